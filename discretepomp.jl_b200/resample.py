"""Index-returning resamplers of the outer (IBIS) layer -- mirror of src/hmm_resample.jl:4-20,44-62,66-83.

The weights are cumulated sequentially in f64 on the host (bit-exact `cumsum`), the search runs on the GPU through
dpomp_resample_indices.  As in the reference, `w` is overwritten with its cumulative sum by rs_stratified /
rs_multinomial (`cumsum!`) but not by rs_systematic (`cumsum`)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _capi


def _search(rs_type: int, w: np.ndarray, u: np.ndarray, n_out: int, on_cumulative: bool, device: int = -1) -> np.ndarray:
    w = _capi.as_f64(w)
    u = _capi.as_f64(np.atleast_1d(u))
    out = np.empty(n_out, dtype=np.int64)
    _capi.check(_capi.lib().dpomp_resample_indices(rs_type, 1 if on_cumulative else 0, _capi.ptr(w), len(w),
                                                   _capi.ptr(u), len(u), n_out, _capi.ptr(out), device))
    return out


def rs_systematic(w: np.ndarray, rng: Optional[np.random.Generator] = None, u: Optional[float] = None) -> np.ndarray:
    """rs_systematic (src/hmm_resample.jl:44-62): 1-based ancestors; consumes one rand()."""
    if u is None:
        u = (rng or np.random.default_rng()).random()
    return _search(_capi.RS_SYSTEMATIC, w, np.asarray([u]), len(w), False)


def rs_stratified(w: np.ndarray, rng: Optional[np.random.Generator] = None, u: Optional[np.ndarray] = None) -> np.ndarray:
    """rs_stratified (src/hmm_resample.jl:66-83): consumes length(w) rand() draws; w becomes cumulative."""
    if u is None:
        u = (rng or np.random.default_rng()).random(len(w))
    out = _search(_capi.RS_STRATIFIED, w, u, len(w), False)
    np.cumsum(w, out=w)
    return out


def rs_multinomial(w: np.ndarray, n: Optional[int] = None, rng: Optional[np.random.Generator] = None,
                   u: Optional[np.ndarray] = None) -> np.ndarray:
    """rs_multinomial (src/hmm_resample.jl:4-20): n offspring (default length(w)); w becomes cumulative."""
    n = len(w) if n is None else int(n)
    if u is None:
        u = (rng or np.random.default_rng()).random(n)
    out = _search(_capi.RS_MULTINOMIAL, w, u, n, False)
    np.cumsum(w, out=w)
    return out


def rsp_indices(rs_type: int, cw: np.ndarray, u: np.ndarray, n_out: Optional[int] = None) -> np.ndarray:
    """Ancestors chosen by rsp_systematic / rsp_stratified / rsp_multinomial (src/hmm_pf_resample.jl) on CUMULATIVE
    weights `cw`, given the raw rand() draws `u`."""
    return _search(rs_type, cw, u, len(cw) if n_out is None else n_out, True)
