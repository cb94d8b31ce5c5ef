"""ctypes wrapper of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
MODE_LITERAL, MODE_DEVICE, MODE_DEVICE_INTERLEAVED = 0, 1, 2
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("dpomp_oracle.c", "dpomp_oracle_ibis.c", "dpomp_oracle_mbp.c", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_compute_ess.restype = C.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def philox(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32); k = np.asarray(key, dtype=np.uint32); out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(c), _p(k), _p(out))
    return out


def pf_partial(desc, theta, n: int, pop: Optional[np.ndarray], ymin: int, ymax: int, rs_type: int = 1, key: int = 0,
               filter_id: int = 0, mode: int = MODE_LITERAL, tile: int = 1024, items: int = 4,
               max_events: int = 1 << 20, threads: int = 1):
    """partial_log_likelihood!; pop is (n, C) int64 (Julia Matrix) and is updated in place.  Returns
    (loglik, logw_last, ancestors_last(1-based), n_events, n_overflow)."""
    th = np.ascontiguousarray(theta, dtype=np.float64)
    c = desc.n_compartments
    popf = np.zeros((c, n), dtype=np.int64) if pop is None else np.ascontiguousarray(np.asarray(pop, dtype=np.int64).T)
    ll = C.c_double(); ev = C.c_int64(); ovf = C.c_int64()
    lw = np.zeros(n); anc = np.zeros(n, dtype=np.int64)
    rc = lib().orc_pf_partial(C.byref(desc), _p(th), C.c_int64(n), _p(popf), ymin, ymax, rs_type, C.c_uint64(key),
                              C.c_uint32(filter_id), mode, tile, items, C.c_int64(max_events), threads, C.byref(ll),
                              _p(lw), _p(anc), C.byref(ev), C.byref(ovf))
    assert rc == 0
    if pop is not None:
        pop[...] = popf.T
    return ll.value, lw, anc, ev.value, ovf.value, popf.T.copy()


def pf_loglik(desc, theta, n: int, rs_type: int = 1, key: int = 0, filter_id: int = 0, mode: int = MODE_LITERAL,
              tile: int = 1024, items: int = 4, max_events: int = 1 << 20, threads: int = 1) -> Tuple[float, int]:
    th = np.ascontiguousarray(theta, dtype=np.float64)
    ll = C.c_double(); ev = C.c_int64()
    rc = lib().orc_pf_loglik(C.byref(desc), _p(th), C.c_int64(n), rs_type, C.c_uint64(key), C.c_uint32(filter_id), mode,
                             tile, items, C.c_int64(max_events), threads, C.byref(ll), C.byref(ev))
    assert rc == 0
    return ll.value, ev.value


def pf_partial_batch(desc, theta_cols, n: int, ymin: int, ymax: int, rs_type: int = 1, key: int = 0, filter0: int = 0,
                     mode: int = MODE_LITERAL, tile: int = 1024, items: int = 4, max_events: int = 1 << 20,
                     threads: int = 1, pops: Optional[np.ndarray] = None):
    """theta_cols: (n_theta, B) Julia layout.  pops: optional (B, C, n) int64 state (updated in place)."""
    th = np.ascontiguousarray(np.asarray(theta_cols, dtype=np.float64).T)
    nb = th.shape[0]
    out = np.zeros(nb); ev = C.c_int64()
    rc = lib().orc_pf_partial_batch(C.byref(desc), _p(th), nb, C.c_int64(n), _p(pops), ymin, ymax, rs_type,
                                    C.c_uint64(key), C.c_uint32(filter0), mode, tile, items, C.c_int64(max_events),
                                    threads, _p(out), C.byref(ev))
    assert rc == 0
    return out, ev.value


def rs(rs_type: int, w, r, n_out: Optional[int] = None) -> np.ndarray:
    """rs_* on RAW weights (src/hmm_resample.jl); returns 1-based ancestors."""
    w = np.array(w, dtype=np.float64); r = np.ascontiguousarray(np.atleast_1d(r), dtype=np.float64)
    n_out = len(w) if n_out is None else n_out
    out = np.zeros(n_out, dtype=np.int64)
    lib().orc_rs(rs_type, _p(w), C.c_int64(len(w)), _p(r), C.c_int64(n_out), _p(out))
    return out


def rsp(rs_type: int, cw, r, n_out: Optional[int] = None) -> np.ndarray:
    """rsp_* on CUMULATIVE weights (src/hmm_pf_resample.jl); returns 1-based ancestors."""
    cw = np.ascontiguousarray(cw, dtype=np.float64); r = np.ascontiguousarray(np.atleast_1d(r), dtype=np.float64)
    n_out = len(cw) if n_out is None else n_out
    out = np.zeros(n_out, dtype=np.int64)
    lib().orc_rsp(rs_type, _p(cw), C.c_int64(len(cw)), _p(r), C.c_int64(n_out), _p(out))
    return out


def compute_ess(w) -> float:
    w = np.ascontiguousarray(w, dtype=np.float64)
    return lib().orc_compute_ess(_p(w), C.c_int64(len(w)))


def compute_is_mu_covar(theta_cols, w):
    th = np.ascontiguousarray(np.asarray(theta_cols, dtype=np.float64).T)  # (n, n_theta) row-major == Julia col-major
    n, nt = th.shape
    w = np.ascontiguousarray(w, dtype=np.float64)
    mu = np.zeros(nt); cv = np.zeros((nt, nt))
    lib().orc_compute_is_mu_covar(_p(mu), _p(cv), _p(th), _p(w), nt, C.c_int64(n))
    return mu, cv


def gillespie_sim(desc, theta, key: int = 0, max_events: int = 1 << 24):
    th = np.ascontiguousarray(theta, dtype=np.float64)
    out = np.zeros((desc.n_obs, desc.n_compartments), dtype=np.int64); ev = C.c_int64()
    lib().orc_gillespie_sim(C.byref(desc), _p(th), C.c_uint64(key), C.c_int64(max_events), _p(out), C.byref(ev))
    return out, ev.value


def cum_rates(desc, theta, x) -> np.ndarray:
    """rate_function + cumsum! (src/hmm_particle_filter.jl:20-21) at one state."""
    th = np.ascontiguousarray(theta, dtype=np.float64); xs = np.ascontiguousarray(x, dtype=np.int64)
    out = np.zeros(desc.n_events)
    lib().orc_cum_rates(C.byref(desc), _p(th), _p(xs), _p(out))
    return out


def choose_event(cum, u: float) -> int:
    """choose_event (src/hmm_cmn.jl:4-10) with the rand() draw given; 1-based event."""
    c = np.ascontiguousarray(cum, dtype=np.float64)
    return int(lib().orc_choose_event(_p(c), len(c), C.c_double(u)))


def obs_model(desc, t: int, x) -> float:
    """gom2 (src/hmm_examples.jl:59-67) for observation t (0-based) at state x."""
    xs = np.ascontiguousarray(x, dtype=np.int64)
    f = lib().orc_obs_model
    f.restype = C.c_double
    return float(f(C.byref(desc), int(t), _p(xs)))


def max_threads() -> int:
    return int(lib().orc_max_threads())


def philox2(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32); out = np.zeros(2, dtype=np.uint32)
    lib().orc_philox2x32_10(_p(c), C.c_uint32(key), _p(out))
    return out


def run_pibis(desc, theta_init, prior_lo, prior_hi, ess_rs_crit=0.3, ind_prop=True, alpha=1.002, npf=200, n_props=1,
              seed=1, threads=1, max_events=1 << 20):
    """run_pibis (src/hmm_ibis.jl:12-135).  theta_init: (n_theta, outer_p).  Returns dict(mu, cv, theta, w, bme, k_log, pf_steps)."""
    th = np.ascontiguousarray(np.asarray(theta_init, dtype=np.float64).T).copy()  # (outer_p, n_theta) == Julia col-major
    outer_p, d = th.shape
    lo = np.ascontiguousarray(prior_lo, dtype=np.float64); hi = np.ascontiguousarray(prior_hi, dtype=np.float64)
    mu = np.zeros(d); cv = np.zeros((d, d)); w = np.zeros(outer_p); bme = np.zeros(2)
    k_log = np.zeros(2, dtype=np.int64); steps = C.c_int64()
    rc = lib().orc_run_pibis(C.byref(desc), _p(th), C.c_int64(outer_p), _p(lo), _p(hi), C.c_double(ess_rs_crit),
                             int(bool(ind_prop)), C.c_double(alpha), C.c_int64(npf), int(n_props), C.c_uint64(seed),
                             int(threads), C.c_int64(max_events), _p(mu), _p(cv), _p(w), _p(bme), _p(k_log), C.byref(steps))
    assert rc == 0
    return dict(mu=mu, cv=cv, theta=th.T.copy(), w=w, bme=bme, k_log=k_log, pf_steps=steps.value)


def run_pmcmc(desc, theta_init, steps, adapt_period, npf, prior_lo, prior_hi, c_initial=0.1, seed=1, threads=1,
              max_events=1 << 20):
    """pMCMC per src/hmm_mcmc.jl:349-365,166-211.  theta_init: (n_theta, chains).  Returns samples (n_theta, steps, chains)."""
    th0 = np.ascontiguousarray(np.asarray(theta_init, dtype=np.float64).T)
    n_chains, d = th0.shape
    lo = np.ascontiguousarray(prior_lo, dtype=np.float64); hi = np.ascontiguousarray(prior_hi, dtype=np.float64)
    samples = np.zeros((n_chains, steps, d)); acc = np.zeros(n_chains, dtype=np.int64)
    rc = lib().orc_run_pmcmc(C.byref(desc), _p(th0), n_chains, int(steps), int(adapt_period), C.c_int64(npf), _p(lo), _p(hi),
                             C.c_double(c_initial), C.c_uint64(seed), int(threads), C.c_int64(max_events), _p(samples), _p(acc))
    assert rc == 0
    return np.ascontiguousarray(samples.transpose(2, 1, 0)), acc


def mbp_iterate(desc, theta, fc, ev_time, ev_type, length, log_like, t_start, obs_i, key, pid):
    """iterate_particle! (src/hmm_sim.jl:6-25) on flat arrays (updated in place).  Returns (log g, new length)."""
    th = np.ascontiguousarray(theta, dtype=np.float64)
    ln = C.c_int64(int(length))
    f = lib().orc_mbp_iterate
    f.restype = C.c_double
    g = f(C.byref(desc), _p(th), _p(fc), _p(ev_time), _p(ev_type), C.byref(ln), C.c_int64(len(ev_time)), _p(log_like),
          C.c_double(t_start), int(obs_i), C.c_uint64(key), C.c_uint32(pid))
    return g, ln.value


def mbp_propose(desc, theta_i, theta_f, xi_time, xi_type, xi_len, cap, ymax, key, pid):
    """partial_model_based_proposal (src/hmm_mbp.jl:83-108).  Returns (times, types(1-based), fc, log_like[2], overflow)."""
    ti = np.ascontiguousarray(theta_i, dtype=np.float64); tf = np.ascontiguousarray(theta_f, dtype=np.float64)
    xt = np.ascontiguousarray(xi_time, dtype=np.float64); xy = np.ascontiguousarray(xi_type, dtype=np.int32)
    ot = np.zeros(cap); oy = np.zeros(cap, dtype=np.int32); ln = C.c_int64(0)
    fc = np.zeros(desc.n_compartments, dtype=np.int64); ll = np.zeros(2)
    rc = lib().orc_mbp_propose(C.byref(desc), _p(ti), _p(tf), _p(xt), _p(xy), C.c_int64(int(xi_len)), _p(ot), _p(oy), C.byref(ln),
                               C.c_int64(cap), _p(fc), _p(ll), int(ymax), C.c_uint64(key), C.c_uint32(pid))
    return ot[: ln.value].copy(), oy[: ln.value].copy(), fc, ll, rc


def run_mbp_ibis(desc, theta_init, prior_lo, prior_hi, ess_rs_crit=0.5, n_props=3, ind_prop=False, alpha=1.002, rs_type=1,
                 cap=4096, seed=1, threads=1):
    """run_mbp_ibis (src/hmm_ibis.jl:140-244).  Returns dict(mu, cv, theta, w, bme, k_log)."""
    th = np.ascontiguousarray(np.asarray(theta_init, dtype=np.float64).T).copy()
    outer_p, d = th.shape
    lo = np.ascontiguousarray(prior_lo, dtype=np.float64); hi = np.ascontiguousarray(prior_hi, dtype=np.float64)
    mu = np.zeros(d); cv = np.zeros((d, d)); w = np.zeros(outer_p); bme = np.zeros(2); k_log = np.zeros(2, dtype=np.int64)
    rc = lib().orc_run_mbp_ibis(C.byref(desc), _p(th), C.c_int64(outer_p), _p(lo), _p(hi), C.c_double(ess_rs_crit), int(n_props),
                                int(bool(ind_prop)), C.c_double(alpha), int(rs_type), C.c_int64(cap), C.c_uint64(seed), int(threads),
                                _p(mu), _p(cv), _p(w), _p(bme), _p(k_log))
    assert rc == 0
    return dict(mu=mu, cv=cv, theta=th.T.copy(), w=w, bme=bme, k_log=k_log)
