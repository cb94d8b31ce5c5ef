"""ARQ-MCMC -- host-side mirror of the caller of the particle-filter closure: run_arq_mcmc_analysis
(src/DiscretePOMP.jl:306-353), run_inner_mcmc_analysis (src/arq_main.jl:48-75), arq_met_hastings! and get_grid_point!
(src/arq_alg_std.jl:4-88) with the helpers of src/arq_alg_cmn.jl.

The reference runs its chains one after the other and evaluates `model.pdf(theta)` (the particle-filter log-likelihood
closure of get_log_pdf_fn, src/hmm_particle_filter.jl:87-101) one grid point at a time.  Here the chains advance in lock
step over the same shared grid cache, and the grid points that need an evaluation in a step are sent to the GPU as ONE
batched dpomp_pf_loglik call.  Requests for the same grid point within a step are served in chain order, the later ones
after the update of the earlier one, exactly as consecutive calls of get_grid_point! would be.  Stated departure: chain
k + 1 does not start from the finished cache of chain k (lock step), which changes the order of cache fills, not the
algorithm.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .ibis import compute_is_mu_covar
from .mcmc import C_DF_MCMC_ADAPT, C_DF_MCMC_STEPS, gelman_diagnostic_sre, handle_rej_samples
from .particle_filter import get_log_pdf_fn, get_private_model
from .structs import DPOMPModel, HiddenMarkovModel, ImportanceSample, RejectionSample

C_DF_ARQ_SL = 1     # sample limit (src/arq_main.jl:9-13)
C_DF_ARQ_SR = 50    # initial sample distribution
C_DF_ARQ_MC = 5     # chains
C_DF_ARQ_AR = 0.33  # targeted acceptance rate
C_DF_ARQ_JT = 0.0   # jitter
Q_JUMP = 0.1        # src/arq_alg_cmn.jl:54-57
Q_J_MIN = 2
N_ADAPT_PERIODS = 100
C_DF_ARQ_CJ = 10
Q_BI_SAMPLE_LIM = 1      # src/arq_alg_std.jl:5
Q_REJECT_TRIGGER = 100   # src/arq_alg_std.jl:64


def df_adapt_period(steps: int) -> int:
    return int(math.floor(steps * C_DF_MCMC_ADAPT))


@dataclass
class ARQModel:
    """ARQModel (src/arq_structs.jl:13-17).  `pdf` maps a parameter vector to a log density; when it also accepts an
    (n_theta, B) matrix (attribute `batched = True`, as the closures of get_log_pdf_fn do) grid points are evaluated in
    batches."""
    pdf: Callable
    sample_interval: np.ndarray
    sample_offset: np.ndarray


@dataclass
class LikelihoodModel:
    """LikelihoodModel (src/arq_structs.jl:20-28)."""
    pdf: Callable
    sample_interval: np.ndarray
    sample_offset: np.ndarray
    sample_limit: int
    sample_dispersal: int
    jitter: float
    prior: Callable


@dataclass
class GridPoint:
    """GridPoint (src/arq_structs.jl:33-38)."""
    sample: np.ndarray
    log_likelihood: float
    visited: int
    sampled: int


@dataclass
class GridRequest:
    """GridRequest (src/arq_structs.jl:40-44)."""
    result: GridPoint
    prior: float
    process_run: bool


@dataclass
class ARQMCMCSample:
    """ARQMCMCSample (src/arq_structs.jl:80-91)."""
    imp_sample: ImportanceSample
    samples: RejectionSample
    sample_interval: np.ndarray
    sample_limit: int
    sample_dispersal: int
    adapt_period: int
    sre: np.ndarray
    run_time: int
    fx: np.ndarray
    sample_cache: Dict[Tuple[int, ...], GridPoint] = field(repr=False, default_factory=dict)


def get_arq_prior(priord) -> Callable:
    """get_arq_prior (src/arq_alg_cmn.jl:16-21)."""
    return lambda theta: float(priord.logpdf(np.asarray(theta, dtype=np.float64)))


def get_theta_val(model: LikelihoodModel, theta: Sequence[int], rng: np.random.Generator) -> np.ndarray:
    """get_theta_val (src/arq_alg_cmn.jl:24-32): realise a grid index as a parameter value."""
    out = model.sample_offset + np.asarray(theta, dtype=np.float64) * model.sample_interval
    if model.jitter > 0.0:
        out = out + ((rng.random(len(out)) * 2) - 1) * model.jitter * model.sample_interval
    return out


def get_theta_f(theta_i: np.ndarray, j_w: np.ndarray, max_dist: int, min_dist: int, rng: np.random.Generator) -> np.ndarray:
    """get_theta_f (src/arq_alg_cmn.jl:36-45): random walk of `d` unit steps on the grid, axes drawn with weights j_w."""
    n = len(theta_i)
    out = np.zeros(n, dtype=np.int64)
    d = max_dist if min_dist == max_dist else int(rng.integers(min_dist, max_dist + 1))
    if d == 0:
        return out + theta_i
    cw = np.cumsum(j_w)
    block = 64
    while True:  # the reference's `while sum(abs.(output)) != d` walk, generated in blocks of steps
        axes = np.minimum(np.searchsorted(cw, rng.random(block) * cw[-1], side="right"), n - 1)
        steps = np.zeros((block, n), dtype=np.int64)
        steps[np.arange(block), axes] = np.where(rng.random(block) < 0.5, 1, -1)
        pos = out + np.cumsum(steps, axis=0)
        hit = np.nonzero(np.abs(pos).sum(axis=1) == d)[0]
        if len(hit):
            return pos[hit[0]] + theta_i
        out = pos[-1]
        block = min(block * 2, 4096)


def adapt_jw(j_w: np.ndarray, lar_j: int, j: int, mc_accepted: np.ndarray, a_h: int, i: int, tgt_ar: float,
             mc_idx: np.ndarray) -> int:
    """adapt_jw! (src/arq_alg_cmn.jl:60-86); `i` is Julia's 1-based step, arrays are 0-based.  Updates j_w in place."""
    recent = int(mc_accepted[i - a_h:i].sum())
    if j == Q_J_MIN and recent == 0:
        j = int(round(C_DF_ARQ_CJ * (i / a_h))) if int(mc_accepted[:i].sum()) == 1 else lar_j
    else:
        j = max(int(round(j * ((recent / a_h) / tgt_ar))), Q_J_MIN)
    sd = mc_idx[:, :i].std(axis=1, ddof=1)
    if sd.sum() == 0.0:
        sd[:] = 1.0
    else:
        sd[sd == 0.0] = sd[sd > 0.0].min()
    j_w[:] = sd
    return j


def _eval_pdf(model: LikelihoodModel, vals: List[np.ndarray]) -> np.ndarray:
    if getattr(model.pdf, "batched", False):
        nb = int(getattr(model.pdf, "n_batch", len(vals)))
        out = [np.asarray(model.pdf(np.stack(vals[k:k + nb], axis=1)), dtype=np.float64) for k in range(0, len(vals), nb)]
        return np.concatenate(out)
    return np.array([float(model.pdf(v)) for v in vals], dtype=np.float64)


def get_grid_points(grid: Dict, thetas: List[np.ndarray], model: LikelihoodModel, burn_in: Sequence[bool],
                    rng: np.random.Generator) -> List[GridRequest]:
    """get_grid_point! (src/arq_alg_std.jl:4-41) for a list of requests, in list order; the density evaluations of a
    round go to `model.pdf` as one batch."""
    n = len(thetas)
    out: List[Optional[GridRequest]] = [None] * n
    pending = list(range(n))
    while pending:
        need, deferred = {}, []
        for r in pending:
            key = tuple(int(v) for v in thetas[r])
            if key in need:  # served after the update of the earlier request of this round
                deferred.append(r)
                continue
            x = grid.get(key)
            exists = x is not None
            visited, sampled, theta_val = (x.visited, x.sampled, x.sample) if exists else (0, 0, get_theta_val(model, key, rng))
            pr = model.prior(theta_val)
            if pr == -np.inf:
                out[r] = GridRequest(GridPoint(theta_val, -np.inf, visited, sampled), pr, False)
            elif visited < (Q_BI_SAMPLE_LIM if burn_in[r] else model.sample_limit):
                need[key] = (r, theta_val, pr, x, visited, sampled)
            else:
                pt = GridPoint(theta_val, x.log_likelihood, visited, sampled + (0 if burn_in[r] else 1))
                grid[key] = pt
                out[r] = GridRequest(pt, pr, False)
        if need:
            keys = list(need)
            lls = _eval_pdf(model, [need[k][1] for k in keys])
            for key, log_like in zip(keys, lls):
                r, theta_val, pr, x, visited, sampled = need[key]
                log_like = float(log_like)
                if x is not None:  # running mean of the density estimate in the linear domain (:27)
                    with np.errstate(divide="ignore", invalid="ignore"):
                        log_like = float(np.log(np.exp(x.log_likelihood) + ((np.exp(log_like) - np.exp(x.log_likelihood)) / visited)))
                pt = GridPoint(theta_val, log_like, visited + 1, sampled + (0 if burn_in[r] else 1))
                grid[key] = pt
                out[r] = GridRequest(pt, pr, True)
        pending = deferred
    return out  # type: ignore[return-value]


def get_initial_samples(model: LikelihoodModel, grid: Dict, n_chains: int, rng: np.random.Generator):
    """get_initial_sample (src/arq_alg_cmn.jl:105-112) for every chain: uniform grid coordinates in 1..sample_dispersal,
    widened by one per retry while the prior rejects the point."""
    d = len(model.sample_interval)
    theta = [None] * n_chains
    x0: List[Optional[GridRequest]] = [None] * n_chains
    disp = [model.sample_dispersal] * n_chains
    todo = list(range(n_chains))
    while todo:
        for c in todo:
            theta[c] = rng.integers(1, disp[c] + 1, size=d).astype(np.int64)
        res = get_grid_points(grid, [theta[c] for c in todo], model, [True] * len(todo), rng)
        nxt = []
        for c, rq in zip(todo, res):
            if rq.prior == -np.inf:
                disp[c] += 1
                nxt.append(c)
            else:
                x0[c] = rq
        todo = nxt
    return theta, x0


def arq_met_hastings(samples: np.ndarray, grid: Dict, model: LikelihoodModel, steps: int, adapt_period: int, tgt_ar: float,
                     rng: np.random.Generator):
    """arq_met_hastings! (src/arq_alg_std.jl:44-88) for all chains of `samples` (n_theta, steps, chains) in lock step.
    Returns per chain (calls to f(theta), acceptance rate, adapted acceptance rate)."""
    d, _, n_chains = samples.shape
    # @init_inner_mcmc (src/arq_alg_cmn.jl:115-142)
    mc_fx = np.zeros((n_chains, 3), dtype=np.int64)
    theta_i, xi = get_initial_samples(model, grid, n_chains, rng)
    for c in range(n_chains):
        mc_fx[c, 0] += int(xi[c].process_run)
    lar_j = int(round(0.2 * model.sample_dispersal * d))
    a_h = int(max(steps / N_ADAPT_PERIODS, 100))
    j = [int(round(Q_JUMP * model.sample_dispersal * d))] * n_chains
    j_w = [np.ones(d) for _ in range(n_chains)]
    mc_idx = np.zeros((n_chains, d, steps), dtype=np.int64)
    mc_accepted = np.zeros((n_chains, steps), dtype=bool)
    for c in range(n_chains):
        samples[:, 0, c] = xi[c].result.sample
        mc_idx[c, :, 0] = theta_i[c]
        mc_accepted[c, 0] = True
    for i in range(2, steps + 1):  # Julia's 1-based step index
        theta_f = [get_theta_f(theta_i[c], j_w[c], j[c], 1, rng) for c in range(n_chains)]
        xf = get_grid_points(grid, theta_f, model, [i < a_h] * n_chains, rng)  # sample limit 1 for the first interval only
        refresh = []
        for c in range(n_chains):
            mc_fx[c, 1] += int(xf[c].process_run)
            lr = xf[c].prior - xi[c].prior + xf[c].result.log_likelihood - xi[c].result.log_likelihood
            mh_prob = math.inf if lr > 700.0 else (math.exp(lr) if lr == lr else math.nan)  # exp(-Inf) = 0, NaN stays NaN
            acc = mh_prob > 1.0 or mh_prob > rng.random()
            mc_accepted[c, i - 1] = acc
            if acc:
                samples[:, i - 1, c] = xf[c].result.sample
                mc_idx[c, :, i - 1] = theta_f[c]
                theta_i[c], xi[c] = theta_f[c], xf[c]
            else:
                samples[:, i - 1, c] = samples[:, i - 2, c]
                mc_idx[c, :, i - 1] = mc_idx[c, :, i - 2]
                if i > Q_REJECT_TRIGGER and not mc_accepted[c, i - Q_REJECT_TRIGGER - 1:i].any():
                    refresh.append(c)  # refresh the current point after Q_REJECT_TRIGGER rejections in a row
        if refresh:
            res = get_grid_points(grid, [theta_i[c] for c in refresh], model, [False] * len(refresh), rng)
            for c, rq in zip(refresh, res):
                xi[c] = rq
                mc_fx[c, 2] += int(rq.process_run)
        if i % a_h == 0:
            for c in range(n_chains):
                j[c] = adapt_jw(j_w[c], lar_j, j[c], mc_accepted[c], a_h, i, tgt_ar, mc_idx[c])
    return [(int(mc_fx[c].sum()), mc_accepted[c].sum() / steps, mc_accepted[c, adapt_period:].sum() / max(steps - adapt_period, 1))
            for c in range(n_chains)]


def collect_theta_weight(grid: Dict, n_theta: int):
    """collect_theta_weight (src/arq_utils.jl:6-14)."""
    theta = np.zeros((n_theta, len(grid)))
    w = np.zeros(len(grid))
    for k, pt in enumerate(grid.values()):
        theta[:, k] = pt.sample
        w[k] = np.exp(pt.log_likelihood)
    return theta, w


def run_inner_mcmc_analysis(mdl: LikelihoodModel, steps: int, burnin: int, chains: int, tgt_ar: float, grid: Dict,
                            rng: np.random.Generator, verbose: bool = True) -> ARQMCMCSample:
    """run_inner_mcmc_analysis (src/arq_main.jl:48-75) with the standard (non data-augmented) algorithm."""
    start_time = time.time_ns()
    n_theta = len(mdl.sample_interval)
    samples = np.zeros((n_theta, steps, chains))
    info = arq_met_hastings(samples, grid, mdl, steps, burnin, tgt_ar, rng)
    fx = np.array([m[0] for m in info], dtype=np.int64)
    rejs = handle_rej_samples(samples, burnin)
    sre = gelman_diagnostic_sre(samples, burnin)
    theta, w = collect_theta_weight(grid, n_theta)
    is_mu, cv = compute_is_mu_covar(theta, w)
    with np.errstate(divide="ignore"):
        bme = np.array([-np.log(w.sum() / len(w)), -np.log(w.sum() / (len(w) ** (1 / n_theta)))])
    imp = ImportanceSample(is_mu, cv, theta, w, 0, bme)
    out = ARQMCMCSample(imp, rejs, mdl.sample_interval, mdl.sample_limit, mdl.sample_dispersal, burnin, sre,
                        time.time_ns() - start_time, fx, grid)
    out.acceptance = np.array([[m[1], m[2]] for m in info])
    if verbose:
        print(f"- finished in {out.run_time / 1e9:.1f} seconds. (Iμ = {is_mu}; Rμ = {rejs.mu}; BME = {bme[0]:.4g}; "
              f"calls to f(θ) := {fx.tolist()})")
    return out


def run_arq_mcmc_analysis(model, obs_data_or_interval, sample_interval=None, sample_offset=None,
                          sample_dispersal: int = C_DF_ARQ_SR, sample_limit: int = C_DF_ARQ_SL, n_chains: int = C_DF_ARQ_MC,
                          steps: int = C_DF_MCMC_STEPS, burnin: Optional[int] = None, tgt_ar: float = C_DF_ARQ_AR,
                          np_: int = 200, ess_crit: float = 0.3, jitter: float = C_DF_ARQ_JT, sample_cache: Optional[Dict] = None,
                          priors: Optional[Sequence[Callable]] = None, seed: int = 1, verbose: bool = True, **pf_kw):
    """run_arq_mcmc_analysis in its three reference forms:
        (model::DPOMPModel, obs_data, sample_interval; ...)      (src/DiscretePOMP.jl:342-353)
        (model::HiddenMarkovModel, sample_interval; ...)         (src/DiscretePOMP.jl:306-317)
        (model::ARQModel, priors; ...)                           (src/arq_main.jl:96-110)
    The particle-filter closure is get_log_pdf_fn(model, np; essc = ess_crit) with room for `n_chains` filters per call."""
    rng = np.random.default_rng(seed)
    if isinstance(model, ARQModel):
        arq = model
        prior_fns = list(priors if priors is not None else obs_data_or_interval)
    else:
        if isinstance(model, DPOMPModel):
            hmm = get_private_model(model, obs_data_or_interval)
        else:
            hmm, sample_interval = model, (obs_data_or_interval if sample_interval is None else sample_interval)
        interval = np.asarray(sample_interval, dtype=np.float64)
        offset = interval / 2 if sample_offset is None else np.asarray(sample_offset, dtype=np.float64)
        pdf = get_log_pdf_fn(hmm, np_, essc=ess_crit, seed=seed, n_batch=max(n_chains, 1), **pf_kw)
        pdf.batched, pdf.n_batch = True, max(n_chains, 1)
        arq = ARQModel(pdf, interval, offset)
        prior_fns = [get_arq_prior(hmm.prior)]
        if verbose:
            print(f"ARQ model initialised: {hmm.model_name}")
    if burnin is None:
        burnin = df_adapt_period(steps)
    grid = {} if sample_cache is None else sample_cache
    output = []
    for k, pr in enumerate(prior_fns):
        if verbose:
            print(f"Running: ARQMCMC analysis {'' if len(prior_fns) == 1 else f'{k + 1} / {len(prior_fns)} -'} ({n_chains} x {steps} steps):")
        mdl = LikelihoodModel(arq.pdf, np.asarray(arq.sample_interval, dtype=np.float64),
                              np.asarray(arq.sample_offset, dtype=np.float64), sample_limit, sample_dispersal, jitter, pr)
        output.append(run_inner_mcmc_analysis(mdl, steps, burnin, n_chains, tgt_ar, grid, rng, verbose))
    return output[0] if len(output) == 1 else output
