"""Iterated batch importance sampling: SMC^2 -- host-side mirror of run_pibis (src/hmm_ibis.jl:12-135) and of the
public wrapper run_ibis_analysis (src/DiscretePOMP.jl:289-302).

The reference runs one particle filter per theta-particle in plain `for` loops (src/hmm_ibis.jl:53-56, 83-116).  Here
the three partial_log_likelihood! call sites become BATCHED C-ABI calls over all theta-particles of this rank
(dpomp_pf_partial), the population gathers of the resample step become dpomp_pf_permute / dpomp_pf_copy_filters, and
everything else (priors, MvNormal proposals, accept/reject, evidence bookkeeping) stays host code, as in the reference.

Stated departure: the mutation sweep is evaluated for all theta-particles at once, so the random-walk
scale `tj` is frozen within a sweep and updated afterwards with the same accept/reject factors.  With ind_prop = true
(the default of run_ibis_analysis for SMC^2) `tj` is not used and the sweep is exactly the reference's.
"""
from __future__ import annotations

import time
from typing import Callable, Optional

import numpy as np

from . import _capi
from .distributed import Comm, migration_plan
from .particle_filter import ParticleFilter, compute_ess, device_model, get_private_model
from .resample import rs_stratified, rs_systematic
from .structs import DPOMPModel, HiddenMarkovModel, ImportanceSample

_M64 = 0xFFFFFFFFFFFFFFFF


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


C_ALG_NM_SMC2 = "SMC2"  # src/DiscretePOMP.jl:37-38
C_ALG_NM_MBPI = "MBPI"
C_DF_MBPI_P = 10000  # src/DiscretePOMP.jl:46-52
C_DF_SMC2_P = 4000
C_DF_PF_P = 200
C_DF_ESS_CRIT = 0.3
C_DF_MBPI_ESS_CRIT = 0.5
C_DF_MBPI_MUT = 3
C_ACCEPTANCE_ALPHA = 1.002  # src/DiscretePOMP.jl:43


def prior_logpdf_columns(prior, theta: np.ndarray) -> np.ndarray:
    """logpdf(prior, theta[:, i]) for every column (src/hmm_ibis.jl:22-24, :88); vectorised when the prior offers it."""
    batch = getattr(prior, "logpdf_batch", None)
    if batch is not None:
        return np.asarray(batch(theta), dtype=np.float64)
    return np.array([prior.logpdf(theta[:, i]) for i in range(theta.shape[1])], dtype=np.float64)


def compute_is_mu_covar(theta: np.ndarray, w: np.ndarray):
    """compute_is_mu_covar! (src/cmn.jl:91-99): weighted mean and (biased) covariance; theta is (n_theta, n)."""
    sw = np.sum(w)
    mu = (theta * w).sum(axis=1) / sw
    d = theta - mu[:, None]
    cv = (d * w) @ d.T / sw
    return mu, cv


class ProposalDensity:
    """Zero-mean MvNormal held as its Cholesky factor (Distributions.MvNormal(cov))."""

    def __init__(self, chol: np.ndarray):
        self.chol = chol

    @staticmethod
    def identity(d: int) -> "ProposalDensity":
        return ProposalDensity(np.eye(d))

    def rand(self, rng: np.random.Generator, n: int) -> np.ndarray:
        return self.chol @ rng.standard_normal((self.chol.shape[0], n))


def get_prop_density(cv: np.ndarray, old: ProposalDensity) -> ProposalDensity:
    """get_prop_density (src/hmm_cmn.jl:33-42): MvNormal(cv) if Hermitian(cv) is positive definite, else keep `old`."""
    try:
        sym = np.triu(cv) + np.triu(cv, 1).T  # LinearAlgebra.Hermitian uses the upper triangle
        return ProposalDensity(np.linalg.cholesky(sym))
    except np.linalg.LinAlgError:
        return old


def get_mv_param(propd: ProposalDensity, sclr, theta_i: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    """get_mv_param (src/hmm_cmn.jl:13-18), vectorised over columns of theta_i."""
    theta_i = np.asarray(theta_i, dtype=np.float64)
    n = theta_i.shape[1] if theta_i.ndim == 2 else 1
    out = propd.rand(rng, n) * sclr
    return out + (theta_i if theta_i.ndim == 2 else theta_i[:, None])


class FilterBank:
    """This rank's share of the theta-particles' particle filters: a resident bank (`main`) and a proposal bank."""

    def __init__(self, mdl: HiddenMarkovModel, outer_p: int, npf: int, comm: Comm, seed: int, pf_factory: Optional[Callable],
                 rs_type: int = 1):
        self.comm, self.outer_p = comm, outer_p
        self.lo, self.hi = comm.bounds(outer_p)
        self.n_local = self.hi - self.lo
        make = pf_factory or (lambda nb, sd: ParticleFilter(device_model(mdl), npf, nb, rs_type, seed=sd))
        nb = max(self.n_local, 1)
        self.main = make(nb, seed)
        self.prop = make(nb, seed + 0x5DEECE66D)
        self.main.set_batch_offset(self.lo)
        self.seed, self._call = seed, 0
        self.timers = {}  # seconds per phase on this rank: kernels vs waiting for the other ranks (scaling diagnostics)

    def _tick(self, name: str, t0: float) -> float:
        t1 = time.perf_counter()
        self.timers[name] = self.timers.get(name, 0.0) + (t1 - t0)
        return t1

    def _next_key(self) -> int:
        """Call keys are derived from a counter every rank advances identically, so the random streams do not depend on
        how the theta-particles are sharded."""
        self._call += 1
        return splitmix64((self.seed & _M64) ^ splitmix64(self._call))

    def partial(self, theta: np.ndarray, ymin: int, ymax: int) -> np.ndarray:
        """Batched partial_log_likelihood! for ALL theta-particles (columns of theta); each rank runs its block."""
        key = self._next_key()
        loc = np.zeros(0)
        t0 = time.perf_counter()
        if self.comm.handle is not None and hasattr(self.main, "partial_allgather"):
            # C ABI: kernels -> ncclAllGather -> one copy to the host on the filters' stream, one synchronisation
            if self.n_local:
                self.main.set_stream_key(key)
            out = self.main.partial_allgather(self.comm, theta[:, self.lo:self.hi], ymin, ymax, self.outer_p)
            self._tick("filter_step+allgather", t0)
            return out
        if self.n_local:
            self.main.set_stream_key(key)
            loc = self.main.partial(theta[:, self.lo:self.hi], ymin, ymax)
        t0 = self._tick("filter_step", t0)
        out = self.comm.allgather_f64(loc, self.outer_p)
        self._tick("allgather", t0)
        return out

    def resample(self, nidx: np.ndarray) -> None:
        """pop2[p] .= pop[nidx[p]] (src/hmm_ibis.jl:74) across ranks; nidx is 1-based global."""
        nidx0 = np.asarray(nidx, dtype=np.int64) - 1
        t0 = time.perf_counter()
        if self.comm.world == 1:
            self.main.permute(nidx0 + 1)
            self._tick("resample_local", t0)
            return
        if self.comm.handle is not None and hasattr(self.main, "resample_migrate"):
            # C ABI: plan, pack, grouped ncclSend/ncclRecv, local gather and unpack on the filters' stream
            self.main.resample_migrate(self.comm, nidx0 + 1, self.outer_p)
            self._tick("resample_migrate", t0)
            return
        local_src, send_slots, send_counts, recv_slots, recv_counts = migration_plan(
            nidx0, self.outer_p, self.comm.world, self.comm.rank)
        t0 = self._tick("migration_plan", t0)
        send = self.main.export_tensor(send_slots + 1)
        t0 = self._tick("migration_pack", t0)
        recv = self.comm.all_to_all_blocks(send, send_counts, recv_counts, self.main.filter_words)
        t0 = self._tick("migration_all_to_all", t0)
        if self.n_local:
            self.main.permute(local_src + 1)
        self.main.import_tensor(recv_slots + 1, recv)
        self._tick("resample_local", t0)

    def propose(self, theta_f: np.ndarray, valid: np.ndarray, obs_i: int):
        """Fresh filters for the valid proposals (src/hmm_ibis.jl:90-101).  Returns global (aw_f, gx_f) and the local
        map proposal-slot -> theta-particle used by `accept`."""
        mine = np.nonzero(valid[self.lo:self.hi])[0] + self.lo
        self._prop_owner = mine
        aw_l, gx_l = np.zeros(self.n_local), np.zeros(self.n_local)
        key1, key2 = self._next_key(), self._next_key()
        t0 = time.perf_counter()
        if len(mine):
            th = theta_f[:, mine]
            self.prop.set_filter_ids(mine)  # proposal slot j simulates theta-particle mine[j] (global id)
            if obs_i == 1:
                self.prop.set_stream_key(key1)
                g = self.prop.partial(th, 1, 1)
                a = g.copy()
            else:
                self.prop.set_stream_key(key1)
                a = self.prop.partial(th, 1, obs_i - 1)
                self.prop.set_stream_key(key2)
                g = self.prop.partial(th, obs_i, obs_i)
                a = a + g
            aw_l[mine - self.lo], gx_l[mine - self.lo] = a, g
        t0 = self._tick("proposal_filters", t0)
        out = self.comm.allgather_f64(np.stack([aw_l, gx_l], axis=1), self.outer_p)
        self._tick("allgather", t0)
        return np.ascontiguousarray(out[:, 0]), np.ascontiguousarray(out[:, 1])

    def accept(self, accepted: np.ndarray) -> None:
        """pop[p] .= pop_f for accepted proposals (src/hmm_ibis.jl:108)."""
        mine = self._prop_owner
        if len(mine) == 0:
            return
        t0 = time.perf_counter()
        acc = accepted[mine]
        src = np.nonzero(acc)[0] + 1
        dst = mine[acc] - self.lo + 1
        self.main.copy_from(self.prop, dst, src)
        self._tick("accept_copy", t0)


def run_pibis(model: HiddenMarkovModel, theta: np.ndarray, ess_rs_crit: float, ind_prop: bool, alpha: float, np_: int,
              n_props: int = 1, rng: Optional[np.random.Generator] = None, seed: int = 1, comm: Optional[Comm] = None,
              pf_factory: Optional[Callable] = None, outer_rs: Callable = rs_systematic, verbose: bool = True,
              hastings_correction: bool = False, deal_offspring: bool = True) -> ImportanceSample:
    """run_pibis(model, theta, ess_rs_crit, ind_prop, alpha, np; n_props = 1) (src/hmm_ibis.jl:12-135).
    `theta` is (n_theta, outer_p); `np_` is the number of state particles per filter.
    `hastings_correction` (not in the reference, default off): with independent proposals the reference accepts with
    exp(aw_f - aw) (src/hmm_ibis.jl:104), which leaves out the proposal-density ratio q(theta) / q(theta_f) and biases the
    evidence; True adds it.
    `deal_offspring` (default on): the resampled theta-particles are dealt over the slots with stride 64 instead of being
    stored in ancestor order (`pop2[p] .= pop[nidx[p]]`, :71-79), so that the copies of a heavy ancestor -- filters of equal
    cost -- spread over the ranks of a sharded run instead of landing on one.  The order of theta-particles is arbitrary
    (nothing in run_pibis depends on it but the next systematic resampling, for which any order is valid); the permutation does
    not depend on the number of ranks, so 1, 2, 4 and 8 ranks still agree bit for bit."""
    comm = comm or Comm(None)
    rng = rng or np.random.default_rng(seed)
    theta = np.array(theta, dtype=np.float64, order="C")
    outer_p = theta.shape[1]
    start_time = time.time_ns()
    ess_crit = ess_rs_crit * outer_p
    w = np.ones(outer_p)
    aw = prior_logpdf_columns(model.prior, theta)
    t_setup = time.perf_counter()
    bank = FilterBank(model, outer_p, np_, comm, seed, pf_factory)
    bank.timers["setup_alloc"] = time.perf_counter() - t_setup
    k_log = np.zeros(2, dtype=np.int64)
    bme = np.zeros(2)
    propd = ProposalDensity.identity(theta.shape[0])
    tj = 0.2
    mu, cv = compute_is_mu_covar(theta, w)
    deal = np.argsort(np.arange(outer_p) % 64, kind="stable")  # slot p takes offspring deal[p]: consecutive offspring 1/64 of the slots apart
    obs_min = 1
    for obs_i in range(1, len(model.obs_data) + 1):
        if model.obs_data[obs_i - 1].obs_id > 0:
            gx = bank.partial(theta, obs_min, obs_i)  # :53-56, batched
            aw = aw + gx
            gx = np.exp(gx)
            lml = np.log(np.sum(w * gx) / np.sum(w))
            bme[0] += lml
            w = w * gx
            if compute_ess(w) < ess_crit:
                # compute_is_mu_covar! (:62): the reference evaluates it at every observation, its values are only used here
                # (and after the last observation), so it is evaluated where it is consumed
                mu, cv = compute_is_mu_covar(theta, w)
                propd = get_prop_density(cv, propd)
                nidx = outer_rs(w.copy(), rng)  # 1-based
                if deal_offspring:
                    nidx = nidx[deal]
                theta = theta[:, nidx - 1]
                aw = aw[nidx - 1]
                bank.resample(nidx)
                mlr = np.mean(gx[nidx - 1]) * np.exp(lml)
                k_log[0] += outer_p
                mtd_gx = gx[nidx - 1].copy()
                for _ in range(n_props):  # the sweep :83-116, all theta-particles at once
                    if ind_prop:
                        theta_f = mu[:, None] + propd.rand(rng, outer_p)
                    else:
                        theta_f = get_mv_param(propd, tj, theta, rng)
                    prtf = prior_logpdf_columns(model.prior, theta_f)
                    valid = prtf != -np.inf
                    aw_f, gx_f = bank.propose(theta_f, valid, obs_i)
                    aw_f = aw_f + np.where(valid, prtf, 0.0)
                    u = rng.random(outer_p)
                    log_ratio = aw_f - aw
                    if ind_prop and hastings_correction:  # + log q(theta) - log q(theta_f) for q = N(mu, cov)
                        zc = np.linalg.solve(propd.chol, theta - mu[:, None])
                        zf = np.linalg.solve(propd.chol, theta_f - mu[:, None])
                        log_ratio = log_ratio + 0.5 * ((zf * zf).sum(axis=0) - (zc * zc).sum(axis=0))
                    with np.errstate(over="ignore", invalid="ignore"):
                        accepted = valid & (np.exp(log_ratio) > u)
                    bank.accept(accepted)
                    mtd_gx[accepted] = np.exp(gx_f[accepted])
                    theta[:, accepted] = theta_f[:, accepted]
                    aw[accepted] = aw_f[accepted]
                    n_acc, n_rej = int(accepted.sum()), int((valid & ~accepted).sum())
                    k_log[1] += n_acc
                    tj *= alpha ** n_acc * 0.999 ** n_rej
                bme[1] += np.log(mlr / np.mean(mtd_gx))
                w = np.ones(outer_p)
            else:
                bme[1] += np.log(np.sum(w * gx) / np.sum(w))  # :122, uses the already updated w as written
            obs_min = obs_i + 1
    mu, cv = compute_is_mu_covar(theta, w)
    output = ImportanceSample(mu, cv, theta, w, time.time_ns() - start_time, -bme)
    if verbose and comm.rank == 0:
        ar = 100.0 * k_log[1] / k_log[0] if k_log[0] else float("nan")
        print(f"- finished in {output.run_time / 1e9:.1f} seconds (AR = {ar:.3g}%)")
    output.k_log = k_log
    output.timers = dict(bank.timers)
    return output


def run_ibis_analysis(model: DPOMPModel, obs_data, algorithm: str = C_ALG_NM_SMC2, np: Optional[int] = None,
                      ind_prop: Optional[bool] = None, ess_rs_crit: Optional[float] = None,
                      alpha: float = C_ACCEPTANCE_ALPHA, npf: int = C_DF_PF_P, n_props: int = C_DF_MBPI_MUT,
                      seed: int = 1, comm: Optional[Comm] = None, **kw) -> ImportanceSample:
    """run_ibis_analysis(model, obs_data; algorithm="SMC2", np, ind_prop, ess_rs_crit, alpha, npf, n_props)
    (src/DiscretePOMP.jl:289-302).  Defaults follow the reference: SMC^2 uses 4000 theta-particles x 200 state particles,
    independent proposals and ess_rs_crit = 0.3; "MBPI" uses 10000 particles, ess 0.5, 3 MBP mutations."""
    import numpy as _np

    smc2 = algorithm == C_ALG_NM_SMC2
    n_outer = (C_DF_SMC2_P if smc2 else C_DF_MBPI_P) if np is None else np
    ind_prop = smc2 if ind_prop is None else ind_prop
    ess_rs_crit = (C_DF_ESS_CRIT if smc2 else C_DF_MBPI_ESS_CRIT) if ess_rs_crit is None else ess_rs_crit
    mdl = get_private_model(model, obs_data)
    rng = _np.random.default_rng(seed)
    theta_init = mdl.prior.rand(n_outer, rng)
    if smc2:
        if kw.get("verbose", True) and (comm is None or comm.rank == 0):
            print(f"Running: {n_outer}-particle SMC^2 analysis (model: {model.model_name})")
        return run_pibis(mdl, theta_init, ess_rs_crit, ind_prop, alpha, npf, rng=rng, seed=seed, comm=comm, **kw)
    from .mbp_ibis import run_mbp_ibis

    return run_mbp_ibis(mdl, theta_init, ess_rs_crit, n_props, ind_prop, alpha, rng=rng, seed=seed, comm=comm, **kw)
