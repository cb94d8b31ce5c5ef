"""Model comparison -- host-side mirror of run_model_comparison_analysis (src/hmm_mcomp.jl:3-23, 55-88): `n_runs`
independent evidence estimates (SMC^2 or MBP-IBIS) for each model.

The (model, run) analyses are independent, so with a communicator they are dealt round-robin to the ranks (each analysis
runs on ONE GPU with its own seed) and the evidence estimates are all-gathered: no other communication."""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np

from .distributed import Comm
from .ibis import (C_ACCEPTANCE_ALPHA, C_ALG_NM_MBPI, C_ALG_NM_SMC2, C_DF_ESS_CRIT, C_DF_MBPI_ESS_CRIT, C_DF_MBPI_MUT, C_DF_MBPI_P,
                   C_DF_PF_P, C_DF_SMC2_P, run_pibis)
from .mbp_ibis import run_mbp_ibis
from .particle_filter import get_private_model
from .structs import DPOMPModel, HiddenMarkovModel


@dataclass
class ModelComparisonResults:
    """ModelComparisonResults (src/hmm_structs.jl:168-176)."""
    names: List[str]
    bme: np.ndarray       # (n_runs, n_models) evidence estimates (-log)
    mu: np.ndarray        # -log(mean(exp(-bme))) by model
    sigma: np.ndarray     # std of the estimates by model
    n_runs: int
    run_time: int
    theta_mu: list        # [run][model] -> posterior mean


def run_model_comparison(models: Sequence[HiddenMarkovModel], n_runs: int, fn_algorithm: Callable, comm: Optional[Comm] = None,
                         verbose: bool = True) -> ModelComparisonResults:
    """run_model_comparison_analysis(models::Array{HiddenMarkovModel,1}, n_runs, fn_algorithm) (src/hmm_mcomp.jl:3-23).
    `fn_algorithm(model, seed_index)` returns an ImportanceSample."""
    comm = comm or Comm(None)
    start_time = time.time_ns()
    n_models = len(models)
    d_max = max(len(m.prior.rand(1, np.random.default_rng(0))[:, 0]) for m in models)
    jobs = [(m, n) for m in range(n_models) for n in range(n_runs)]
    mine = [k for k in range(len(jobs)) if k % comm.world == comm.rank]
    loc = np.full((len(jobs), 1 + d_max), np.nan)
    for k in mine:
        m, n = jobs[k]
        if verbose:
            print(f" [rank {comm.rank}] model m{m + 1}: {models[m].model_name}, analysis {n + 1}")
        rs = fn_algorithm(models[m], k)
        loc[k, 0] = rs.bme[0]
        loc[k, 1:1 + len(rs.mu)] = rs.mu
    if comm.world > 1:  # every job was run by exactly one rank: gather the rows in job order
        order = [k for r in range(comm.world) for k in range(len(jobs)) if k % comm.world == r]
        mine_rows = loc[mine] if mine else np.zeros((0, 1 + d_max))
        counts = [len([k for k in range(len(jobs)) if k % comm.world == r]) for r in range(comm.world)]
        pad = max(counts)
        buf = np.full((pad, 1 + d_max), np.nan)
        buf[: len(mine_rows)] = mine_rows
        allr = comm.allgather_f64(buf.reshape(1, -1), comm.world).reshape(comm.world, pad, 1 + d_max)
        rows = np.concatenate([allr[r, : counts[r]] for r in range(comm.world)])
        loc = np.empty_like(loc)
        loc[order] = rows
    bme = np.zeros((n_runs, n_models))
    theta_mu = [[None] * n_models for _ in range(n_runs)]
    for k, (m, n) in enumerate(jobs):
        bme[n, m] = loc[k, 0]
        theta_mu[n][m] = loc[k, 1:][~np.isnan(loc[k, 1:])]
    mu = -np.log(np.mean(np.exp(-bme), axis=0))
    sigma = bme.std(axis=0, ddof=1) if n_runs > 1 else np.full(n_models, np.nan)
    out = ModelComparisonResults([m.model_name for m in models], bme, mu, sigma, n_runs, time.time_ns() - start_time, theta_mu)
    if verbose and comm.rank == 0:
        print(f"Analysis complete (total runtime := {round(out.run_time / 1e9)}s)")
    return out


def run_model_comparison_analysis(models: Sequence[DPOMPModel], y, n_runs: int = 3, algorithm: str = C_ALG_NM_SMC2,
                                  np_: Optional[int] = None, ess_rs_crit: Optional[float] = None, npf: int = C_DF_PF_P,
                                  n_props: int = C_DF_MBPI_MUT, seed: int = 1, comm: Optional[Comm] = None,
                                  verbose: bool = True) -> ModelComparisonResults:
    """run_model_comparison_analysis(models::Array{DPOMPModel,1}, y; n_runs = 3, algorithm = "SMC2", np, ess_rs_crit, npf,
    n_props) (src/hmm_mcomp.jl:55-88)."""
    smc2 = algorithm == C_ALG_NM_SMC2
    if not smc2 and not (algorithm[:4] == C_ALG_NM_MBPI or algorithm == "MIBIS"):
        print(f" WARNING - algorithm unknown: {algorithm}\n - defaulting to SMC2")
        smc2 = True
    outer_p = np_ if np_ is not None else (C_DF_SMC2_P if smc2 else C_DF_MBPI_P)
    crit = ess_rs_crit if ess_rs_crit is not None else (C_DF_ESS_CRIT if smc2 else C_DF_MBPI_ESS_CRIT)
    if verbose and (comm is None or comm.rank == 0):
        print(f"Running: {n_runs}-run {len(models)}-model Bayesian evidence analysis (algorithm := {algorithm})")

    def alg(mdl: HiddenMarkovModel, k: int):
        rng = np.random.default_rng([seed, k])
        theta_init = mdl.prior.rand(outer_p, rng)
        if smc2:  # alg_smc2 (:62-65)
            return run_pibis(mdl, theta_init, crit, True, C_ACCEPTANCE_ALPHA, npf, rng=rng, seed=seed * 1000003 + k, verbose=False)
        return run_mbp_ibis(mdl, theta_init, crit, n_props, False, C_ACCEPTANCE_ALPHA, rng=rng, seed=seed * 1000003 + k,
                            verbose=False)  # alg_mibis (:66-69)

    hmm = [get_private_model(m, y) for m in models]
    return run_model_comparison(hmm, n_runs, alg, comm=comm, verbose=verbose)
