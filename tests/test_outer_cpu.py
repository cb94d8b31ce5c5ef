"""CPU tests of the host-side outer layers (SMC^2, pMCMC) and their multi-rank sharding.  The CUDA particle filter is
replaced by an oracle-backed stand-in (tests/fake_pf.py) so that the bookkeeping, the migration plan and the
world_size-2 gloo path can be exercised without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT, load_case


def test_partition_and_migration_plan(dp):
    n, world = 11, 3
    bounds = [dp.partition_bounds(n, world, r) for r in range(world)]
    assert bounds == [(0, 4), (4, 8), (8, 11)]
    assert dp.partition_owner(n, world, np.arange(n)).tolist() == [0] * 4 + [1] * 4 + [2] * 3
    rng = np.random.default_rng(0)
    nidx0 = np.sort(rng.integers(0, n, n))
    # replay the plan on plain arrays: every destination must end up with its ancestor's payload
    payload = [np.arange(lo, hi) * 10 for lo, hi in bounds]
    plans = [dp.migration_plan(nidx0, n, world, r) for r in range(world)]
    sent = {}
    for r, (local_src, send_slots, send_counts, recv_slots, recv_counts) in enumerate(plans):
        pos = 0
        for dst in range(world):
            sent[(r, dst)] = payload[r][send_slots[pos:pos + send_counts[dst]]]
            pos += send_counts[dst]
    for r, (local_src, send_slots, send_counts, recv_slots, recv_counts) in enumerate(plans):
        new = payload[r][local_src].copy()
        pos = 0
        for src in range(world):
            got = sent[(src, r)]
            assert len(got) == recv_counts[src]
            new[recv_slots[pos:pos + len(got)]] = got
            pos += len(got)
        lo, hi = bounds[r]
        assert np.array_equal(new, nidx0[lo:hi] * 10)


def _pibis_worker(rank, world, port, out_path, kind):
    import sys
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import dpomp_b200 as dp
    from fake_pf import OraclePF
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    comm = None
    if world > 1:
        torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
        comm = dp.Comm()
    model = dp.generate_model("SIS", [100, 1])
    model.prior = dp.UniformProduct([0, 0], [0.01, 0.5])
    y = dp.get_observations(os.path.join(ROOT, "tests", "golden", "pooley.csv"))
    hmm = dp.get_private_model(model, y)
    desc = dp.compile_model(model, y)
    factory = lambda nb, sd: OraclePF(desc.desc, 40, nb, 1, sd)
    if kind == "pibis":
        theta0 = model.prior.rand(24, np.random.default_rng(3))
        res = dp.run_pibis(hmm, theta0, 0.5, True, 1.002, 40, rng=np.random.default_rng(9), seed=5, comm=comm,
                           pf_factory=factory, outer_rs=lambda w, rng: _host_rs_systematic(w, rng), verbose=False)
        out = dict(bme=res.bme, mu=res.mu, theta=res.theta, w=res.weight, k=res.k_log)
    elif kind == "mbp_mcmc":
        from fake_mbp import OracleMbp
        theta0 = model.prior.rand(5, np.random.default_rng(4)) * 0.5 + np.array([[0.002], [0.05]])
        mk = lambda n, sd: OracleMbp(desc.desc, [o.time for o in y], 0, n, 4096, sd)
        res = dp.run_mbp_mcmc(hmm, theta0, 120, 50, False, seed=8, comm=comm, particles_factory=mk, verbose=False)
        out = dict(theta=res.samples.theta, mu=res.samples.mu, acc=res.a_cnt)
    else:
        theta0 = model.prior.rand(5, np.random.default_rng(4)) * 0.5 + np.array([[0.002], [0.05]])
        res = dp.run_pmcmc(hmm, theta0, steps=40, adapt_period=20, p=40, seed=6, comm=comm, pf_factory=factory, verbose=False)
        out = dict(theta=res.samples.theta, mu=res.samples.mu, acc=res.accepted)
    if rank == 0:
        np.savez(out_path, **out)
    if world > 1:
        torch.distributed.destroy_process_group()


def _host_rs_systematic(w, rng):
    # CPU stand-in for the GPU search hook (tests only): the oracle's literal rs_systematic
    from oracle import oracle as orc
    return orc.rs(1, w, [rng.random()])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("kind", ["pibis", "pmcmc", "mbp_mcmc"])
def test_world_size_2_gloo_matches_single_process(tmp_path, kind):
    """Sharding theta-particles / chains over 2 ranks (gloo) gives bit-identical results to one process: the streams are
    keyed by global filter ids and call counters, the host RNG is replicated."""
    one = str(tmp_path / "one.npz"); two = str(tmp_path / "two.npz")
    _pibis_worker(0, 1, 0, one, kind)
    mp.spawn(_pibis_worker, args=(2, _free_port(), two, kind), nprocs=2, join=True)
    a, b = np.load(one), np.load(two)
    for k in a.files:
        assert np.array_equal(a[k], b[k]), k


def test_run_pibis_bookkeeping_against_oracle(dp, orc):
    """Host driver (batched sweep) vs the oracle's literal run_pibis on the same tiny problem: same law.  With so few
    particles only coarse agreement is expected; the GPU suite does the real z-tests."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_pf import OraclePF
    model, y, hmm, theta = load_case(dp, "sis_pooley")
    model.prior = dp.UniformProduct([0, 0], [0.01, 0.5])
    hmm = dp.get_private_model(model, y)
    cm = dp.compile_model(model, y)
    ours, ref = [], []
    for s in range(6):
        th0 = model.prior.rand(200, np.random.default_rng(s))
        r = dp.run_pibis(hmm, th0, 0.3, True, 1.002, 50, rng=np.random.default_rng(100 + s), seed=s,
                         pf_factory=lambda nb, sd: OraclePF(cm.desc, 50, nb, 1, sd), outer_rs=_host_rs_systematic, verbose=False)
        ours.append(r.bme[0])
        ref.append(orc.run_pibis(cm.desc, th0, model.prior.lower, model.prior.upper, npf=50, seed=200 + s, threads=orc.max_threads())["bme"][0])
    assert abs(np.mean(ours) - np.mean(ref)) < 0.5, (ours, ref)
    assert 19.0 < np.mean(ours) < 21.5


def test_prop_density_guard(dp):
    old = dp.ibis.ProposalDensity.identity(2) if hasattr(dp, "ibis") else None
    from dpomp_b200.ibis import ProposalDensity
    old = ProposalDensity.identity(2)
    assert dp.get_prop_density(np.array([[1.0, 2.0], [2.0, 1.0]]), old) is old  # not positive definite -> keep old
    new = dp.get_prop_density(np.array([[2.0, 0.5], [0.5, 1.0]]), old)
    assert np.allclose(new.chol @ new.chol.T, [[2.0, 0.5], [0.5, 1.0]])
    mu, cv = dp.compute_is_mu_covar(np.array([[1.0, 3.0], [2.0, 2.0]]), np.array([1.0, 3.0]))
    assert np.allclose(mu, [2.5, 2.0]) and np.allclose(cv, [[0.75, 0.0], [0.0, 0.0]])


def test_mbp_mcmc_host_driver_on_oracle_store(dp, orc):
    """run_mbp_mcmc (src/hmm_mcmc.jl:330-345, met_hastings_alg! :123-141) on the oracle-backed trajectory store: the
    posterior of SIS / pooley.csv against the anchors of SURVEY.md 8c (theta ~ (0.00327, 0.109)), acceptance bookkeeping
    and the shape contract of the result."""
    from fake_mbp import OracleMbp
    model = dp.generate_model("SIS", [100, 1])
    model.prior = dp.UniformProduct([0, 0], [0.01, 0.5])
    y = dp.get_observations(os.path.join(ROOT, "tests", "golden", "pooley.csv"))
    hmm = dp.get_private_model(model, y)
    cm = dp.compile_model(model, y)
    mk = lambda n, sd: OracleMbp(cm.desc, [o.time for o in y], 0, n, 4096, sd)
    # the configuration of the reference's own test (test/runtests.jl:40-44: run_mcmc_analysis defaults, 3 chains from the
    # prior, 50000 steps, adaptation 10000), whose seeded run gives samples.mu[1] = 0.003318
    th0 = model.prior.rand(3, np.random.default_rng(1))
    steps, adapt = 50000, 10000
    r = dp.run_mbp_mcmc(hmm, th0, steps, adapt, False, seed=1, particles_factory=mk, verbose=False)
    assert r.samples.theta.shape == (2, steps, 3) and r.adapt_period == adapt
    assert abs(r.samples.mu[0] - 0.003318) < 0.0003, r.samples.mu
    assert np.array_equal(r.samples.theta[:, 0, :], th0)  # theta[:,1,mc] .= xi.theta
    assert np.all(r.a_cnt[:, 0] >= 1) and np.all(r.a_cnt.sum(axis=1) <= steps)
    # accepted moves change theta, rejected ones repeat it
    moved = np.any(np.diff(r.samples.theta, axis=1) != 0, axis=0).sum(axis=0) + 1
    assert np.array_equal(moved, r.a_cnt.sum(axis=1))
    assert abs(r.samples.mu[0] - 0.00327) < 0.0003 and abs(r.samples.mu[1] - 0.109) < 0.015, r.samples.mu
    assert np.all(r.sre[:, 1] < 1.1) and np.all(np.abs(r.a_cnt[:, 1] / (steps - adapt) - 0.333) < 0.03)  # 1.002 / 0.999 scaling
    # finite adaptation freezes the scale after the adaptation period; an impossible prior rejects every proposal
    model.prior = dp.UniformProduct([0, 0], [1e-9, 1e-9])
    hmm2 = dp.get_private_model(model, y)
    r2 = dp.run_mbp_mcmc(hmm2, th0[:, :2], 30, 10, True, seed=2, particles_factory=mk, verbose=False)
    assert np.all(r2.samples.theta == th0[:, :2][:, None, :]) and np.array_equal(r2.a_cnt, [[1, 0], [1, 0]])


def test_smc2_hastings_correction_removes_the_evidence_bias(dp, orc):
    """run_pibis(...; hastings_correction = True) -- an option beyond the reference: with the proposal-density ratio in
    the acceptance step of the independent proposals, the evidence of the exactly solvable pure-death case
    (tests/test_oracle.py::test_smc2_evidence_against_exact_quadrature, -ln p(y) = 13.147) is recovered; the default
    (reference behaviour) stays ~0.16 lower.  Host driver on the oracle-backed filter bank."""
    from conftest import death_rate_case
    from fake_pf import OraclePF
    case = death_rate_case(dp, 801)
    model, hmm, cm, bme_exact = case["model"], case["hmm"], case["cm"], case["bme"]
    factory = lambda nb, sd: OraclePF(cm.desc, 200, nb, 1, sd)
    res = {}
    for corr in (True, False):
        b = []
        for s in range(4):
            rng = np.random.default_rng(500 + s)
            th0 = model.prior.rand(1500, rng)
            r = dp.run_pibis(hmm, th0, 0.3, True, 1.002, 200, rng=rng, seed=600 + s, pf_factory=factory,
                             outer_rs=lambda w, rng: _host_rs_systematic(w, rng), verbose=False, hastings_correction=corr)
            b.append(r.bme[0])
        res[corr] = -np.log(np.mean(np.exp(-np.array(b))))
    assert abs(res[True] - bme_exact) < 0.07, (res, bme_exact)
    assert res[False] < res[True] - 0.05, (res, bme_exact)


def test_mbp_mcmc_posterior_against_exact_quadrature(dp, orc):
    """run_mbp_mcmc on the exactly solvable pure-death case: posterior mean and standard deviation of the death rate against
    quadrature of the forward-algorithm likelihood (prior U(0, 0.2)).  Host driver on the oracle-backed trajectory store."""
    from conftest import death_rate_case
    from fake_mbp import OracleMbp
    case = death_rate_case(dp, 801)
    model, y, hmm, cm, mean, sd = case["model"], case["y"], case["hmm"], case["cm"], case["mean"], case["sd"]
    mk = lambda n, sd_: OracleMbp(cm.desc, [o.time for o in y], 0, n, 4096, sd_)
    r = dp.run_mbp_mcmc(hmm, model.prior.rand(4, np.random.default_rng(1)), 20000, 4000, False, seed=3, particles_factory=mk,
                        verbose=False)
    assert abs(r.samples.mu[0] - mean) < 0.03 * mean, (r.samples.mu, mean)
    assert abs(np.sqrt(r.samples.cv[0, 0]) - sd) < 0.12 * sd, (r.samples.cv, sd)
    assert r.sre[0, 1] < 1.05


def test_pmcmc_posterior_against_exact_quadrature(dp, orc):
    """run_pmcmc (src/hmm_mcmc.jl:349-365, 166-211) on the exactly solvable pure-death case: the pseudo-marginal chain
    targets the exact posterior whatever the particle count; mean and sd against quadrature.  Host driver on the
    oracle-backed particle filter."""
    from conftest import death_rate_case
    from fake_pf import OraclePF
    case = death_rate_case(dp, 801)
    factory = lambda nb, sd: OraclePF(case["cm"].desc, 300, nb, 1, sd)
    th0 = np.array([[0.03, 0.05, 0.07, 0.09]])
    r = dp.run_pmcmc(case["hmm"], th0, steps=6000, adapt_period=1500, p=300, seed=4, pf_factory=factory, verbose=False)
    assert r.samples.theta.shape == (1, 6000, 4)
    assert abs(r.samples.mu[0] - case["mean"]) < 0.04 * case["mean"], (r.samples.mu, case["mean"])
    assert abs(np.sqrt(r.samples.cv[0, 0]) - case["sd"]) < 0.15 * case["sd"], (r.samples.cv, case["sd"])
    assert np.all(r.accepted > 300)


def test_pmcmc_adapt_interval_is_float_like_the_reference(dp, orc):
    """ADAPT_INTERVAL = adapt_period / 10 is a Float64 (src/hmm_mcmc.jl:168,200): with adapt_period = 25 no
    step index is a multiple of 2.5 except 5, 10, 15, 20; with adapt_period = 27 the proposal covariance never adapts."""
    from conftest import death_rate_case
    from fake_pf import OraclePF
    import dpomp_b200 as dpm
    import math
    case = death_rate_case(dp, 201)
    factory = lambda nb, sd: OraclePF(case["cm"].desc, 64, nb, 1, sd)
    th0 = np.array([[0.05, 0.07]])
    calls = []
    real = np.linalg.cholesky

    def spy(a):
        calls.append(1)
        return real(a)
    np.linalg.cholesky = spy
    try:
        dp.run_pmcmc(case["hmm"], th0, steps=40, adapt_period=27, p=64, seed=4, pf_factory=factory, verbose=False)
        n27 = len(calls)
        calls.clear()
        dp.run_pmcmc(case["hmm"], th0, steps=40, adapt_period=25, p=64, seed=4, pf_factory=factory, verbose=False)
        n25 = len(calls)
    finally:
        np.linalg.cholesky = real
    assert n27 == 0
    expected = sum(1 for i in range(2, 25) if math.fmod(i, 2.5) == 0)  # Julia steps 2..24 with i % 2.5 == 0: 5, 10, 15, 20
    assert expected == 4 and n25 <= expected * th0.shape[1] and n25 > 0
