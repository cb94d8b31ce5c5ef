"""CPU tests of the caller layers that sit on the particle-filter closure: ARQ-MCMC (src/arq_alg_std.jl, src/arq_alg_cmn.jl,
src/arq_main.jl) and the model-comparison fan-out (src/hmm_mcomp.jl).  No GPU: the density is an analytic function."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gauss_pdf(mu, sd):
    mu, sd = np.asarray(mu, dtype=float), np.asarray(sd, dtype=float)

    def pdf(theta):
        pdf.calls += 1
        return float(-0.5 * np.sum(((np.asarray(theta) - mu) / sd) ** 2))
    pdf.calls = 0
    return pdf


def test_get_grid_points_cache_semantics(dp):
    """get_grid_point! (src/arq_alg_std.jl:4-41): cache hits, the burn-in sample limit, the running mean of repeated
    evaluations, prior rejection, and same-step requests served in order."""
    rng = np.random.default_rng(0)
    pdf = _gauss_pdf([0.0, 0.0], [1.0, 1.0])
    prior = lambda th: -np.inf if np.any(th < 0) else 0.0
    mdl = dp.LikelihoodModel(pdf, np.array([0.5, 0.5]), np.array([0.25, 0.25]), 3, 50, 0.0, prior)
    grid = {}
    a, b, c = dp.get_grid_points(grid, [np.array([1, 2]), np.array([1, 2]), np.array([-3, 0])], mdl, [False] * 3, rng)
    assert np.allclose(a.result.sample, [0.75, 1.25])  # offset + index * interval (src/arq_alg_cmn.jl:24-32)
    assert a.process_run and a.result.visited == 1 and a.result.sampled == 1
    assert b.process_run and b.result.visited == 2 and b.result.sampled == 2  # second request of the step: after the first
    assert np.isclose(b.result.log_likelihood, a.result.log_likelihood)       # mean of two equal evaluations
    assert c.prior == -np.inf and not c.process_run and c.result.log_likelihood == -np.inf and (-3, 0) not in grid
    assert pdf.calls == 2
    d, = dp.get_grid_points(grid, [np.array([1, 2])], mdl, [True], rng)  # burn-in: limit 1, already visited -> cached
    assert not d.process_run and d.result.visited == 2 and d.result.sampled == 2 and pdf.calls == 2
    e, = dp.get_grid_points(grid, [np.array([1, 2])], mdl, [False], rng)
    f, = dp.get_grid_points(grid, [np.array([1, 2])], mdl, [False], rng)
    assert e.process_run and e.result.visited == 3 and not f.process_run and f.result.sampled == 4 and pdf.calls == 3


def test_get_grid_points_batched_density(dp):
    """A density closure that accepts (n_theta, B) matrices (like get_log_pdf_fn's, attribute `batched`) receives the
    evaluations of a round in chunks of at most n_batch columns, and yields the same cache as the scalar closure."""
    rng = np.random.default_rng(0)
    mu, sd = np.array([0.4, 0.9]), np.array([0.2, 0.3])
    calls = []

    def pdf(theta):
        th = np.asarray(theta, dtype=float)
        calls.append(th.shape)
        z = (th - (mu[:, None] if th.ndim == 2 else mu)) / (sd[:, None] if th.ndim == 2 else sd)
        return -0.5 * np.sum(z * z, axis=0)
    pdf.batched, pdf.n_batch = True, 2
    prior = lambda th: 0.0
    mdl = dp.LikelihoodModel(pdf, np.array([0.1, 0.1]), np.array([0.05, 0.05]), 1, 50, 0.0, prior)
    grid = {}
    pts = [np.array([i, 2 * i]) for i in range(5)] + [np.array([1, 2])]  # five distinct points and one repeat
    res = dp.get_grid_points(grid, pts, mdl, [False] * 6, rng)
    assert calls == [(2, 2), (2, 2), (2, 1)] and len(grid) == 5  # 5 evaluations in chunks of n_batch = 2; the repeat is cached
    assert [r.process_run for r in res] == [True] * 5 + [False]
    for r, p in zip(res, pts):
        val = mdl.sample_offset + p * mdl.sample_interval
        assert np.allclose(r.result.sample, val) and np.isclose(r.result.log_likelihood, float(pdf(val)))
    assert res[5].result.sampled == 2 and res[1].result.sampled == 1


def test_theta_f_and_adapt_jw(dp):
    rng = np.random.default_rng(1)
    th = np.array([10, 20, 30])
    for _ in range(200):  # get_theta_f (src/arq_alg_cmn.jl:36-45): L1 distance between 1 and j
        f = dp.get_theta_f(th, np.array([1.0, 2.0, 0.5]), 4, 1, rng)
        assert 1 <= np.abs(f - th).sum() <= 4
    assert np.abs(dp.get_theta_f(th, np.ones(3), 3, 3, rng) - th).sum() == 3
    # adapt_jw! (src/arq_alg_cmn.jl:60-86)
    acc = np.zeros(400, dtype=bool); acc[0] = True; acc[100:200:2] = True
    idx = np.zeros((2, 400), dtype=np.int64); idx[0, :200] = np.arange(200) % 7
    jw = np.ones(2)
    j = dp.adapt_jw(jw, 20, 10, acc, 100, 200, 0.33, idx)
    assert j == max(int(round(10 * (0.5 / 0.33))), 2)
    assert jw[0] > 0 and jw[1] == jw[0]  # zero spread -> minimum positive spread
    assert dp.adapt_jw(jw, 20, 2, acc, 100, 400, 0.33, idx) == 20  # low acceptance at the minimum jump: lar_j
    acc1 = np.zeros(400, dtype=bool); acc1[0] = True
    assert dp.adapt_jw(jw, 20, 2, acc1, 100, 300, 0.33, idx) == 30  # only the initial sample accepted: contingency jump


def test_arq_mcmc_on_analytic_density(dp):
    """run_arq_mcmc_analysis(model::ARQModel, priors) (src/arq_main.jl:96-110): both the importance sample over the grid
    cache and the rejection samples recover a Gaussian target; result struct contract."""
    pdf = _gauss_pdf([0.31, 1.22], [0.05, 0.2])
    arq = dp.ARQModel(pdf, np.array([0.01, 0.04]), np.array([0.005, 0.02]))
    prior = lambda th: -np.inf if np.any(th < 0) else 0.0
    r = dp.run_arq_mcmc_analysis(arq, [prior], steps=6000, n_chains=3, seed=4, verbose=False)
    assert isinstance(r, dp.ARQMCMCSample) and r.samples.theta.shape == (2, 6000, 3) and r.adapt_period == 1200
    assert np.allclose(r.imp_sample.mu, [0.31, 1.22], atol=0.02) and np.allclose(r.samples.mu, [0.31, 1.22], atol=0.03)
    assert np.allclose(np.sqrt(np.diag(r.samples.cv)), [0.05, 0.2], rtol=0.3)
    assert len(r.sample_cache) == r.imp_sample.theta.shape[1] == len(r.imp_sample.weight)
    assert r.fx.sum() == pdf.calls == sum(p.visited for p in r.sample_cache.values())  # sample_limit 1: one call per point
    assert np.all(r.sre[:, 1] < 1.2) and np.all(r.acceptance[:, 1] > 0.1)
    # every sample lies on the grid
    k = (r.samples.theta[:, -1, :] - arq.sample_offset[:, None]) / arq.sample_interval[:, None]
    assert np.allclose(k, np.round(k), atol=1e-9)
    # a retained cache is reused: no new evaluations at sample_limit 1 inside the explored region
    calls = pdf.calls
    r2 = dp.run_arq_mcmc_analysis(arq, [prior], steps=600, n_chains=2, seed=5, sample_cache=r.sample_cache, verbose=False)
    assert pdf.calls - calls < 0.2 * 600 * 2 and r2.sample_cache is r.sample_cache


def _mcomp_worker(rank, world, port, out_path):
    import sys
    sys.path.insert(0, ROOT)
    import dpomp_b200 as dp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    comm = None
    if world > 1:
        torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
        comm = dp.Comm()
    m1 = dp.get_private_model(dp.generate_model("SIS", [100, 1]), [])
    m2 = dp.get_private_model(dp.generate_model("SEIR", [100, 0, 1, 0]), [])

    def alg(mdl, k):  # stand-in analysis: deterministic in (model, job index)
        d = 2 if mdl.model_name == "SIS" else 3
        return dp.ImportanceSample(np.arange(d) + 0.5 * k, np.eye(d), np.zeros((d, 1)), np.ones(1), 0, np.array([20.0 + k, 0.0]))
    res = dp.run_model_comparison([m1, m2], 3, alg, comm=comm, verbose=False)
    if rank == 0:
        np.savez(out_path, bme=res.bme, mu=res.mu, sigma=res.sigma, t00=res.theta_mu[0][0], t21=res.theta_mu[2][1])
    if world > 1:
        torch.distributed.destroy_process_group()


def test_model_comparison_fanout_and_two_ranks(tmp_path):
    """run_model_comparison_analysis (src/hmm_mcomp.jl:3-23): bookkeeping of the (run, model) evidence matrix; dealing the
    independent analyses to 2 ranks (gloo) gives the same result."""
    one, two = str(tmp_path / "one.npz"), str(tmp_path / "two.npz")
    _mcomp_worker(0, 1, 0, one)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_mcomp_worker, args=(2, port, two), nprocs=2, join=True)
    a, b = np.load(one), np.load(two)
    for k in a.files:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["bme"], [[20, 23], [21, 24], [22, 25]])
    assert np.allclose(a["mu"], -np.log(np.mean(np.exp(-a["bme"]), axis=0))) and np.allclose(a["sigma"], 1.0)
    assert np.array_equal(a["t00"], [0.0, 1.0]) and np.array_equal(a["t21"], [2.5, 3.5, 4.5])


def test_save_to_file_writes_the_reference_csv_formats(dp, tmp_path):
    """save_to_file(results::SimResults, dpath) (src/hmm_utils.jl:35-72): sim.csv + obs.csv; Julia float formatting."""
    from dpomp_b200.sim import _jl_float
    assert [_jl_float(v) for v in (20.0, 1e-5, 0.0001, 1234567.8, 100000.0, 0.30000000000000004)] == \
        ["20.0", "1.0e-5", "0.0001", "1.2345678e6", "100000.0", "0.30000000000000004"]
    part = dp.Particle(np.array([0.003, 0.1]), np.array([100, 1]), np.array([99, 1]),
                       [dp.Event(1.25, 1), dp.Event(2.5, 2)], 0.0, np.zeros(2))
    res = dp.SimResults("SIS", part, [np.array([99, 2]), np.array([100, 1])],
                        [dp.Observation(20.0, 1, 1.0, np.array([100, 1])), dp.Observation(40.0, 1, 1.0, np.array([90, 11]))])
    d = str(tmp_path) + os.sep
    dp.save_to_file(res, d)
    assert open(d + "sim.csv").read() == "time, event,1,2\n1.25,1,99,2\n2.5,2,100,1"
    assert open(d + "obs.csv").read() == "time,id,1,2\n20.0,1,100,1\n40.0,1,90,11"
    dp.save_to_file(res, d, literal=True)  # the reference's literal bytes: the state printed as one Julia array
    assert open(d + "sim.csv").read() == "time, event,1\n1.25,1,[99, 2]\n2.5,2,[100, 1]"
