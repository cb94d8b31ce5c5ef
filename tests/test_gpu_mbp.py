"""GPU tests of the MBP-IBIS layer (row a12): iterate_particle!, partial_model_based_proposal and run_mbp_ibis through
the C ABI, against the oracle's literal restatement on identical Philox draws."""
import numpy as np
import pytest

from conftest import load_case

pytestmark = pytest.mark.gpu


def _setup(dp, case="sis_pooley", t0=False):
    model, y, hmm, theta = load_case(dp, case)
    if t0:
        def rf(out, p, x):
            out[0] = p[0] * x[0] * x[1]; out[1] = p[1] * x[1]
        model = dp.generate_custom_model("SIS", rf, [100, 1], [[-1, 1], [1, -1]], prior=dp.generate_weak_prior(3), t0_index=3)
        theta = np.array([0.003, 0.1, 4.0])
    hmm = dp.get_private_model(model, y)
    return model, y, hmm, theta


def _setup_any(dp, case, t0):
    """_setup plus "sis_freq": the frequency-dependent SIS model (src/hmm_examples.jl:126-131), whose rate table has
    denominators and therefore runs the generic (run-time table) trajectory kernels instead of a predefined-model instantiation."""
    if case != "sis_freq":
        return _setup(dp, case, t0)
    model, y, hmm, theta = load_case(dp, "sis_pooley")
    model = dp.generate_model("SIS", [100, 1], freq_dep=True)
    return model, y, dp.get_private_model(model, y), np.array([0.3, 0.1])


@pytest.mark.parametrize("case,t0", [("sis_pooley", False), ("sis_pooley", True), ("seir_c3", False), ("sis_freq", False)])
def test_iterate_and_propose_match_oracle(dp, orc, case, t0):
    model, y, hmm, theta = _setup_any(dp, case, t0)
    dm = dp.device_model(hmm)
    desc = dm.compiled.desc
    n, cap = 64, 4096
    rng = np.random.default_rng(3)
    thetas = theta[:, None] * rng.uniform(0.8, 1.25, size=(len(theta), n))
    if t0:
        thetas[2] = rng.uniform(1.0, 8.0, n)
    pt = dp.MbpParticles(dm, n, cap, seed=5)
    n_obs = min(len(y), 4)
    # oracle state per particle
    o_fc = np.tile(np.asarray(model.initial_condition, dtype=np.int64), (n, 1))
    o_t = np.zeros((n, cap)); o_y = np.zeros((n, cap), dtype=np.int32); o_len = np.zeros(n, dtype=np.int64); o_ll = np.zeros((n, 2))
    for obs_i in range(1, n_obs + 1):
        key = 9000 + obs_i
        pt.set_stream_key(key)
        lg = pt.iterate(thetas, obs_i, fresh=(obs_i == 1))
        for p in range(n):
            t_start = (thetas[2, p] if t0 else 0.0) if obs_i == 1 else y[obs_i - 2].time
            g, o_len[p] = orc.mbp_iterate(desc, thetas[:, p], o_fc[p], o_t[p], o_y[p], o_len[p], o_ll[p], t_start, obs_i, key, p)
            assert np.isclose(lg[p], g, rtol=1e-12, atol=0)
    for p in (0, 7, n - 1):
        fc, times, types, ll = pt.get_particle(p + 1)
        assert np.array_equal(fc, o_fc[p]) and len(times) == o_len[p]
        assert np.array_equal(types, o_y[p, : o_len[p]]) and np.allclose(times, o_t[p, : o_len[p]], rtol=1e-12, atol=0)
        assert np.isclose(ll[0], o_ll[p, 0], rtol=1e-12)
    # model-based proposals conditional on those trajectories
    theta_f = thetas * rng.uniform(0.85, 1.2, size=thetas.shape)
    valid = np.ones(n, dtype=bool); valid[3] = False
    key = 777
    pt.set_stream_key(key)
    ll_f = pt.propose(thetas, theta_f, valid, n_obs)
    assert np.all(np.isneginf(ll_f[3]))
    for p in (0, 1, 7, 20, n - 1):
        times, types, fc, ll, rc = orc.mbp_propose(desc, thetas[:, p], theta_f[:, p], o_t[p], o_y[p], o_len[p], cap, n_obs, key, p)
        g_fc, g_times, g_types, g_ll = pt.get_particle(p + 1, proposal=True)
        assert rc == 0 and np.array_equal(g_fc, fc) and np.array_equal(g_types, types)
        assert np.allclose(g_times, times, rtol=1e-12, atol=0) and np.allclose(g_ll, ll, rtol=1e-12) and np.allclose(ll_f[p], ll, rtol=1e-12)
    # accept / permute move whole particles
    pt.accept([2, 8])
    a = pt.get_particle(2); b = pt.get_particle(2, proposal=True)
    assert all(np.array_equal(u, v) for u, v in zip(a, b))
    before = [pt.get_particle(p + 1) for p in range(n)]
    nidx = np.sort(rng.integers(1, n + 1, n))
    pt.permute(nidx)
    for p in (0, 5, n - 1):
        got = pt.get_particle(p + 1)
        assert all(np.array_equal(u, v) for u, v in zip(got, before[nidx[p] - 1]))


def test_trajectory_overflow_gives_minus_inf(dp):
    model, y, hmm, theta = _setup(dp)
    pt = dp.MbpParticles(dp.device_model(hmm), 32, 16, seed=1)  # 16 events is far too few for 20 time units
    lg = pt.iterate(np.tile(theta[:, None], (1, 32)), 1, fresh=True)
    dead = np.isneginf(lg)
    assert dead.mean() > 0.4  # the rest are early extinctions (the single infective recovers) with <= 16 events
    for p in range(32):
        fc, times, types, ll = pt.get_particle(p + 1)
        assert len(times) <= 16 and np.isneginf(ll[0]) == dead[p]
        if not dead[p]:
            assert fc[1] == 0  # extinct


def test_run_mbp_ibis_against_oracle_and_anchor(dp, orc):
    """run_ibis_analysis(model, y; algorithm = "MBPI") defaults (10000 particles, 3 mutations, ess 0.5) on SIS/pooley."""
    model, y, hmm, theta = _setup(dp)
    model.prior = dp.UniformProduct([0, 0], [0.01, 0.5])
    cm = dp.compile_model(model, y)
    reps = 4
    ours, ref, mus, rmus = [], [], [], []
    for s in range(reps):
        r = dp.run_ibis_analysis(model, y, algorithm="MBPI", seed=300 + s, verbose=False)
        ours.append(r.bme.copy()); mus.append(r.mu.copy())
        th0 = model.prior.rand(10000, np.random.default_rng(400 + s))
        o = orc.run_mbp_ibis(cm.desc, th0, model.prior.lower, model.prior.upper, seed=500 + s, threads=orc.max_threads(), cap=8192)
        ref.append(o["bme"].copy()); rmus.append(o["mu"].copy())
    ours, ref, mus, rmus = map(np.array, (ours, ref, mus, rmus))
    z = (ours[:, 0].mean() - ref[:, 0].mean()) / np.sqrt(ours[:, 0].var(ddof=1) / reps + ref[:, 0].var(ddof=1) / reps)
    assert abs(z) < 4.5, (ours, ref)
    assert 19.8 < ours[:, 0].mean() < 20.6  # prior-IS anchor 20.18 +- 0.1, SMC^2 reference run 19.98
    assert np.all(np.abs(mus.mean(axis=0) - rmus.mean(axis=0)) < 0.12 * np.abs(rmus.mean(axis=0)))
    # posterior mean: theta ~ (0.00327, 0.109)
    assert abs(mus[:, 0].mean() - 0.00327) < 0.0005 and abs(mus[:, 1].mean() - 0.109) < 0.02


def test_mbp_ibis_seir_stratified_against_oracle(dp, orc):
    """BASELINE config C5 at reduced scale: SEIR [100,0,1,0] on the first 30 observations of seir_c3.csv, prior
    U(0,(0.02,1,0.5)), 2048 theta-particles, n_props = 3, ind_prop = false, ess 0.5, STRATIFIED outer resampling
    (src/hmm_resample.jl:66-83; the reference hard-codes systematic at src/hmm_ibis.jl:194 -- stated departure).  Evidence
    and posterior mean are z-tested over replicates against the oracle's literal run_mbp_ibis with rs_type = 2."""
    model, y, hmm, theta = load_case(dp, "seir_c3")
    model.prior = dp.UniformProduct([0, 0, 0], [0.02, 1.0, 0.5])
    y = y[:30]
    hmm = dp.get_private_model(model, y)
    cm = dp.compile_model(model, y)
    reps, n_o = 5, 2048
    ours, ref, mus, rmus = [], [], [], []
    for s in range(reps):
        th0 = model.prior.rand(n_o, np.random.default_rng(600 + s))
        r = dp.run_mbp_ibis(hmm, th0, 0.5, 3, False, 1.002, seed=620 + s, outer_rs=dp.rs_stratified, verbose=False)
        assert r.k_log[1] > 0
        ours.append(r.bme.copy()); mus.append(r.mu.copy())
        th1 = model.prior.rand(n_o, np.random.default_rng(640 + s))
        o = orc.run_mbp_ibis(cm.desc, th1, model.prior.lower, model.prior.upper, ess_rs_crit=0.5, n_props=3, ind_prop=False,
                             rs_type=2, seed=660 + s, threads=orc.max_threads(), cap=8192)
        ref.append(o["bme"].copy()); rmus.append(o["mu"].copy())
    ours, ref, mus, rmus = map(np.array, (ours, ref, mus, rmus))

    def z(a, b):
        return (a.mean() - b.mean()) / np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b) + 1e-300)
    for k in range(2):
        assert abs(z(ours[:, k], ref[:, k])) < 4.5, (k, ours[:, k], ref[:, k])
    for j in range(3):
        assert abs(z(mus[:, j], rmus[:, j])) < 4.5, (j, mus[:, j], rmus[:, j])
    assert np.all(mus > 0) and np.all(mus < [0.02, 1.0, 0.5])


def test_export_import_roundtrip(dp):
    model, y, hmm, theta = _setup(dp)
    dm = dp.device_model(hmm)
    a = dp.MbpParticles(dm, 40, 4096, seed=2)
    b = dp.MbpParticles(dm, 40, 4096, seed=9)
    th = np.tile(theta[:, None], (1, 40)) * np.linspace(0.8, 1.3, 40)[None, :]
    a.iterate(th, 1, True); a.iterate(th, 2, False)
    slots = np.array([3, 17, 40, 1])
    lens = a.lengths(slots)
    fixed, times, types = a.export_particles(slots, lens)
    assert times.numel() == lens.sum() and fixed.numel() == 4 * 16
    dst = np.array([5, 6, 7, 8])
    b.import_particles(dst, lens, fixed, times, types)
    for s, d in zip(slots, dst):
        assert all(np.array_equal(u, v) for u, v in zip(a.get_particle(int(s)), b.get_particle(int(d))))


def _mbp_worker(rank, world, port, out_path):
    import os, sys
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    import torch
    import dpomp_b200 as dp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    comm = None
    if world > 1:
        torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
        comm = dp.Comm()
    torch.cuda.set_device(0)
    model = dp.generate_model("SIS", [100, 1])
    model.prior = dp.UniformProduct([0, 0], [0.01, 0.5])
    y = dp.get_observations(os.path.join(ROOT, "tests", "golden", "pooley.csv"))
    hmm = dp.get_private_model(model, y)
    th0 = model.prior.rand(501, np.random.default_rng(5))
    res = dp.run_mbp_ibis(hmm, th0, 0.5, 3, False, 1.002, rng=np.random.default_rng(6), seed=7, comm=comm, max_traj=4096, verbose=False,
                          device_outer=False)  # gloo exchanges = the host-driven loop on both sides; the device-resident loop over NCCL: test_gpu_multi.py
    if rank == 0:
        np.savez(out_path, bme=res.bme, mu=res.mu, theta=res.theta, w=res.weight)
    if world > 1:
        torch.distributed.destroy_process_group()


def test_mbp_ibis_two_ranks_equal_one_rank_bitwise(tmp_path):
    """theta-particles and their trajectories sharded over 2 ranks (gloo collectives, both on cuda:0): the two-phase
    trajectory migration must reproduce the single-process run bit for bit."""
    import socket
    import torch.multiprocessing as mp
    one, two = str(tmp_path / "one.npz"), str(tmp_path / "two.npz")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_mbp_worker, args=(1, port, one), nprocs=1, join=True)
    mp.spawn(_mbp_worker, args=(2, port, two), nprocs=2, join=True)
    a, b = np.load(one), np.load(two)
    for k in a.files:
        assert np.array_equal(a[k], b[k]), k


def test_gillespie_sim_entry_point(dp, orc):
    """gillespie_sim(model, theta; tmax, num_obs, n_sims) (src/DiscretePOMP.jl:134-152): result structs, consistency of the
    recorded trajectory with the observed states, and the law of the final state against the oracle's simulator."""
    model = dp.generate_model("SIS", [100, 1])
    theta = np.array([0.003, 0.1])
    x = dp.gillespie_sim(model, theta, verbose=False)  # reference test: test/runtests.jl:22-26
    assert isinstance(x, dp.SimResults) and len(x.observations) == 5 and [o.time for o in x.observations] == [20.0, 40.0, 60.0, 80.0, 100.0]
    assert len(x.population) == len(x.particle.trajectory) and np.all(np.diff([e.time for e in x.particle.trajectory]) >= 0)
    if x.population:
        assert np.array_equal(x.population[-1], x.particle.final_condition) and x.population[-1].sum() == 101
    assert np.array_equal(x.observations[-1].val, x.particle.final_condition)  # dmy_obs_fn: y.val .= population
    # state at each observation time = initial condition + transitions of the events up to that time
    tm = model.m_transition
    for o in x.observations:
        k = sum(e.time <= o.time for e in x.particle.trajectory)
        st = model.initial_condition + sum((tm[e.event_type - 1] for e in x.particle.trajectory[:k]), np.zeros(2, dtype=np.int64))
        assert np.array_equal(o.val, st)
    sims = dp.gillespie_sim(model, theta, n_sims=2000, seed=5, verbose=False)
    gpu_final = np.array([s.particle.final_condition[1] for s in sims], dtype=float)
    y = [dp.Observation(20.0 * (i + 1), 1, 1.0, [0, 0]) for i in range(5)]
    cm = dp.compile_model(model, y)
    ref_final = np.array([orc.gillespie_sim(cm.desc, theta, key=900 + i)[0][-1, 1] for i in range(2000)], dtype=float)
    # mixture of early extinctions (I = 0) and the endemic state (I ~ 67): compare extinction rate and endemic mean
    assert abs((gpu_final == 0).mean() - (ref_final == 0).mean()) < 0.05
    assert abs(gpu_final[gpu_final > 0].mean() - ref_final[ref_final > 0].mean()) < 1.0


def test_mbp_mcmc_chains_match_oracle_store(dp, orc):
    """run_mbp_mcmc (src/hmm_mcmc.jl:330-345): the same host driver on the CUDA trajectory store and on the oracle-backed
    store, identical host streams and Philox keys -- the chains must agree step for step (proposal log-likelihoods agree
    to ~1e-12, so accept/reject decisions are identical away from knife edges)."""
    import os, sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_mbp import OracleMbp
    for case, prior_hi, t0 in (("sis_pooley", [0.01, 0.5], False), ("seir_c3", [0.02, 1.0, 0.5], False), ("sis_pooley", [0.01, 0.5, 10.0], True)):
        model, y, hmm, theta = _setup(dp, case, t0)
        y = y[:20]
        model.prior = dp.UniformProduct([0.0] * len(prior_hi), prior_hi)
        hmm = dp.get_private_model(model, y)
        cm = dp.compile_model(model, y)
        mk = lambda n, sd: OracleMbp(cm.desc, [o.time for o in y], hmm.t0_index, n, 4096, sd)
        th0 = theta[:, None] * np.random.default_rng(2).uniform(0.8, 1.2, size=(len(theta), 4))
        a = dp.run_mbp_mcmc(hmm, th0, 300, 100, False, seed=31, max_traj=4096, verbose=False)
        b = dp.run_mbp_mcmc(hmm, th0, 300, 100, False, seed=31, particles_factory=mk, verbose=False)
        assert np.allclose(a.samples.theta, b.samples.theta, rtol=1e-9, atol=0), case
        assert np.array_equal(a.a_cnt, b.a_cnt) and a.a_cnt[:, 1].sum() > 0


def test_run_mcmc_analysis_default_algorithm_posterior(dp):
    """run_mcmc_analysis(model, y) -- MBP-MCMC is the reference's default (src/DiscretePOMP.jl:185-193): posterior of
    SIS / pooley.csv against the anchors (theta ~ (0.00327, 0.109); the reference's seeded run of
    test/runtests.jl:40-44 gives samples.mu[1] = 0.003318) with 16 lock-step chains started around the mode."""
    model, y, hmm, theta = _setup(dp)
    model.prior = dp.UniformProduct([0, 0], [0.01, 0.5])
    th0 = theta[:, None] * np.random.default_rng(8).uniform(0.7, 1.3, size=(2, 16))
    r = dp.run_mcmc_analysis(model, y, n_chains=16, initial_parameters=th0, steps=3000, adapt_period=600, seed=5, verbose=False)
    assert r.samples.theta.shape == (2, 3000, 16) and np.array_equal(r.samples.theta[:, 0, :], th0)
    # short chains (one slow-mixing mode sits near 0.0028, see the 50000-step CPU test for the tight anchor)
    assert abs(r.samples.mu[0] - 0.00327) < 0.0006 and abs(r.samples.mu[1] - 0.109) < 0.03, r.samples.mu
    assert np.all(r.a_cnt[:, 1] > 0.05 * 2400)
    # chains drawn from the prior by default, like the reference
    r2 = dp.run_mcmc_analysis(model, y, n_chains=3, steps=50, seed=6, verbose=False)
    assert r2.samples.theta.shape == (2, 50, 3) and r2.adapt_period == 10
    with pytest.raises(NotImplementedError):
        dp.run_mcmc_analysis(model, y, mbp=False)


@pytest.mark.parametrize("case,t0", [("sis_pooley", False), ("sis_pooley", True), ("seir_c3", False), ("sis_freq", False)])
def test_warp_and_thread_per_trajectory_kernels_are_identical(dp, case, t0):
    """dpomp_mbp_set_mode: the warp-per-trajectory kernels (shared-memory windows of the event lists) and the
    thread-per-trajectory kernels execute the same walk -- trajectories, states and log-likelihoods bit for bit, including a
    tiny event capacity (overflow) and event lists longer than one window."""
    model, y, hmm, theta = _setup_any(dp, case, t0)
    dm = dp.device_model(hmm)
    rng = np.random.default_rng(4)
    for n, cap in ((70, 4096), (33, 40)):
        thetas = theta[:, None] * rng.uniform(0.8, 1.25, size=(len(theta), n))
        if t0:
            thetas[2] = rng.uniform(1.0, 8.0, n)
        theta_f = thetas * rng.uniform(0.85, 1.2, size=thetas.shape)
        valid = np.ones(n, dtype=bool); valid[3] = False
        res = []
        for mode in (1, 2):
            pt = dp.MbpParticles(dm, n, cap, seed=5)
            pt.set_mode(mode)
            lg = []
            for obs_i in range(1, len(y) + 1 if cap > 100 else 3):
                pt.set_stream_key(9000 + obs_i)
                lg.append(pt.iterate(thetas, obs_i, fresh=(obs_i == 1)))
            pt.set_stream_key(777)
            ll_f = pt.propose(thetas, theta_f, valid, len(lg))
            cur = [pt.get_particle(p + 1) for p in range(n)]
            prop = [pt.get_particle(p + 1, proposal=True) for p in range(n) if valid[p]]
            res.append((np.array(lg), ll_f, cur, prop))
        a, b = res
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1], equal_nan=True)
        for u, v in zip(a[2] + a[3], b[2] + b[3]):
            assert all(np.array_equal(x, z, equal_nan=True) for x, z in zip(u, v))
        if cap > 100:
            if case == "sis_pooley":
                assert max(len(c[1]) for c in a[2]) > 256  # event lists longer than one shared-memory window
        elif case == "sis_pooley":
            assert np.isneginf(a[0]).any()  # the tiny capacity overflows


@pytest.mark.parametrize("mode", [1, 2])  # one thread / one warp per trajectory
def test_store_grows_on_demand_and_results_do_not_depend_on_the_stride(dp, mode):
    """max_traj = MAX_TRAJ = 196000 (src/DiscretePOMP.jl:40) is the hard limit; the device store starts with a stride of 1024
    events per trajectory and doubles when a walk reaches it (the walk is not committed and re-runs with the same random
    streams).  LOTKA trajectories (~190 events per observation) cross 1024, 2048 and 4096 within 30 observations: every
    trajectory, final state and log-likelihood equals the one of a store that reserved 16384 events from the start."""
    model, y, hmm, theta = load_case(dp, "lotka_c4")
    dm = dp.device_model(hmm)
    n = 48
    th = np.tile(theta[:, None], (1, n)) * np.linspace(0.9, 1.1, n)[None, :]
    th_f = th * 1.03
    res = []
    for reserve in (None, 16384):
        pt = dp.MbpParticles(dm, n, seed=5)
        pt.set_mode(mode)
        assert pt.capacity() == (1024, 196000)
        if reserve:
            pt.reserve(reserve)
            assert pt.capacity()[0] == reserve
        lg = []
        for i in range(1, len(y) + 1):
            pt.set_stream_key(1000 + i)
            lg.append(pt.iterate(th, i, fresh=(i == 1)))
        pt.set_stream_key(77)
        ll = pt.propose(th, th_f, np.ones(n, dtype=np.uint8), len(y))
        cur = [pt.get_particle(p + 1) for p in (0, 7, n - 1)]
        prop = [pt.get_particle(p + 1, proposal=True) for p in (0, 7, n - 1)]
        res.append((np.array(lg), ll, cur, prop, pt.capacity()[0]))
    a, b = res
    assert a[4] >= 4096 and b[4] == 16384 and max(len(c[1]) for c in a[2]) > 4096  # the store did grow past several strides
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.all(np.isfinite(a[0]))
    for u, v in zip(a[2] + a[3], b[2] + b[3]):
        for x, z in zip(u, v):
            assert np.array_equal(x, z)


def test_device_resident_outer_layer_matches_the_host_driven_loop(dp):
    """dpomp_mbp_outer_*: theta, weights, priors and the accept test on the device.  Without mutation sweeps (n_props = 0) both
    drivers consume the same draws, so the resampled theta-particles are identical and the evidence agrees to rounding (the
    reductions associate differently); with sweeps the proposals come from Philox streams instead of numpy, so the two are
    compared by a z-test over replicates."""
    model, y, hmm, theta = _setup(dp)
    model.prior = dp.UniformProduct([0, 0], [0.01, 0.5])
    hmm = dp.get_private_model(model, y)
    th0 = model.prior.rand(3001, np.random.default_rng(2))
    a = dp.run_mbp_ibis(hmm, th0, 0.5, 0, False, 1.002, rng=np.random.default_rng(3), seed=4, verbose=False, device_outer=True)
    b = dp.run_mbp_ibis(hmm, th0, 0.5, 0, False, 1.002, rng=np.random.default_rng(3), seed=4, verbose=False, device_outer=False)
    assert np.array_equal(a.theta, b.theta) and np.allclose(a.weight, b.weight, rtol=1e-12)
    assert np.allclose(a.bme, b.bme, rtol=1e-11) and np.allclose(a.mu, b.mu, rtol=1e-11) and np.allclose(a.cv, b.cv, rtol=1e-9)
    # determinism of the device path
    a2 = dp.run_mbp_ibis(hmm, th0, 0.5, 3, False, 1.002, rng=np.random.default_rng(3), seed=4, verbose=False, device_outer=True)
    a3 = dp.run_mbp_ibis(hmm, th0, 0.5, 3, False, 1.002, rng=np.random.default_rng(3), seed=4, verbose=False, device_outer=True)
    assert np.array_equal(a2.theta, a3.theta) and np.array_equal(a2.bme, a3.bme) and a2.k_log[1] > 0
    reps = 6
    dev, host = [], []
    for s in range(reps):
        th = model.prior.rand(4000, np.random.default_rng(700 + s))
        r1 = dp.run_mbp_ibis(hmm, th, 0.5, 3, False, 1.002, seed=720 + s, verbose=False, device_outer=True)
        r2 = dp.run_mbp_ibis(hmm, th, 0.5, 3, False, 1.002, seed=740 + s, verbose=False, device_outer=False)
        dev.append(np.concatenate([r1.bme, r1.mu, [r1.k_log[1] / r1.k_log[0]]]))
        host.append(np.concatenate([r2.bme, r2.mu, [r2.k_log[1] / r2.k_log[0]]]))
    dev, host = np.array(dev), np.array(host)
    for j in range(dev.shape[1]):
        z = (dev[:, j].mean() - host[:, j].mean()) / np.sqrt(dev[:, j].var(ddof=1) / reps + host[:, j].var(ddof=1) / reps + 1e-300)
        assert abs(z) < 4.5, (j, dev[:, j], host[:, j])
