"""GPU parity tests of the particle-filter hot path, through the C ABI, against the CPU oracle.

Bars (BASELINE.json north_star):
  * ancestors / populations / event counts: bit-exact on identical Philox draws (f64 event loop)
  * per-trajectory log-likelihoods (log weights): <= 1e-6 relative (f64 loop: exact; f32 loop: on the trajectories that
    did not flip an event)
  * PF log-likelihood estimates: z-test against the oracle's Monte-Carlo distribution (f32 loop)
"""
import numpy as np
import pytest

from conftest import load_case, pure_death_case

pytestmark = pytest.mark.gpu


def _pf(dp, hmm, n, nb=1, rs=1, seed=1, f64=False, **kw):
    pf = dp.ParticleFilter(dp.device_model(hmm), n, nb, rs, seed=seed,
                           sim_precision=dp._capi.SIM_F64 if f64 else dp._capi.SIM_F32, **kw)
    pf.set_record_ancestors(True)  # diagnostics: log weights and ancestors of the last observation
    return pf


@pytest.mark.parametrize("case", ["sis_pooley", "sir_c2", "seir_c3", "lotka_c4"])
@pytest.mark.parametrize("rs_type", [1, 2, 3])
@pytest.mark.parametrize("scatter", [0, 1])  # offspring rows: the reference's order / chunk-interleaved over the tiles
def test_f64_loop_is_bit_exact_against_oracle(dp, orc, case, rs_type, scatter):
    model, y, hmm, theta = load_case(dp, case)
    ymax = min(len(y), 4)
    for n in (200, 1024, 3000, 4133):  # 256-particle tile, exactly one 1024 tile, ragged multi-tile (93 / 129 full chunks)
        if case == "lotka_c4" and n > 1024:
            continue
        if scatter and n <= 1024:  # one tile: the interleaved placement is the identity (covered by scatter = 0)
            continue
        pf = _pf(dp, hmm, n, rs=rs_type, f64=True)
        pf.set_record_ancestors(True)
        pf.set_scatter(scatter)
        tile, items = pf.geometry()
        key = 0xABCDEF00 + n + rs_type
        pf.set_stream_key(key)
        ll = pf.partial(theta, 1, ymax)[0]
        o_ll, o_lw, o_anc, o_ev, o_ovf, o_pop = orc.pf_partial(pf.dmodel.compiled.desc, theta, n, None, 1, ymax, rs_type,
                                                               key, 0, orc.MODE_DEVICE_INTERLEAVED if scatter else orc.MODE_DEVICE,
                                                               tile, items)
        assert pf.last_event_count() == o_ev
        assert np.array_equal(pf.last_logw(), o_lw)
        assert np.array_equal(pf.last_ancestors(), o_anc)
        assert np.array_equal(pf.get_pop(1), o_pop)
        assert abs(ll - o_ll) <= 1e-12 * max(1.0, abs(o_ll))


def test_literal_reference_arithmetic_gives_same_ancestors(dp, orc):
    # the oracle's LITERAL mode (running linear cumsum, sequential walk: the reference's own arithmetic) picks the same
    # ancestors as the device tree at these sizes, and the log-likelihood agrees to rounding
    model, y, hmm, theta = load_case(dp, "sis_pooley")
    n = 2000
    pf = _pf(dp, hmm, n, f64=True)
    pf.set_scatter(0)  # the reference's row order (offspring i in row i)
    key = 424242
    pf.set_stream_key(key)
    ll = pf.loglik(theta)[0]
    o = orc.pf_partial(pf.dmodel.compiled.desc, theta, n, None, 1, 5, 1, key, 0, orc.MODE_LITERAL)
    assert np.array_equal(pf.get_pop(1), o[5]) and abs(ll - o[0]) < 1e-10


def test_batched_filters_match_single_filters(dp, orc):
    # filter b of a batch == a single filter with global id b (batch_offset), bit for bit: sharding invariance
    model, y, hmm, theta = load_case(dp, "sir_c2")
    n, nb = 1500, 5
    rng = np.random.default_rng(1)
    thetas = theta[:, None] * rng.uniform(0.8, 1.2, size=(2, nb))
    pf = _pf(dp, hmm, n, nb, f64=True)
    key = 777
    pf.set_stream_key(key)
    lls = pf.partial(thetas, 1, 10)
    tile, items = pf.geometry()
    for b in (0, 3, 4):
        single = _pf(dp, hmm, n, 1, f64=True)
        single.set_batch_offset(b)
        single.set_stream_key(key)
        assert single.partial(thetas[:, b], 1, 10)[0] == lls[b]
        assert np.array_equal(single.get_pop(1), pf.get_pop(b + 1))
        o = orc.pf_partial(pf.dmodel.compiled.desc, thetas[:, b], n, None, 1, 10, 1, key, b, orc.MODE_DEVICE, tile, items)
        assert np.array_equal(pf.get_pop(b + 1), o[5]) and abs(lls[b] - o[0]) < 1e-11


def test_partial_calls_compose_on_device(dp, orc):
    # run_pibis' call pattern (src/hmm_ibis.jl:53-56): one observation per call on device-resident populations
    model, y, hmm, theta = load_case(dp, "sis_pooley")
    n = 1024
    pf = _pf(dp, hmm, n, f64=True)
    desc = pf.dmodel.compiled.desc
    tile, items = pf.geometry()
    pop = np.zeros((n, 2), dtype=np.int64)
    for i in range(1, 6):
        key = 1000 + i
        pf.set_stream_key(key)
        g = pf.partial(theta, i, i)[0]
        o = orc.pf_partial(desc, theta, n, pop, i, i, 1, key, 0, orc.MODE_DEVICE, tile, items)
        assert abs(g - o[0]) < 1e-12 and np.array_equal(pf.get_pop(1), pop)
    with pytest.raises(dp._capi.DpompError):
        _pf(dp, hmm, n).partial(theta, 2, 3)  # ymin > 1 on a filter that was never started


@pytest.mark.parametrize("case,n_keys,min_same", [("sir_c2", 1, 0.999), ("sir_c2", 8, 0.999), ("seir_c3", 8, 0.999),
                                                  ("sis_pooley", 1, 0.998), ("lotka_c4", 1, 0.998)])
def test_f32_loop_trajectories_match_within_1e6(dp, orc, case, n_keys, min_same):
    """Same Philox draws, f32 event loop against the f64 literal oracle over the first observation interval (n_keys
    independent stream keys x 4096 trajectories).  A trajectory whose f32 waiting times / event choices never cross a decision
    boundary ends in the identical integer state, and its log-weight -- f64 from integers -- is then identical (<= 1e-6
    relative asserted); the others differ by an event.  The fraction of identical trajectories is the per-trajectory statement
    of the f32 loop (2 events per interval for SIR / SEIR, ~200 for SIS-pooley and LOTKA); the f64 loop (DPOMP_SIM_F64) is
    identical on ALL of them (tests above)."""
    model, y, hmm, theta = load_case(dp, case)
    n = 4096
    key = 31337
    same_all, lw_all, ref_all = [], [], []
    for r in range(n_keys):
        pf = _pf(dp, hmm, n)
        pf.set_stream_key(key + r)
        pf.partial(theta, 1, 1)
        o = orc.pf_partial(pf.dmodel.compiled.desc, theta, n, None, 1, 1, 1, key + r, 0, orc.MODE_LITERAL)
        lw = pf.last_logw()
        same_all.append(lw == o[1]); lw_all.append(lw); ref_all.append(o[1])
    same, lw, ref = np.concatenate(same_all), np.concatenate(lw_all), np.concatenate(ref_all)
    print(f"{case}: {same.mean():.5f} of {same.size} f32 trajectories identical to the f64 oracle")
    assert same.mean() > min_same, (case, same.mean())
    fin = same & np.isfinite(ref)
    assert np.all(np.abs(lw[fin] - ref[fin]) <= 1e-6 * np.abs(ref[fin]))
    # the trajectories that differ do so by a few events: their log-weights stay finite and the weighted mean moves by << MC error
    both = np.isfinite(lw) & np.isfinite(ref)
    assert both.mean() > 0.999
    w_gpu, w_ref = np.exp(lw[both] - ref[both].max()), np.exp(ref[both] - ref[both].max())
    assert abs(np.log(w_gpu.mean()) - np.log(w_ref.mean())) < 5.0 / np.sqrt(both.sum())


@pytest.mark.parametrize("case,n,reps", [("sis_pooley", 200, 512), ("sir_c2", 1024, 128), ("seir_c3", 1024, 128),
                                         ("lotka_c4", 512, 64)])
def test_f32_loglik_z_test_against_oracle(dp, orc, case, n, reps):
    """z-test (|z| < 4.5) of the mean PF log-likelihood, f32 GPU path vs literal oracle, independent draws."""
    model, y, hmm, theta = load_case(dp, case)
    pf = _pf(dp, hmm, n, reps, seed=99)
    gpu = pf.loglik(np.tile(theta[:, None], (1, reps)))
    tile, items = pf.geometry()
    ref, _ = orc.pf_partial_batch(pf.dmodel.compiled.desc, np.tile(theta[:, None], (1, reps)), n, 1, len(y), 1,
                                  key=20261018, threads=orc.max_threads())
    z = (gpu.mean() - ref.mean()) / np.sqrt(gpu.var(ddof=1) / reps + ref.var(ddof=1) / reps)
    assert abs(z) < 4.5, (case, gpu.mean(), ref.mean(), z)
    assert 0.5 < gpu.std() / ref.std() < 2.0


def test_sis_pooley_anchor(dp):
    model, y, hmm, theta = load_case(dp, "sis_pooley")
    f = dp.get_particle_filter_lpdf(model, y, np=1 << 16, n_batch=8)
    lls = f(np.tile(theta[:, None], (1, 8)))
    assert abs(lls.mean() + 15.69) < 0.03  # SURVEY.md 8c anchor: -15.69 +- 0.01 as N -> inf
    assert isinstance(f(theta), float)


def test_determinism_and_seed_dependence(dp):
    model, y, hmm, theta = load_case(dp, "sir_c2")
    a = _pf(dp, hmm, 5000, seed=5).loglik(theta)[0]
    b = _pf(dp, hmm, 5000, seed=5).loglik(theta)[0]
    c = _pf(dp, hmm, 5000, seed=6).loglik(theta)[0]
    assert a == b and a != c
    pf = _pf(dp, hmm, 5000, seed=5)
    assert pf.loglik(theta)[0] == a and pf.loglik(theta)[0] != a  # consecutive calls use fresh streams


def test_event_cap_overflow_is_flagged(dp):
    model, y, hmm, theta = load_case(dp, "sis_pooley")
    pf = _pf(dp, hmm, 512, max_events=10)
    ll = pf.partial(theta, 1, 1)[0]
    assert pf.overflow_count() > 0 and np.isneginf(pf.last_logw()).sum() == pf.overflow_count()
    assert np.isfinite(ll) or np.isneginf(ll)


def test_zero_rate_absorption_and_all_zero_weights(dp, orc):
    # theta = 0: no events ever; the state stays at the initial condition and the weights are the closed form
    model, y, hmm, theta = load_case(dp, "sis_pooley")
    pf = _pf(dp, hmm, 300, f64=True)
    ll = pf.loglik(np.zeros(2))[0]
    assert pf.last_event_count() == 0 and np.all(pf.get_pop(1) == np.array([100, 1]))
    want = sum(np.log(1 / (np.sqrt(2 * np.pi) * 2.0)) - (v - 1) ** 2 / 8.0 for v in (18, 65, 70, 66, 67))
    assert abs(ll - want) < 1e-9
    # an observation so far away that every weight underflows in the linear domain: the reference returns -Inf,
    # the LSE path stays finite (documented divergence)
    far = [dp.Observation(o.time, 1, 1.0, [0, 1000]) for o in y]
    hmm2 = dp.get_private_model(model, far)
    ll2 = _pf(dp, hmm2, 300).loglik(np.zeros(2))[0]
    assert np.isfinite(ll2) and ll2 < -1e5


def test_obs_id_zero_skips_likelihood_and_resampling(dp, orc):
    model, y, hmm, theta = load_case(dp, "sis_pooley")
    y2 = [dp.Observation(o.time, 0 if i == 1 else 1, 1.0, o.val) for i, o in enumerate(y)]
    hmm2 = dp.get_private_model(model, y2)
    pf = _pf(dp, hmm2, 1024, f64=True)
    tile, items = pf.geometry()
    pf.set_stream_key(55)
    ll = pf.loglik(theta)[0]
    o = orc.pf_partial(pf.dmodel.compiled.desc, theta, 1024, None, 1, 5, 1, 55, 0, orc.MODE_DEVICE, tile, items)
    assert abs(ll - o[0]) < 1e-12 and np.array_equal(pf.get_pop(1), o[5])


def test_t0_index_parameter(dp, orc):
    model, y, hmm, theta = load_case(dp, "sis_pooley")
    def rf(out, p, x):
        out[0] = p[0] * x[0] * x[1]; out[1] = p[1] * x[1]
    m3 = dp.generate_custom_model("SIS", rf, [100, 1], [[-1, 1], [1, -1]], prior=dp.generate_weak_prior(3), t0_index=3)
    hmm3 = dp.get_private_model(m3, y)
    th = np.array([0.003, 0.1, 5.0])
    pf = _pf(dp, hmm3, 1024, f64=True)
    tile, items = pf.geometry()
    pf.set_stream_key(91)
    ll = pf.loglik(th)[0]
    o = orc.pf_partial(pf.dmodel.compiled.desc, th, 1024, None, 1, 5, 1, 91, 0, orc.MODE_DEVICE, tile, items)
    assert abs(ll - o[0]) < 1e-12 and np.array_equal(pf.get_pop(1), o[5])


def test_full_size_c2_properties(dp):
    """BASELINE config C2 at full size (2^20 particles x 100 observations): size-independent properties."""
    model, y, hmm, theta = load_case(dp, "sir_c2")
    n = 1 << 20
    pf = _pf(dp, hmm, n, seed=2)
    pf.set_record_ancestors(True)
    ll = pf.loglik(theta)[0]
    pop = pf.get_pop(1)
    assert np.all(pop.sum(axis=1) == 101) and pop.min() >= 0  # SIR conserves the population
    assert pf.overflow_count() == 0
    # the last resampling step (observation 99): ancestors are sorted (systematic) and cover 1..N
    pf.partial(theta, 1, 2)
    anc = pf.last_ancestors()
    pf2 = _pf(dp, hmm, n, seed=2)
    assert np.all(np.diff(anc) >= 0) and anc.min() >= 1 and anc.max() <= n
    # offspring counts follow the weights: |count_j - N w_j| < 1 for systematic resampling (first observation only)
    pf2.set_record_ancestors(True)
    pf2.partial(theta, 1, 1)
    lw = pf2.last_logw(); anc1 = pf2.last_ancestors()
    w = np.exp(lw - lw.max()); w /= w.sum()
    counts = np.bincount(anc1 - 1, minlength=n)
    assert np.max(np.abs(counts - n * w)) < 1.0 + 1e-6
    # two independent 2^20-particle estimates agree to Monte-Carlo error (sd of one estimate ~ 0.01)
    ll_b = _pf(dp, hmm, n, seed=3).loglik(theta)[0]
    assert abs(ll - ll_b) < 0.1


@pytest.mark.parametrize("name,ic,theta,fd", [("ROSSMAC", [50, 5, 60, 5], [0.02, 0.05, 0.05, 0.1, 0.5, 0.5], False),
                                              ("SIR", [100, 2, 0], [0.3, 0.1], True), ("SEIS", [60, 0, 3], [0.01, 0.3, 0.2], False),
                                              ("SI", [50, 1], [0.004], False), ("SEI", [80, 0, 2], [0.01, 0.4], True)])
def test_generic_rate_table_models_bit_exact(dp, orc, name, ic, theta, fd):
    """Models that run on the generic rate-table kernel (denominators, six events) or on other built-in structures."""
    model = dp.generate_model(name, ic, freq_dep=fd)
    c = len(ic)
    val = [0] * c
    val[(2 if name in ("SEI", "SEIS") else 1)] = 4
    y = [dp.Observation(float(t), 1, 1.0, val) for t in (1.0, 2.5, 4.0)]
    hmm = dp.get_private_model(model, y)
    th = np.asarray(theta)
    for n in (700, 2500):
        pf = _pf(dp, hmm, n, f64=True)
        tile, items = pf.geometry()
        pf.set_stream_key(4242 + n)
        ll = pf.loglik(th)[0]
        o = orc.pf_partial(pf.dmodel.compiled.desc, th, n, None, 1, 3, 1, 4242 + n, 0, orc.MODE_DEVICE, tile, items)
        assert pf.last_event_count() == o[3] and np.array_equal(pf.get_pop(1), o[5]) and abs(ll - o[0]) <= 1e-11 * max(1, abs(o[0]))
    # f32 loop: same law (mean log-lik of 64 filters within 5 sd of the oracle's)
    pf = _pf(dp, hmm, 1024, 64, seed=8)
    g = pf.loglik(np.tile(th[:, None], (1, 64)))
    r, _ = orc.pf_partial_batch(pf.dmodel.compiled.desc, np.tile(th[:, None], (1, 64)), 1024, 1, 3, 1, key=99, threads=orc.max_threads())
    assert abs(g.mean() - r.mean()) < 5 * np.sqrt(g.var(ddof=1) / 64 + r.var(ddof=1) / 64) + 1e-9


def test_more_tiles_than_one_scan_chunk_bit_exact(dp, orc):
    """N > 1024 tiles of 1024 particles: the per-filter combine scans the tile partials in several chunks."""
    model, y, hmm, theta = load_case(dp, "sir_c2")
    n = 1024 * 1024 + 5000  # 1029 tiles, ragged last tile
    pf = _pf(dp, hmm, n, f64=True)
    tile, items = pf.geometry()
    pf.set_stream_key(2026)
    ll = pf.partial(theta, 1, 2)[0]
    o = orc.pf_partial(pf.dmodel.compiled.desc, theta, n, None, 1, 2, 1, 2026, 0, orc.MODE_DEVICE, tile, items,
                       threads=orc.max_threads())
    assert pf.last_event_count() == o[3]
    assert np.array_equal(pf.last_ancestors(), o[2]) and np.array_equal(pf.get_pop(1), o[5])
    assert abs(ll - o[0]) <= 1e-12 * abs(o[0])


def test_degenerate_weights_collapse_to_few_ancestors(dp, orc):
    """An observation that only a handful of particles explain: one ancestor tile feeds (almost) every offspring, so a
    single warp of the resample kernel loops over many windows.  Bit-exact against the oracle."""
    model, y, hmm, theta = load_case(dp, "sir_c2")
    y2 = [dp.Observation(1.0, 1, 1.0, [0, 9, 0]), dp.Observation(2.0, 1, 1.0, [0, 9, 0])]  # I = 9 after one time unit is rare
    model2 = dp.generate_model("SIR", [100, 1, 0], obs_error=0.25)
    hmm2 = dp.get_private_model(model2, y2)
    n = 20000
    pf = _pf(dp, hmm2, n, f64=True)
    tile, items = pf.geometry()
    pf.set_stream_key(77)
    ll = pf.partial(theta, 1, 1)[0]
    o = orc.pf_partial(pf.dmodel.compiled.desc, theta, n, None, 1, 1, 1, 77, 0, orc.MODE_DEVICE, tile, items)
    anc = pf.last_ancestors()
    assert np.array_equal(anc, o[2]) and np.array_equal(pf.get_pop(1), o[5]) and abs(ll - o[0]) < 1e-9
    assert len(np.unique(anc)) < n // 20  # heavy collapse


@pytest.mark.parametrize("case,n,nb,f64", [("sir_c2", 3000, 3, True), ("sir_c2", 200, 5, False), ("seir_c3", 70000, 2, False),
                                            ("lotka_c4", 1024, 4, False), ("sis_pooley", 1 << 18, 1, False),
                                            ("sis_pooley", 200, 1300, False)])
@pytest.mark.parametrize("rs_type", [1, 2])
def test_fused_step_kernel_equals_two_kernel_path(dp, case, n, nb, f64, rs_type):
    """The fused simulate+resample launch and the two-kernel path are the same computation: bit-identical log-likelihoods,
    populations and ancestors (f32 and f64 loops, ragged tiles, several filters; 1300 one-tile filters = more CTAs than the
    device holds at once: the arrival-order tickets of the fused kernel, which co-resident launches skip)."""
    model, y, hmm, theta = load_case(dp, case)
    thetas = theta[:, None] * np.linspace(0.9, 1.1, nb)[None, :]
    out = []
    for fused in (True, False):
        pf = _pf(dp, hmm, n, nb, rs=rs_type, f64=f64, seed=4)
        pf.set_fused(fused)
        pf.set_stream_key(555)
        ll = pf.partial(thetas, 1, min(len(y), 6))
        out.append((ll, [pf.get_pop(b + 1) for b in range(nb)], pf.last_ancestors(nb), pf.last_timing()[1]))
    (ll_a, pops_a, anc_a, launches_a), (ll_b, pops_b, anc_b, launches_b) = out
    assert np.array_equal(ll_a, ll_b) and np.array_equal(anc_a, anc_b)
    assert all(np.array_equal(u, v) for u, v in zip(pops_a, pops_b))
    assert launches_a < launches_b  # one launch per resampling observation instead of two


@pytest.mark.parametrize("case,n,nb", [("sir_c2", 70000, 3), ("seir_c3", 1 << 20, 1), ("sis_pooley", 3000, 7)])
@pytest.mark.parametrize("rs_type", [1, 2, 3])
def test_deferred_level2_equals_two_ticket_levels(dp, case, n, nb, rs_type, monkeypatch):
    """Latency regime: level 2 of the weight combine runs in the resample kernel (every CTA for itself, same tree) instead of
    behind a second ticket level in the simulate kernel.  DPOMP_DEFER_L2=0 (read when the handle is created) switches it off:
    log-likelihoods, populations and ancestors are bit-identical for all three resamplers (multinomial: the gather kernel
    reads the totals the tile-0 CTA of the resample kernel stored)."""
    model, y, hmm, theta = load_case(dp, case)
    thetas = theta[:, None] * np.linspace(0.95, 1.05, nb)[None, :]
    out = []
    for knob in ("1", "0"):
        monkeypatch.setenv("DPOMP_DEFER_L2", knob)
        pf = _pf(dp, hmm, n, nb, rs=rs_type, seed=8)
        pf.set_fused(0)
        pf.set_stream_key(4242)
        ll = pf.partial(thetas, 1, min(len(y), 5))
        out.append((ll, [pf.get_pop(b + 1) for b in range(nb)], pf.last_ancestors(nb)))
    (ll_a, pops_a, anc_a), (ll_b, pops_b, anc_b) = out
    assert np.array_equal(ll_a, ll_b) and np.array_equal(anc_a, anc_b)
    assert all(np.array_equal(u, v) for u, v in zip(pops_a, pops_b))


@pytest.mark.parametrize("case,n,nb", [("lotka_c4", 4096, 6), ("sir_c2", 5000, 3), ("seir_c3", 70000, 2), ("sis_pooley", 3000, 4)])
@pytest.mark.parametrize("knob", ["1", "auto"])
def test_two_per_lane_loop_on_large_tiles_is_bit_identical(dp, case, n, nb, knob, monkeypatch):
    """1024-particle tiles of the predefined models: the two-particles-per-lane loop with the Philox words drawn ahead (picked
    for event-heavy models in under-filled launches from the event intensity of the handle's previous call) uses the same
    counters and draws as the one-particle-per-lane loop: log-likelihoods, populations and ancestors are bit-identical, forced
    (DPOMP_TWO_PER_LANE=1) and selected automatically (second call of the handle; LOTKA / SIS-pooley qualify, SIR / SEIR do not)."""
    model, y, hmm, theta = load_case(dp, case)
    thetas = theta[:, None] * np.linspace(0.95, 1.05, nb)[None, :]
    out = []
    for k in (knob, "0"):
        if k == "auto":
            monkeypatch.delenv("DPOMP_TWO_PER_LANE", raising=False)
        else:
            monkeypatch.setenv("DPOMP_TWO_PER_LANE", k)
        pf = _pf(dp, hmm, n, nb, seed=8)
        res = []
        for call in range(2):  # the second call of a handle knows the event intensity of the first
            pf.set_stream_key(900 + call)
            ll = pf.partial(thetas, 1, min(len(y), 4))
            res.append((ll, [pf.get_pop(b + 1) for b in range(nb)], pf.last_ancestors(nb)))
        out.append(res)
    for (ll_a, pops_a, anc_a), (ll_b, pops_b, anc_b) in zip(*out):
        assert np.array_equal(ll_a, ll_b) and np.array_equal(anc_a, anc_b)
        assert all(np.array_equal(u, v) for u, v in zip(pops_a, pops_b))


@pytest.mark.parametrize("variant", ["wide_spread", "huge_observation", "tiny_sigma"])
def test_weight_pass_paths_are_bit_exact_against_oracle(dp, orc, variant):
    """The integer-domain weight pass of the simulate kernel (tile maximum by integer min, per-CTA exp table indexed by
    |y - x| - d_min) against the oracle's direct f64 evaluation, on the paths beside the table: offsets beyond its 128
    entries (population 1000: |y - x| spreads over hundreds), observations >= 2^30 (the direct f64 path), and a sigma for
    which 2 sigma^2 is not a power of two (the f64 division instead of the multiplication)."""
    if variant == "wide_spread":
        model, y, hmm, theta = load_case(dp, "sir_dense")
        # one long interval without resampling up to the epidemic peak: infectives range from 0 (early extinction) to hundreds
        y = [dp.Observation(40.0, 1, 1.0, [0, 300, 0]), dp.Observation(41.0, 1, 1.0, [0, 300, 0])]
        hmm = dp.get_private_model(model, y)
    elif variant == "huge_observation":
        model, y, hmm, theta = load_case(dp, "sis_pooley")
        y = [dp.Observation(20.0, 1, 1.0, [0, 18]), dp.Observation(40.0, 1, 1.0, [0, (1 << 31) + 5]), dp.Observation(60.0, 1, 1.0, [0, 70])]
        hmm = dp.get_private_model(model, y)
    else:
        model, y, hmm, theta = load_case(dp, "sis_pooley")
        model = dp.generate_model("SIS", [100, 1], obs_error=1.7)
        hmm = dp.get_private_model(model, y)
    for n in (1500, 5000):
        pf = _pf(dp, hmm, n, f64=True)
        tile, items = pf.geometry()
        key = 0x5EED + n
        pf.set_stream_key(key)
        ymax = len(y) - 1  # the last evaluated step is followed by a resampling step (obs_i < length(obs_data))
        ll = pf.partial(theta, 1, ymax)[0]
        o_ll, o_lw, o_anc, o_ev, o_ovf, o_pop = orc.pf_partial(pf.dmodel.compiled.desc, theta, n, None, 1, ymax, 1, key, 0,
                                                               orc.MODE_DEVICE, tile, items)
        assert np.array_equal(pf.last_logw(), o_lw) and np.array_equal(pf.last_ancestors(), o_anc)
        assert np.array_equal(pf.get_pop(1), o_pop)
        assert (np.isnan(ll) and np.isnan(o_ll)) or ll == o_ll or abs(ll - o_ll) <= 1e-12 * max(1.0, abs(o_ll))
        if variant == "wide_spread":
            lw = pf.last_logw()
            d = np.sqrt(np.maximum(0.0, (lw.max() - lw[np.isfinite(lw)]) * 8.0))  # sigma = 2: logw = c - d^2 / 8
            assert d.max() > 160  # offsets well beyond the 128-entry table


@pytest.mark.parametrize("f64", [False, True])
@pytest.mark.parametrize("rs_type", [1, 2, 3])
def test_pf_loglik_against_exact_forward_algorithm(dp, f64, rs_type):
    """The CUDA filter against an EXACT likelihood (no oracle, no reference statistic): pure-death process, p(y_1..y_5) by
    the forward recursion over its 61 hidden states; the PF likelihood estimate is unbiased, 2^18 particles x 8 filters."""
    model, y, hmm, theta, ll_exact = pure_death_case(dp)
    nb = 8
    pf = dp.ParticleFilter(dp.device_model(hmm), 1 << 18, nb, rs_type, seed=17 + rs_type,
                           sim_precision=dp._capi.SIM_F64 if f64 else dp._capi.SIM_F32)
    lls = pf.loglik(np.tile(theta[:, None], (1, nb)))
    assert abs(np.log(np.mean(np.exp(lls - ll_exact)))) < 0.004, (lls, ll_exact)
    assert np.all(np.abs(lls - ll_exact) < 0.02)


def test_f32_event_uniform_is_strictly_below_one(dp):
    """the f32 event-choice uniform must stay below 1 so that choose_event (src/hmm_cmn.jl:4-10) never falls
    through to a zero-rate last event; the waiting-time uniform must stay above 0 (finite log)."""
    import ctypes as C
    words = np.array([0, 1, 127, 128, 0x7FFFFFFF, 0x80000000, 0xFFFFFF00, 0xFFFFFF7F, 0xFFFFFF80, 0xFFFFFFFE, 0xFFFFFFFF],
                     dtype=np.uint32)
    wait, evt = np.zeros(len(words), dtype=np.float32), np.zeros(len(words), dtype=np.float32)
    dp._capi.check(dp._capi.lib().dpomp_debug_uniforms_f32(dp._capi.ptr(words), len(words), dp._capi.ptr(wait), dp._capi.ptr(evt)))
    assert np.all(wait > 0) and np.all(wait <= 1)
    assert np.all(evt >= 0) and np.all(evt < 1) and evt.max() == np.float32(1 - 2.0 ** -24)
    # fl(u * R) < R for every positive total rate R: some event with a positive rate is always chosen
    rates = np.float32(np.random.default_rng(0).uniform(1e-6, 1e6, 4096))
    rates = np.concatenate([rates, np.float32([1.0, 2.0, 3.0, 0.1, 1e-30, 1.5e38])])
    assert np.all(evt.max() * rates < rates)


def test_f32_loop_never_produces_negative_compartments(dp):
    """SEIR (last event I->R has rate 0 while I == 0 and E > 0): 2^22 particle-steps, no compartment may go negative and
    the population size is conserved."""
    model, y, hmm, theta = load_case(dp, "seir_c3")
    pf = _pf(dp, hmm, 1 << 18, seed=5)
    pf.partial(theta, 1, 16)
    pop = pf.get_pop(1)
    assert pop.min() >= 0 and np.all(pop.sum(axis=1) == 101)


def test_multinomial_seam_matches_flat_search(dp, orc):
    """a multinomial draw between the last cw of a tile and the tile's end belongs to the NEXT tile's first
    particle with cw > chs (src/hmm_resample.jl:9-16); ancestors equal the oracle's on many multi-tile steps."""
    model, y, hmm, theta = load_case(dp, "sir_c2")
    n = 5000
    for key in range(900, 906):
        pf = _pf(dp, hmm, n, rs=3, f64=True)
        tile, items = pf.geometry()
        pf.set_stream_key(key)
        pf.partial(theta, 1, 6)
        o = orc.pf_partial(pf.dmodel.compiled.desc, theta, n, None, 1, 6, 3, key, 0, orc.MODE_DEVICE, tile, items)
        assert np.array_equal(pf.last_ancestors(), o[2]) and np.array_equal(pf.get_pop(1), o[5])


def test_interleaved_scatter_is_a_row_permutation_of_the_reference_order(dp):
    """One resampling step from identical inputs: the interleaved placement holds exactly the reference-order offspring,
    32-row chunk k moved to chunk sigma(k) (include/dpomp.h), the trailing partial chunk in place; ancestors move with
    their rows.  Also at 2^17 particles x 2 filters (multi-group combine)."""
    model, y, hmm, theta = load_case(dp, "sir_c2")
    for n, nb in ((4133, 1), (1 << 17, 2)):
        pops, ancs = [], []
        for mode in (0, 1):
            pf = _pf(dp, hmm, n, nb=nb, f64=False, seed=9)
            pf.set_scatter(mode)
            pf.set_stream_key(77)
            pf.partial(np.tile(theta[:, None], (1, nb)), 1, 1)  # observation 1 resamples (obs_id > 0, not the last)
            pops.append(pf.get_pop(nb)); ancs.append(pf.last_ancestors(nb))
        tile, _ = pf.geometry()
        m = -(-n // tile)
        ncf = n >> 5
        q, r = divmod(ncf, m)
        k = np.arange(n) >> 5
        rr, qq = k % m, k // m
        row = np.where(k < ncf, ((rr * q + np.minimum(rr, r) + qq) << 5) | (np.arange(n) & 31), np.arange(n))
        assert np.array_equal(np.sort(row), np.arange(n))  # a bijection
        assert np.array_equal(pops[1][row], pops[0]) and np.array_equal(ancs[1][row], ancs[0])


def test_interleaved_scatter_estimate_has_the_same_law(dp):
    """z-test of the PF log-likelihood estimate between the two row orders (f32 loop, 64 replicate filters each)."""
    model, y, hmm, theta = load_case(dp, "seir_c3")
    lls = []
    for mode in (0, 1):
        pf = _pf(dp, hmm, 8192, nb=64, seed=31 + mode)
        pf.set_scatter(mode)
        lls.append(pf.loglik(np.tile(theta[:, None], (1, 64))))
    z = (lls[0].mean() - lls[1].mean()) / np.sqrt(lls[0].var(ddof=1) / 64 + lls[1].var(ddof=1) / 64)
    assert abs(z) < 4.5, (lls[0].mean(), lls[1].mean(), z)


@pytest.mark.parametrize("case,n,nb", [("sir_c2", 3000, 3), ("sir_c2", 1 << 17, 1), ("seir_c3", 70000, 2), ("lotka_c4", 2048, 4),
                                       ("sis_pooley", 200, 5), ("sis_pooley", 1024, 1)])
@pytest.mark.parametrize("rs_type", [1, 2])
def test_persistent_kernel_equals_launch_chain(dp, case, n, nb, rs_type):
    """One cooperative launch for all observations of a call (dataflow between tiles instead of kernel boundaries) is the
    same computation as the per-observation launch chain: bit-identical log-likelihoods, populations, ancestors and event
    counts; ragged tiles, several filters, several groups of tiles, partial calls that compose."""
    model, y, hmm, theta = load_case(dp, case)
    thetas = theta[:, None] * np.linspace(0.9, 1.1, nb)[None, :]
    ymax = min(len(y), 12)
    split = max(1, ymax // 2)
    out = []
    for mode in (2, 0):
        pf = _pf(dp, hmm, n, nb, rs=rs_type, seed=4)
        pf.set_persistent(mode)
        pf.set_fused(0)
        pf.set_stream_key(555)
        ll1 = pf.partial(thetas, 1, split)
        l1 = pf.last_timing()[1]
        ev1 = pf.last_event_count()
        pf.set_stream_key(556)
        ll2 = pf.partial(thetas, split + 1, ymax) if split < ymax else np.zeros(nb)
        out.append((ll1, ll2, [pf.get_pop(b + 1) for b in range(nb)], pf.last_ancestors(nb), ev1, pf.last_event_count(), l1, pf.overflow_count()))
    a, b = out
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert all(np.array_equal(u, v) for u, v in zip(a[2], b[2]))
    if ymax < len(y):  # the last observation of the call resampled: ancestors are defined
        assert np.array_equal(a[3], b[3])
    assert a[4] == b[4] and a[5] == b[5] and a[7] == b[7] == 0
    assert a[6] == 1 and b[6] > 1  # one launch for the whole call


def test_persistent_kernel_with_non_likelihood_observations(dp):
    """obs_id <= 0 observations neither weight nor resample (src/hmm_particle_filter.jl:58): the persistent kernel keeps the
    tile in shared memory across them; the call's last observation is the data set's last (no resampling after it)."""
    model = dp.generate_model("SIR", [100, 1, 0])
    ys = [(1.0, 1, 3), (2.0, 0, 0), (3.0, 1, 6), (4.0, 0, 0), (5.0, 0, 0), (6.0, 1, 12), (7.0, 1, 14)]
    y = [dp.Observation(t, oid, 1.0, [0, v, 0]) for t, oid, v in ys]
    hmm = dp.get_private_model(model, y)
    res = []
    for mode in (2, 0):
        pf = _pf(dp, hmm, 5000, 2, seed=8)
        pf.set_persistent(mode)
        pf.set_stream_key(99)
        ll = pf.loglik(np.array([[0.003, 0.0035], [0.1, 0.12]]))
        res.append((ll, pf.get_pop(1), pf.get_pop(2), pf.last_logw(2)))
    assert np.array_equal(res[0][0], res[1][0]) and np.all(np.isfinite(res[0][0]))
    for k in (1, 2, 3):
        assert np.array_equal(res[0][k], res[1][k])


def test_small_filter_latency_kernel_equals_throughput_kernel(dp):
    """Filters of <= 256 particles run two particles per lane at once when the launch is small (<= 296 CTAs: the latency
    regime of the reference's default 200-particle filter) and one at a time in large batches: identical results, because the
    random streams are keyed by (filter, particle, event), not by the schedule."""
    model, y, hmm, theta = load_case(dp, "sis_pooley")
    dm = dp.device_model(hmm)
    nb_big = 400
    thetas = theta[:, None] * np.linspace(0.8, 1.2, nb_big)[None, :]
    big = dp.ParticleFilter(dm, 200, nb_big, 1, seed=3)
    big.set_stream_key(4242)
    ll_big = big.loglik(thetas)
    ids = np.array([10, 11, 12, 399])
    small = dp.ParticleFilter(dm, 200, len(ids), 1, seed=3)
    small.set_filter_ids(ids)
    small.set_stream_key(4242)
    ll_small = small.loglik(thetas[:, ids])
    assert np.array_equal(ll_small, ll_big[ids])
    for j, b in enumerate(ids):
        assert np.array_equal(small.get_pop(j + 1), big.get_pop(int(b) + 1))
