"""GPU tests of the resampling searches (dpomp_resample_indices): bit-exact 1-based ancestors against the oracle's
literal restatement of rs_* / rsp_* given identical weights and uniforms (north_star: "resampling ancestor indices are
bit-exact against the reference algorithm given identical uniforms and weights")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _weights(rng, n, kind):
    if kind == "uniform":
        return rng.random(n)
    if kind == "skewed":
        return np.exp(rng.normal(0, 4, n))
    if kind == "sparse":
        w = np.zeros(n); idx = rng.choice(n, max(1, n // 50), replace=False); w[idx] = rng.random(len(idx)); return w
    if kind == "ties":
        return np.repeat(rng.integers(0, 3, (n + 3) // 4), 4)[:n].astype(float) + (np.arange(n) == n - 1)
    if kind == "one":
        w = np.zeros(n); w[n // 3] = 2.5; return w
    raise ValueError(kind)


@pytest.mark.parametrize("rs_type", [1, 2, 3])
@pytest.mark.parametrize("kind", ["uniform", "skewed", "sparse", "ties", "one"])
def test_indices_bit_exact(dp, orc, rs_type, kind):
    rng = np.random.default_rng(rs_type * 100 + len(kind))
    for n in (1, 2, 7, 200, 4000, 8192, 16384, 100003):
        if rs_type == 3 and n > 16384:
            continue  # the literal multinomial oracle is O(n^2)
        w = _weights(rng, n, kind)
        r = rng.random(n)
        want = orc.rs(rs_type, w, r if rs_type != 1 else r[:1])
        if rs_type == 1:
            got = dp.rs_systematic(w.copy(), u=r[0])
        elif rs_type == 2:
            got = dp.rs_stratified(w.copy(), u=r)
        else:
            got = dp.rs_multinomial(w.copy(), u=r)
        assert got.dtype == np.int64 and np.array_equal(got, want), (rs_type, kind, n)
        # rsp_* semantics on cumulative weights (src/hmm_pf_resample.jl)
        cw = np.cumsum(w)
        assert np.array_equal(dp.rsp_indices(rs_type, cw, r if rs_type != 1 else r[:1]), orc.rsp(rs_type, cw, r if rs_type != 1 else r[:1]))


def test_edge_uniforms(dp, orc):
    # u exactly on bin edges, u = 0, and the largest uniform below 1
    w = np.array([1.0, 1.0, 1.0, 1.0])
    for u in (0.0, 0.5, 0.25, np.nextafter(1.0, 0.0)):
        assert np.array_equal(dp.rs_systematic(w.copy(), u=u), orc.rs(1, w, [u]))
    assert dp.rs_systematic(w.copy(), u=0.0).tolist() == [1, 1, 2, 3]
    assert dp.rs_systematic(np.zeros(5), u=0.3).tolist() == [1, 1, 1, 1, 1]
    # multinomial with n_out != n
    r = np.array([0.1, 0.9, 0.5])
    assert np.array_equal(dp.rs_multinomial(np.array([1.0, 2.0, 1.0, 4.0]), n=3, u=r), orc.rs(3, [1.0, 2.0, 1.0, 4.0], r, n_out=3))


def test_cumsum_side_effects_match_reference(dp):
    # rs_stratified / rs_multinomial cumsum! their argument in place, rs_systematic does not (src/hmm_resample.jl:5,45,67)
    w = np.array([1.0, 2.0, 3.0])
    dp.rs_systematic(w, u=0.5); assert w.tolist() == [1.0, 2.0, 3.0]
    dp.rs_stratified(w, u=np.array([0.5, 0.5, 0.5])); assert w.tolist() == [1.0, 3.0, 6.0]
