"""
Keeps the selectable kernel flavours honest: every DMMA tile-kernel generation (v1 cp.async,
v2 TMA/mbarrier, v2 with 16 consumer warps, v3 persistent) must produce the same normal matrix and
forces, and the fused in-shared-memory Gram variant must reproduce the staged Gram matrix to the
accuracy of the fast Green's function.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import verde_b200 as vb
    from verde_b200 import _cabi, kernels

    return vb, kernels, _cabi.require_device()


def sample(n, seed=0):
    rng = np.random.RandomState(seed)
    e, nn = rng.uniform(0, 5000, n), rng.uniform(-5000, 0, n)
    return e, nn, 1000.0 * np.sin(e / 400.0) * np.cos(nn / 400.0)


def test_gemm_generations_agree(env):
    vb, kernels, lib = env
    e, nn, d = sample(1100)
    w = np.random.RandomState(1).uniform(0.5, 2, e.size)
    results = {}
    default = vb._cabi.get_option("gemm_variant")
    try:
        for variant in (3, 0, 1, 2, 4):
            assert lib.vb200_set_option(b"gemm_variant", float(variant)) == 0
            system = kernels.GramSystem(e.size).accumulate(0, e, nn, d, w, e, nn, 1e-5)
            gram = system.gram.cpu().numpy()
            results[variant] = (gram, system.solve(e.size, 1e-3))
        # the dynamic tile queue with a capped grid (SMs left to a concurrent stream): same tiles, same bits
        assert lib.vb200_set_option(b"gemm_grid_limit", 37.0) == 0
        system = kernels.GramSystem(e.size).accumulate(0, e, nn, d, w, e, nn, 1e-5)
        results["4 capped"] = (system.gram.cpu().numpy(), system.solve(e.size, 1e-3))
    finally:
        lib.vb200_set_option(b"gemm_variant", default)
        lib.vb200_set_option(b"gemm_grid_limit", 0.0)
    g_ref, f_ref = results[3]
    for variant in (0, 1, 2):
        gram, force = results[variant]
        np.testing.assert_allclose(gram, g_ref, rtol=1e-13, atol=1e-13 * np.abs(g_ref).max())
        assert np.max(np.abs(force - f_ref)) <= 1e-9 * np.max(np.abs(f_ref)), variant
    for variant in (4, "4 capped"):  # one add per element per launch, whichever CTA draws the tile: bit-identical
        np.testing.assert_array_equal(results[variant][0], g_ref)
        np.testing.assert_array_equal(results[variant][1], f_ref)


@pytest.mark.parametrize("mindist,weighted", [(0.0, False), (1e-5, True)])
def test_fused_in_smem_gram_matches_staged(env, mindist, weighted):
    vb, kernels, lib = env
    e, nn, d = sample(900, seed=3)
    fe, fn, _ = sample(300, seed=4)
    w = np.random.RandomState(2).uniform(0.5, 2, e.size) if weighted else None
    staged = kernels.GramSystem(fe.size).accumulate(0, e, nn, d, w, fe, fn, mindist)
    try:
        assert lib.vb200_set_option(b"gram_fused", 1.0) == 0
        fused = kernels.GramSystem(fe.size).accumulate(0, e, nn, d, w, fe, fn, mindist)
    finally:
        lib.vb200_set_option(b"gram_fused", 0.0)
    g0, g1 = staged.gram.cpu().numpy(), fused.gram.cpu().numpy()
    # tile-packed lower triangle; compare on the scale of the diagonal (entries span many decades)
    assert np.max(np.abs(g1 - g0)) <= 1e-12 * np.abs(g0).max()
    np.testing.assert_array_equal(staged.stats.cpu().numpy(), fused.stats.cpu().numpy())
    f0, f1 = staged.solve(e.size, 1e-2), fused.solve(e.size, 1e-2)
    assert np.max(np.abs(f1 - f0)) <= 1e-6 * np.max(np.abs(f0))


def test_single_launch_sweeps_match_per_block_kernels(env):
    vb, kernels, lib = env
    e, nn, d = sample(2100, seed=7)
    forces = []
    try:
        for variant in (1, 0):
            assert lib.vb200_set_option(b"solve_variant", float(variant)) == 0
            forces.append(kernels.fit_spline(e, nn, d, None, e, nn, 0.0, 1e-2))
    finally:
        lib.vb200_set_option(b"solve_variant", 1.0)
    assert np.max(np.abs(forces[1] - forces[0])) <= 1e-6 * np.max(np.abs(forces[0]))


def test_chol_panel_width_does_not_change_the_solution(env):
    vb, kernels, lib = env
    e, nn, d = sample(1500, seed=5)
    forces = []
    try:
        for width in (1, 3, 8):
            assert lib.vb200_set_option(b"chol_panel_tiles", float(width)) == 0
            forces.append(kernels.fit_spline(e, nn, d, None, e, nn, 0.0, 1e-2))
    finally:
        lib.vb200_set_option(b"chol_panel_tiles", 8.0)
    for f in forces[1:]:
        # different panel widths reorder the trailing updates: agreement to ~cond*eps (cond ~ 3e9 here)
        assert np.max(np.abs(f - forces[0])) <= 1e-6 * np.max(np.abs(forces[0]))


@pytest.mark.parametrize("n", [700, 2100])
def test_batched_damping_sweep_equals_one_solve_per_damping(env, n):
    """vb200_gram_solve_multi factors all dampings side by side (batched panel kernels, one DMMA launch per
    update for the whole batch, one cooperative sweep): same arithmetic per system -> same bits."""
    vb, kernels, lib = env
    e, nn, d = sample(n, seed=11)
    w = np.random.RandomState(3).uniform(0.5, 2, e.size)
    dampings = [1e-6, 1e-4, 1e-3, 1e-2, 1e-1, 1.0, 10.0]
    system = kernels.GramSystem(e.size).accumulate(0, e, nn, d, w, e, nn, 1e-5)
    gram_before = system.gram.clone()
    batched = system.solve_multi(e.size, dampings)
    assert batched.shape == (len(dampings), e.size)
    assert bool((system.gram == gram_before).all()), "the sweep must leave the Gram matrix untouched"
    for k, damping in enumerate(dampings):
        single = system.solve(e.size, damping, keep=True)
        np.testing.assert_array_equal(batched[k], single)
    # the older tile kernels walk the batch matrix by matrix: same answer
    default = vb._cabi.get_option("gemm_variant")
    try:
        assert lib.vb200_set_option(b"gemm_variant", 1.0) == 0
        v2 = system.solve_multi(e.size, dampings[:3])
        assert lib.vb200_set_option(b"gemm_variant", 4.0) == 0
        v4 = system.solve_multi(e.size, dampings)
    finally:
        lib.vb200_set_option(b"gemm_variant", default)
    np.testing.assert_array_equal(v4, batched)  # dynamic tile queue over the batch: bit-identical
    assert np.max(np.abs(v2 - batched[:3])) <= 1e-9 * np.max(np.abs(batched[:3]))


def test_batched_sweep_reports_a_non_positive_definite_member(env):
    vb, kernels, lib = env
    e, nn, d = sample(300, seed=12)
    system = kernels.GramSystem(e.size).accumulate(0, e, nn, d, None, e, nn, 0.0)
    system.gram.mul_(-1.0)  # -G - tiny*I is negative definite, -G + huge*I is fine
    with pytest.raises(np.linalg.LinAlgError):
        system.solve_multi(e.size, [1e-30, 1e30])


@pytest.mark.parametrize("mindist", [0.0, 1e-5])
def test_wide_predict_kernel_is_bit_identical(env, mindist):
    """predict_wide = 1: 512-thread CTAs with the conflict-free 8-copy log table; the lanes split the forces
    exactly as in the default kernel, so the sums are the same bits."""
    vb, kernels, lib = env
    rng = np.random.RandomState(21)
    npts, m = 80_000, 3000  # large enough for the 8-points-per-warp path the wide kernel replaces
    e, nn = rng.uniform(0, 5000, npts), rng.uniform(-5000, 0, npts)
    fe, fn, f = rng.uniform(0, 5000, m), rng.uniform(-5000, 0, m), rng.normal(size=m)
    outs = []
    try:
        for wide in (0.0, 1.0):
            assert lib.vb200_set_option(b"predict_wide", wide) == 0
            outs.append(kernels.predict_b200(e, nn, fe, fn, mindist, f, np.empty(npts)).copy())
    finally:
        lib.vb200_set_option(b"predict_wide", 0.0)
    np.testing.assert_array_equal(outs[0], outs[1])


def test_batched_sweep_in_groups_when_memory_is_short(env):
    "sweep_group caps how many damped systems are factored side by side (what a nearly full HBM does)"
    vb, kernels, lib = env
    e, nn, d = sample(900, seed=13)
    dampings = [1e-5, 1e-4, 1e-3, 1e-2, 1e-1]
    system = kernels.GramSystem(e.size).accumulate(0, e, nn, d, None, e, nn, 0.0)
    whole = system.solve_multi(e.size, dampings)
    try:
        assert lib.vb200_set_option(b"sweep_group", 2.0) == 0
        grouped = system.solve_multi(e.size, dampings)  # batches of 2, 2, 1
    finally:
        lib.vb200_set_option(b"sweep_group", 0.0)
    np.testing.assert_array_equal(grouped, whole)


@pytest.mark.parametrize("n,panel_tiles", [(700, 8), (2100, 3), (2100, 8)])
def test_stepwise_cholesky_equals_the_fused_solve(env, n, panel_tiles):
    """vb200_chol_begin / chol_panel / chol_update / chol_finish (the multi-GPU panel-cyclic driver) on one
    rank: every panel is owned locally, no broadcast, and the arithmetic is that of vb200_gram_solve."""
    vb, kernels, lib = env
    e, nn, d = sample(n, seed=17)
    system = kernels.GramSystem(e.size).accumulate(0, e, nn, d, None, e, nn, 1e-5)
    fused = system.solve(e.size, 1e-3, keep=True)
    gram = system.gram.clone()
    phases = {}
    stepwise = system.solve_distributed(e.size, 1e-3, panel_tiles=panel_tiles, phase_times=phases)
    system.gram.copy_(gram)
    if panel_tiles == 8:  # the fused solve's panel width: same launches, same bits
        np.testing.assert_array_equal(stepwise, fused)
    else:
        assert np.max(np.abs(stepwise - fused)) <= 1e-6 * np.max(np.abs(fused))
    assert set(phases) == {"panel", "broadcast", "update"} and system.info.factor_ms > 0
    # the panel ranges tile the packed triangle exactly
    import ctypes
    T = (n + 127) // 128
    off, cnt, end = ctypes.c_int64(), ctypes.c_int64(), 0
    for k0 in range(0, T, panel_tiles):
        w = min(panel_tiles, T - k0)
        assert lib.vb200_chol_panel_range(n, k0, w, ctypes.byref(off), ctypes.byref(cnt)) == 0
        assert off.value == end
        end = off.value + cnt.value
    assert end == lib.vb200_gram_doubles(n)
    # misuse is refused: no begin -> no panel / finish
    assert lib.vb200_chol_panel(system.gram.data_ptr(), n, 0, 1) != 0
    assert lib.vb200_chol_panel_range(n, T, 1, ctypes.byref(off), ctypes.byref(cnt)) != 0


def test_multi_range_update_is_bit_identical(env):
    """vb200_chol_update_ranges: one DMMA launch over several strided block-column ranges (the panels ONE rank owns
    in the panel-cyclic factorisation) against one vb200_chol_update per panel.  Same tiles, operands and order."""
    import ctypes

    vb, kernels, lib = env
    n, pt = 4300, 3
    e, nn, d = sample(n, seed=19)
    system = kernels.GramSystem(n).accumulate(0, e, nn, d, None, e, nn, 1e-5)
    gram = system.gram.clone()
    T = (n + 127) // 128
    panels = [(k0, min(pt, T - k0)) for k0 in range(0, T, pt)]
    forces = {}
    for mode in ("per_panel", "ranges"):
        system.gram.copy_(gram)
        ptr = system.gram.data_ptr()
        assert lib.vb200_chol_begin(ptr, system.stats.data_ptr(), n, n, 1e-3) == 0
        for p, (k0, w) in enumerate(panels):
            assert lib.vb200_chol_panel(ptr, n, k0, w) == 0
            later = panels[p + 1:]
            if mode == "per_panel":
                for q0, qw in later:
                    assert lib.vb200_chol_update(ptr, n, k0, w, q0, q0 + qw) == 0
            else:
                for phase in (0, 1, 2):  # three "ranks", each with its strided share of the later panels
                    mine = [(q0, q0 + qw) for q, (q0, qw) in enumerate(later) if q % 3 == phase]
                    lo = (ctypes.c_int32 * len(mine))(*[a for a, _ in mine])
                    hi = (ctypes.c_int32 * len(mine))(*[b for _, b in mine])
                    assert lib.vb200_chol_update_ranges(ptr, n, k0, w, len(mine), lo, hi) == 0
        out = np.empty(n)
        info = vb._cabi.FitInfo()
        assert lib.vb200_chol_finish(ptr, n, out.ctypes.data, ctypes.byref(info)) == 0
        forces[mode] = out
    np.testing.assert_array_equal(forces["ranges"], forces["per_panel"])
    # overlapping or unsorted ranges are refused
    system.gram.copy_(gram)
    assert lib.vb200_chol_begin(system.gram.data_ptr(), system.stats.data_ptr(), n, n, 1e-3) == 0
    lo, hi = (ctypes.c_int32 * 2)(6, 4), (ctypes.c_int32 * 2)(9, 7)
    assert lib.vb200_chol_update_ranges(system.gram.data_ptr(), n, 0, 3, 2, lo, hi) != 0
    out = np.empty(n)
    lib.vb200_chol_finish(system.gram.data_ptr(), n, out.ctypes.data, None)


@pytest.mark.parametrize("m,panel,min_rows", [(2500, 2, 4), (14_000, 8, 96)])
def test_lookahead_cholesky_is_bit_identical(env, m, panel, min_rows):
    """Single-GPU look-ahead: panel p+1 is factored on a few SMs beside the trailing update of panel p (two grids
    draining one dynamic tile queue).  Every tile sees the same operands in the same order -> the same bits."""
    vb, kernels, lib = env
    e, nn, d = sample(m, seed=31)
    saved = {k: vb._cabi.get_option(k) for k in ("gemm_variant", "chol_lookahead", "chol_panel_tiles", "chol_lookahead_min_rows")}
    forces = {}
    try:
        lib.vb200_set_option(b"chol_panel_tiles", float(panel))
        lib.vb200_set_option(b"chol_lookahead_min_rows", float(min_rows))
        for tag, variant, ahead in (("static", 3, 0), ("dynamic", 4, 0), ("ahead16", 4, 16), ("ahead40", 4, 40)):
            lib.vb200_set_option(b"gemm_variant", float(variant))
            lib.vb200_set_option(b"chol_lookahead", float(ahead))
            forces[tag] = kernels.fit_spline(e, nn, d, None, e, nn, 1e-5, 1e-4)
    finally:
        for k, v in saved.items():
            lib.vb200_set_option(k.encode(), v)
    for tag in ("dynamic", "ahead16", "ahead40"):
        np.testing.assert_array_equal(forces[tag], forces["static"], err_msg=tag)


def test_lookahead_batched_sweep_is_bit_identical(env):
    vb, kernels, lib = env
    e, nn, d = sample(1800, seed=33)
    dampings = [1e-6, 1e-4, 1e-2]
    system = kernels.GramSystem(e.size).accumulate(0, e, nn, d, None, e, nn, 0.0)
    plain = system.solve_multi(e.size, dampings)
    saved = {k: vb._cabi.get_option(k) for k in ("gemm_variant", "chol_lookahead", "chol_panel_tiles", "chol_lookahead_min_rows")}
    try:
        for k, v in (("gemm_variant", 4), ("chol_lookahead", 16), ("chol_panel_tiles", 2), ("chol_lookahead_min_rows", 4)):
            lib.vb200_set_option(k.encode(), float(v))
        ahead = system.solve_multi(e.size, dampings)
        lib.vb200_set_option(b"chol_lookahead", 0.0)
        same_panels = system.solve_multi(e.size, dampings)
    finally:
        for k, v in saved.items():
            lib.vb200_set_option(k.encode(), v)
    np.testing.assert_array_equal(ahead, same_panels)
    # panel width 2 vs 8 only re-associates the trailing sums: the well-conditioned member (damping 1e-2) agrees closely
    assert np.max(np.abs(ahead[2] - plain[2])) <= 1e-7 * np.max(np.abs(plain[2]))
