"""The tensor-core prefilter tier (k_tc_sweep, tc_kernels.cuh): tcgen05 bf16x3 scoring of every candidate ->
exact FP32 re-scoring of the candidates inside the prefilter's window -> FP32 window -> reference f64 -> arg-min.
Bar: the selected candidate, its wrapped angle and its f64 distance are BIT-IDENTICAL to the CPU oracle (and
therefore to the dense FP32 path); the prefilter's error on d^2 stays inside half its window for EVERY candidate
(the soundness condition: the f64 arg-min can then never be filtered out)."""
import numpy as np
import pytest

from multimodars import _native as nat
from oracle import oracle_py as ora
from tests.test_sweep_gpu import make_units

pytestmark = pytest.mark.gpu

WINDOW = 4e-6  # mmrs_sweep_opts.prefilter_abs default, in units of Rmax^2


@pytest.fixture(scope="module")
def ctx():
    c = nat.Context(0)
    yield c
    c.close()


def run_both(ctx, sizes, step, rng_deg, mode, seed):
    rng = np.random.default_rng(seed)
    tests, refs, cents, txy, toff, rxy, roff = make_units(rng, sizes)
    g = nat.make_grid(step, rng_deg)
    dense = ctx.sweep_batched(txy, toff, rxy, roff, cents, [g], mode=mode, prefilter=1)
    d32 = [ctx.dist32(u, g.n_cand).astype(np.float64) for u in range(len(sizes))]
    assert not ctx.prefilter_info()["ran"]
    tc = ctx.sweep_batched(txy, toff, rxy, roff, cents, [g], mode=mode, prefilter=2)
    info = ctx.prefilter_info()
    dtc = [ctx.dist32(u, g.n_cand).astype(np.float64) for u in range(len(sizes))]
    return tests, refs, cents, g, dense, d32, tc, dtc, info


@pytest.mark.parametrize("mode", [0, 1])
def test_prefilter_selection_is_bit_identical(ctx, mode):
    sizes = [(520, 520), (510, 510), (64, 300), (130, 64), (128, 128), (129, 255), (600, 555), (1000, 1024)]
    tests, refs, cents, g, dense, d32, tc, dtc, info = run_both(ctx, sizes, 0.5, 90.0, mode, 21 + mode)
    assert info["ran"] and info["rescored"] >= len(sizes)
    for u, (t, r, c) in enumerate(zip(tests, refs, cents)):
        o = ora.sweep(t, r, c, mode, 0.5, 90.0)
        for res in (dense, tc):
            assert res["best_idx"][u] == o["index"], (u, res[u], o["index"])
            assert res["best_angle"][u] == o["angle"] and res["best_dist"][u] == o["cost"]
        assert tc["best_dist_f32"][u] == dense["best_dist_f32"][u]   # the same exact FP32 minimum
        rmax = max(np.abs(t - c).max(), np.abs(r - c).max())
        err = np.abs(dtc[u] ** 2 - d32[u] ** 2).max() / rmax ** 2
        assert err <= 0.5 * WINDOW, (u, sizes[u], err)
    assert info["max_err"] <= 0.5 * WINDOW
    # the window is narrow: a handful of candidates per unit are re-scored in FP32, not the grid
    assert info["rescored"] <= 0.1 * len(sizes) * g.n_cand, info


def test_prefilter_oct_resolution_and_fine_grid(ctx):
    """Config-4 shaped units (N = M = 2020: 16 row tiles, single operand buffer) and a 36 000-candidate grid."""
    sizes = [(2020, 2020), (2020, 1999), (700, 2020)]
    tests, refs, cents, g, dense, d32, tc, dtc, info = run_both(ctx, sizes, 0.25, 90.0, 0, 9)
    assert info["ran"]
    for u, (t, r, c) in enumerate(zip(tests, refs, cents)):
        o = ora.sweep(t, r, c, 0, 0.25, 90.0, threads=8)
        assert tc["best_idx"][u] == o["index"] and tc["best_dist"][u] == o["cost"] and tc["best_angle"][u] == o["angle"]
        rmax = max(np.abs(t - c).max(), np.abs(r - c).max())
        assert np.abs(dtc[u] ** 2 - d32[u] ** 2).max() / rmax ** 2 <= 0.5 * WINDOW
    sizes = [(520, 520), (500, 520)]
    tests, refs, cents, g, dense, d32, tc, dtc, info = run_both(ctx, sizes, 0.01, 180.0, 0, 2)
    assert g.n_cand == 36000 and info["ran"]
    for u, (t, r, c) in enumerate(zip(tests, refs, cents)):
        o = ora.sweep(t, r, c, 0, 0.01, 180.0, threads=8)
        assert tc["best_idx"][u] == o["index"] and tc["best_dist"][u] == o["cost"]
        assert dense["best_idx"][u] == o["index"]
        rmax = max(np.abs(t - c).max(), np.abs(r - c).max())
        assert np.abs(dtc[u] ** 2 - d32[u] ** 2).max() / rmax ** 2 <= 0.5 * WINDOW
    assert info["rescored"] < 2000, info


def test_prefilter_plateau_and_regrid(ctx):
    """A circle against a circle is a plateau (every candidate inside every window): the tier-1 list takes the whole
    grid, the result is still the reference's leftmost arg-min. Then the coarse-to-fine regrid path."""
    n = 256
    phi = np.linspace(0, 2 * np.pi, n, endpoint=False)
    circ = np.stack([2.0 * np.cos(phi) + 4.5, 2.0 * np.sin(phi) + 4.5], 1)
    g = nat.make_grid(1.0, 90.0)
    res = ctx.sweep_batched(circ, [0, n], circ, [0, n], [[4.5, 4.5]], [g], mode=0, prefilter=2)
    o = ora.sweep(circ, circ, (4.5, 4.5), 0, 1.0, 90.0)
    assert res["best_idx"][0] == o["index"] and res["best_dist"][0] == o["cost"]
    rng = np.random.default_rng(5)
    tests, refs, cents, txy, toff, rxy, roff = make_units(rng, [(520, 520)] * 6)
    g = nat.make_grid(1.0, 90.0)
    ctx.sweep_upload(txy, toff, rxy, roff, cents, [g], mode=0, prefilter=2)
    ctx.sweep_run()
    coarse = ctx.sweep_download()
    assert ctx.prefilter_info()["ran"]
    grids = [nat.make_grid(0.5, 5.0, center=float(c), limes_deg=90.0) for c in coarse["best_angle"]]
    ctx.sweep_regrid(grids, np.arange(6))
    ctx.sweep_run()
    fine = ctx.sweep_download()
    for u, (t, r, c) in enumerate(zip(tests, refs, cents)):
        o1 = ora.sweep(t, r, c, 0, 1.0, 90.0)
        o2 = ora.sweep(t, r, c, 0, 0.5, 5.0, center=coarse["best_angle"][u], limes_deg=90.0)
        assert coarse["best_idx"][u] == o1["index"] and coarse["best_dist"][u] == o1["cost"]
        assert fine["best_idx"][u] == o2["index"] and fine["best_dist"][u] == o2["cost"]


def test_prefilter_required_rejects_unfit_shapes(ctx):
    rng = np.random.default_rng(1)
    tests, refs, cents, txy, toff, rxy, roff = make_units(rng, [(520, 520), (6, 6)])
    g = nat.make_grid(0.5, 90.0)
    with pytest.raises(nat.MmrsError, match="prefilter required"):
        ctx.sweep_batched(txy, toff, rxy, roff, cents, [g], mode=0, prefilter=2)
    res = ctx.sweep_batched(txy, toff, rxy, roff, cents, [g], mode=0)   # auto: the dense sweep
    assert not ctx.prefilter_info()["ran"] and (res["best_idx"] >= 0).all()
