"""The expanded-form tier K1x (k_sweep<.., XF>, mmrs_sweep_opts.prefilter = 3): every candidate is scored with
|a - b|^2 = |b|^2 - 2 a.b + |a|^2 (6 instead of 8 packed FP32 instructions per 2 x 2 block), the candidates inside the
tier's error window of the minimum are re-scored by the direct-form kernel, and K2 / K3 / K4 run unchanged on exact
FP32 values. Bars: the selected candidate, its wrapped angle, its f64 distance, the FP32 minimum and the tie count are
BIT-IDENTICAL to the dense direct path and to the CPU oracle; the tier's error on d^2 stays below the proven bound
13 u Rn^2 <= 1.6e-6 Rmax^2 for EVERY candidate (so the window of 8e-6 Rmax^2 can never filter the f64 arg-min out)."""
import numpy as np
import pytest

from multimodars import _native as nat
from oracle import oracle_py as ora
from tests.test_sweep_gpu import make_units

pytestmark = pytest.mark.gpu

WINDOW = 8e-6          # default tier-1 window in units of Rmax^2
BOUND = 1.6e-6         # proven bound of the tier's absolute error on d^2, in units of Rmax^2 (sweep_kernels.cuh)


@pytest.fixture(scope="module")
def ctx():
    c = nat.Context(0)
    yield c
    c.close()


def run_both(ctx, sizes, step, rng_deg, mode, seed, centre=(4.5, 4.5)):
    rng = np.random.default_rng(seed)
    tests, refs, cents, txy, toff, rxy, roff = make_units(rng, sizes, centre=centre)
    g = nat.make_grid(step, rng_deg)
    dense = ctx.sweep_batched(txy, toff, rxy, roff, cents, [g], mode=mode, prefilter=1, tie_margin=1e-9)
    d32 = [ctx.dist32(u, g.n_cand).astype(np.float64) for u in range(len(sizes))]
    assert not ctx.prefilter_info()["ran"]
    xf = ctx.sweep_batched(txy, toff, rxy, roff, cents, [g], mode=mode, prefilter=3, tie_margin=1e-9)
    info = ctx.prefilter_info()
    dxf = [ctx.dist32(u, g.n_cand).astype(np.float64) for u in range(len(sizes))]
    return tests, refs, cents, g, dense, d32, xf, dxf, info


def check(tests, refs, cents, sizes, dense, d32, xf, dxf, info, oracle=None):
    assert info["kind"] == "expanded-form tier" and info["rescored"] >= len(sizes)
    for f in ("best_idx", "best_angle", "best_dist", "best_dist_f32", "n_ties", "flags"):
        assert np.array_equal(dense[f], xf[f]), f
    for u, (t, r, c) in enumerate(zip(tests, refs, cents)):
        if oracle is not None:
            o = oracle(t, r, c)
            assert xf["best_idx"][u] == o["index"] and xf["best_angle"][u] == o["angle"] and xf["best_dist"][u] == o["cost"]
        rmax = max(np.abs(t - c).max(), np.abs(r - c).max())
        err = np.abs(dxf[u] ** 2 - d32[u] ** 2).max() / rmax ** 2
        assert err <= BOUND, (u, sizes[u], err)
    assert info["max_err"] <= BOUND and info["window"] == WINDOW


@pytest.mark.parametrize("mode", [0, 1])
def test_expanded_tier_selection_is_bit_identical(ctx, mode):
    """Every kernel flavour: exact tiling (520), padded even / odd register tiles, chunked test sets, tiny sets."""
    sizes = [(520, 520), (510, 510), (64, 300), (130, 64), (33, 257), (129, 255), (600, 555), (1000, 1024), (6, 6), (543, 512)]
    tests, refs, cents, g, dense, d32, xf, dxf, info = run_both(ctx, sizes, 0.5, 90.0, mode, 41 + mode)
    check(tests, refs, cents, sizes, dense, d32, xf, dxf, info, lambda t, r, c: ora.sweep(t, r, c, mode, 0.5, 90.0))
    assert info["rescored"] <= 0.2 * len(sizes) * g.n_cand, info


def test_expanded_tier_oct_resolution_and_fine_grid(ctx):
    sizes = [(2020, 2020), (2020, 1999), (700, 2020)]
    tests, refs, cents, g, dense, d32, xf, dxf, info = run_both(ctx, sizes, 0.25, 90.0, 0, 9)
    check(tests, refs, cents, sizes, dense, d32, xf, dxf, info, lambda t, r, c: ora.sweep(t, r, c, 0, 0.25, 90.0, threads=8))
    sizes = [(520, 520), (500, 520)]
    tests, refs, cents, g, dense, d32, xf, dxf, info = run_both(ctx, sizes, 0.01, 180.0, 0, 2)
    assert g.n_cand == 36000
    check(tests, refs, cents, sizes, dense, d32, xf, dxf, info, lambda t, r, c: ora.sweep(t, r, c, 0, 0.01, 180.0, threads=8))
    assert info["rescored"] < 0.02 * 2 * 36000, info


def test_expanded_tier_off_centre_and_plateau(ctx):
    """Rotation centres far from the contours make the point norms (and with them the cancellation) large: the error
    bound scales with Rmax^2 and must still hold. A circle on a circle is a near-plateau: a large part of the grid is re-scored."""
    sizes = [(520, 520), (300, 310)]
    tests, refs, cents, g, dense, d32, xf, dxf, info = run_both(ctx, sizes, 0.5, 30.0, 1, 3, centre=(-2.0, 11.0))
    check(tests, refs, cents, sizes, dense, d32, xf, dxf, info, lambda t, r, c: ora.sweep(t, r, c, 1, 0.5, 30.0))
    n = 256
    phi = np.linspace(0, 2 * np.pi, n, endpoint=False)
    circ = np.stack([2.0 * np.cos(phi) + 4.5, 2.0 * np.sin(phi) + 4.5], 1)
    g = nat.make_grid(1.0, 90.0)
    res = ctx.sweep_batched(circ, [0, n], circ, [0, n], [[4.5, 4.5]], [g], mode=0, prefilter=3)
    o = ora.sweep(circ, circ, (4.5, 4.5), 0, 1.0, 90.0)
    assert res["best_idx"][0] == o["index"] and res["best_dist"][0] == o["cost"]
    assert ctx.prefilter_info()["rescored"] >= 1
