"""Exact lower-bound pruning (mmrs_sweep_opts.prune, k_lb / k_lb_argmin in sweep_kernels.cuh): candidates whose lower
bound exceeds the exact distance of an already scored candidate are never scored. The bar is the dense path's:
selected index, wrapped angle, f64 distance, FP32 minimum and tie count identical to the dense sweep and to the CPU
oracle; additionally every pruned candidate's bound must really be a lower bound of its exact FP32 distance."""
import numpy as np
import pytest

import multimodars as mm
from multimodars import _native as nat
from oracle import oracle_py as ora
from tests import golden_io as gio
from tests.test_sweep_gpu import make_units

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = nat.Context(0)
    yield c
    c.close()


def both(ctx, sizes, step, rng_deg, mode, seed, **kw):
    rng = np.random.default_rng(seed)
    tests, refs, cents, txy, toff, rxy, roff = make_units(rng, sizes)
    g = nat.make_grid(step, rng_deg)
    dense = ctx.sweep_batched(txy, toff, rxy, roff, cents, [g], mode=mode, prune=-1, **kw)
    d32 = [ctx.dist32(u, g.n_cand) for u in range(len(sizes))]
    assert not ctx.prefilter_info()["ran"]
    pr = ctx.sweep_batched(txy, toff, rxy, roff, cents, [g], mode=mode, prune=1, **kw)
    info = ctx.prefilter_info()
    lb = [ctx.dist32(u, g.n_cand) for u in range(len(sizes))]
    return tests, refs, cents, g, dense, d32, pr, lb, info


@pytest.mark.parametrize("mode", [0, 1])
def test_pruned_sweep_is_identical_and_bounds_hold(ctx, mode):
    sizes = [(520, 520), (510, 510), (128, 300), (130, 200), (600, 555), (1000, 1024), (2020, 2020), (700, 2020)]
    tests, refs, cents, g, dense, d32, pr, lb, info = both(ctx, sizes, 0.5, 90.0, mode, 31 + mode, tie_margin=1e-9)
    assert info["kind"] == "lower-bound pruning" and g.n_cand >= 256
    for name in ("best_idx", "best_angle", "best_dist", "best_dist_f32", "n_ties", "n_shortlist", "flags"):
        assert (dense[name] == pr[name]).all(), name
    for u, (t, r, c) in enumerate(zip(tests, refs, cents)):
        o = ora.sweep(t, r, c, mode, 0.5, 90.0)
        assert pr["best_idx"][u] == o["index"] and pr["best_dist"][u] == o["cost"] and pr["best_angle"][u] == o["angle"]
        # after a pruned run dist32 holds the exact FP32 distance of scored candidates and the BOUND of the others
        scored = lb[u] == d32[u]
        assert scored[o["index"]]
        slack = 4e-6 * max(np.abs(t - c).max(), np.abs(r - c).max())
        assert (lb[u] <= d32[u] * (1 + 4e-6) + slack).all(), u
        # ... and no unscored candidate could have won
        assert (lb[u][~scored] > d32[u].min()).all()
    assert 0 < info["rescored"] < 0.8 * len(sizes) * g.n_cand, info


def test_pruned_fine_grid_and_plateau(ctx):
    sizes = [(520, 520), (500, 520)]
    tests, refs, cents, g, dense, d32, pr, lb, info = both(ctx, sizes, 0.01, 180.0, 0, 2)
    assert g.n_cand == 36000 and info["kind"] == "lower-bound pruning"
    for name in ("best_idx", "best_dist", "best_dist_f32"):
        assert (dense[name] == pr[name]).all(), name
    for u, (t, r, c) in enumerate(zip(tests, refs, cents)):
        o = ora.sweep(t, r, c, 0, 0.01, 180.0, threads=8)
        assert pr["best_idx"][u] == o["index"] and pr["best_dist"][u] == o["cost"]
    assert info["rescored"] < 0.5 * 2 * g.n_cand, info
    # a circle against itself: every candidate has the same cost, every bound survives, leftmost index wins
    n = 256
    phi = np.linspace(0, 2 * np.pi, n, endpoint=False)
    circ = np.stack([2.0 * np.cos(phi) + 4.5, 2.0 * np.sin(phi) + 4.5], 1)
    g = nat.make_grid(0.5, 90.0)
    a = ctx.sweep_batched(circ, [0, n], circ, [0, n], [[4.5, 4.5]], [g], mode=0, prune=-1)
    b = ctx.sweep_batched(circ, [0, n], circ, [0, n], [[4.5, 4.5]], [g], mode=0, prune=1)
    assert ctx.prefilter_info()["kind"] == "lower-bound pruning"
    o = ora.sweep(circ, circ, (4.5, 4.5), 0, 0.5, 90.0)
    assert a["best_idx"][0] == b["best_idx"][0] == o["index"] and a["best_dist"][0] == b["best_dist"][0] == o["cost"]


def test_pruning_does_not_apply_to_small_units_or_short_grids(ctx):
    rng = np.random.default_rng(1)
    tests, refs, cents, txy, toff, rxy, roff = make_units(rng, [(520, 520), (64, 64)])
    ctx.sweep_batched(txy, toff, rxy, roff, cents, [nat.make_grid(0.5, 90.0)], mode=0, prune=1)
    assert not ctx.prefilter_info()["ran"]                     # a 64-point unit: dense sweep
    tests, refs, cents, txy, toff, rxy, roff = make_units(rng, [(520, 520)] * 3)
    ctx.sweep_batched(txy, toff, rxy, roff, cents, [nat.make_grid(1.0, 90.0)], mode=0, prune=1)
    assert not ctx.prefilter_info()["ran"]                     # 181 candidates per unit: not worth two extra passes


def test_process_level_pruning_matches_golden():
    """from_array_full(bruteforce=True) with the context-wide pruning default: logs and geometries bit-identical to
    the committed config-1 goldens (0.5 deg over +-90 deg = 361 candidates per unit)."""
    pack, gold = gio.inputs(), gio.oracle_outputs()
    full = [("rest", True), ("rest", False), ("stress", True), ("stress", False)]
    ins = [gio.py_input(mm, pack, n, d, f"{n}_{'dia' if d else 'sys'}") for n, d in full]
    ctx = mm.get_context()
    ctx.set_prune(True)
    try:
        ab, cd, ac, bd, logs = mm.from_array_full(*ins, sample_size=500, write_obj=False, postprocessing=False,
                                                  step_rotation_deg=0.5, range_rotation_deg=90.0, bruteforce=True,
                                                  smooth=False)
    finally:
        ctx.set_prune(False)
    for i in range(4):
        assert np.array_equal(np.array(logs[i], dtype=np.float64).reshape(-1, 7), gold[f"cfg1_brute0p5_logs_{i}"])
    outs = [ab.geom_a, ab.geom_b, cd.geom_a, cd.geom_b, ac.geom_a, ac.geom_b, bd.geom_a, bd.geom_b]
    for i, g in enumerate(outs):
        assert gio.sha(g.to_blob()) == str(gold[f"cfg1_brute0p5_out_sha_{i}"])
