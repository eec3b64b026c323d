"""Small single-GPU program for compute-sanitizer: dense sweep, prefilter tier (tcgen05), list re-scoring, f64 recheck."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "multimoda-rs_b200")]
from multimodars import _native as nat
from scripts.quick_bench import contour

ctx = nat.Context(0)
rng = np.random.default_rng(3)
sizes = [520, 130, 2020]
for n in sizes:
    U = 2
    t = np.concatenate([contour(rng, n, rng.normal(0, .2)) for _ in range(U)])
    r = np.concatenate([contour(rng, n) for _ in range(U)])
    off = np.arange(U + 1) * n
    g = nat.make_grid(2.0, 90.0)
    a = ctx.sweep_batched(t, off, r, off, np.full((U, 2), 4.5), [g], mode=0, prefilter=1)
    b = ctx.sweep_batched(t, off, r, off, np.full((U, 2), 4.5), [g], mode=0, prefilter=2)
    assert (a["best_idx"] == b["best_idx"]).all() and (a["best_dist"] == b["best_dist"]).all(), n
    print(n, a["best_idx"], ctx.prefilter_info())
print("ok")
