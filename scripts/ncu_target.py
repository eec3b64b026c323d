"""Short single-GPU program for ncu captures: one config-2 shaped sweep launch (U units x 36000 candidates)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "multimoda-rs_b200")]
import bench
from multimodars import _native as nat

U = int(sys.argv[1]) if len(sys.argv) > 1 else 40
txy, toff, rxy, roff, cen, _, n = bench.make_units(20261018)
ctx = nat.Context(0)
g = nat.make_grid(bench.STEP_DEG, bench.RANGE_DEG)
ctx.sweep_upload(txy[:U * n], toff[:U + 1], rxy[:U * n], roff[:U + 1], cen[:U], [g], mode=0)
for _ in range(2):
    ctx.sweep_run()
    res = ctx.sweep_download()
print("units", U, "timings", ctx.timings(), "mean shortlist", res["n_shortlist"].mean())
