import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "multimoda-rs_b200")]
import bench
from multimodars import _native as nat
ctx = nat.Context(0)
fr = bench.synthetic_pullback(400, 1000, 300)
th = 2.0 * np.pi * np.arange(20) / 20
cath = np.stack([4.5 + 0.5 * np.cos(th), 4.5 + 0.5 * np.sin(th)], 1)[::2]
pts = [np.concatenate([f[::2], cath]) - f.mean(axis=0) for f in fr]
tests, refs = pts[1:], pts[:-1]
U = len(tests); n = len(pts[0])
off = np.arange(U + 1) * n
g = nat.make_grid(1.0, 180.0)
res = ctx.sweep_batched(np.concatenate(tests), off, np.concatenate(refs), off, np.zeros((U, 2)), [g], mode=0, tie_margin=1e-9)
print("stage1 n_ties hist", np.bincount(res["n_ties"]), "n_shortlist hist", np.bincount(res["n_shortlist"])[:10])
bad = np.nonzero(res["n_ties"] != 1)[0]
for u in bad[:5]:
    idx, d = ctx.shortlist(int(u))
    o = np.argsort(idx)
    print("unit", u, "best", res["best_idx"][u], "shortlist idx", idx[o], "d", d[o])
grids = [nat.make_grid(0.1, 5.0, center=float(a), limes_deg=180.0) for a in res["best_angle"]]
ctx.sweep_regrid(grids, grid_of_unit=np.arange(U), tie_margin=1e-9)
ctx.sweep_run(); r2 = ctx.sweep_download()
print("stage2 n_ties hist", np.bincount(r2["n_ties"]), "n_shortlist hist", np.bincount(r2["n_shortlist"])[:10])
bad = np.nonzero(r2["n_ties"] != 1)[0]
for u in bad[:5]:
    idx, d = ctx.shortlist(int(u))
    o = np.argsort(idx)
    print("unit", u, "best", r2["best_idx"][u], "ncand", grids[u].n_cand, "shortlist idx", idx[o], "d", d[o], "diff", np.diff(d[o]))
grids3 = [nat.make_grid(0.01, 0.1, center=float(a), limes_deg=180.0) for a in r2["best_angle"]]
ctx.sweep_regrid(grids3, grid_of_unit=np.arange(U), tie_margin=1e-9)
ctx.sweep_run(); r3 = ctx.sweep_download()
print("stage3 n_ties hist", np.bincount(r3["n_ties"]), "n_shortlist hist", np.bincount(r3["n_shortlist"])[:12])
bad = np.nonzero(r3["n_ties"] != 1)[0]
for u in bad[:5]:
    idx, d = ctx.shortlist(int(u))
    o = np.argsort(idx)
    print("unit", u, "best", r3["best_idx"][u], "shortlist idx", idx[o], "d", d[o], "diff", np.diff(d[o]))
