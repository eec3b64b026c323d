"""Device-side timing of the tensor-core prefilter path against the dense FP32 sweep on config-2 / config-3 /
config-4 shaped batches (the same synthetic contours as scripts/quick_bench.py)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "multimoda-rs_b200")]
from multimodars import _native as nat
from scripts.quick_bench import contour


def run(ctx, U, N, step, rng_deg, reps=3):
    rng = np.random.default_rng(0)
    t = np.concatenate([contour(rng, N, rng.normal(0, .2)) for _ in range(U)])
    r = np.concatenate([contour(rng, N) for _ in range(U)])
    off = np.arange(U + 1) * N
    g = nat.make_grid(step, rng_deg)
    out = {}
    for name, pf, prune in (("dense", 1, -1), ("tc", 2, -1), ("prune", 1, 1)):
        if name == "tc" and "--no-tc" in sys.argv:
            continue
        ctx.sweep_upload(t, off, r, off, np.full((U, 2), 4.5), [g], mode=0, prefilter=pf, prune=prune)
        best = None
        for _ in range(reps):
            ctx.sweep_run(); res = ctx.sweep_download(); tm = ctx.timings()
            best = tm if best is None or tm["total_ms"] < best["total_ms"] else best
        out[name] = (best, res, ctx.prefilter_info())
    evals = U * g.n_cand
    d, tcr = out["dense"], out.get("tc", out["prune"])
    p = out["prune"]
    same_p = bool((d[1]["best_idx"] == p[1]["best_idx"]).all() and (d[1]["best_dist"] == p[1]["best_dist"]).all())
    print(f"U={U} N={N} C={g.n_cand}: dense {d[0]['total_ms']:.2f} ms | pruned total {p[0]['total_ms']:.2f} ms "
          f"(bounds {p[2]['tc_ms']:.2f} ms, survivors {p[2]['rescore_ms']:.2f} ms, {p[2]['rescored']} scored = "
          f"{p[2]['rescored'] / (U * g.n_cand):.3f} of all) speed-up {d[0]['total_ms'] / p[0]['total_ms']:.2f}x identical={same_p}", flush=True)
    same = bool((d[1]["best_idx"] == tcr[1]["best_idx"]).all() and (d[1]["best_dist"] == tcr[1]["best_dist"]).all())
    info = tcr[2]
    print(f"U={U} N={N} C={g.n_cand}: dense {d[0]['total_ms']:.2f} ms ({evals / d[0]['total_ms'] * 1e3:.4g} evals/s) | "
          f"prefilter total {tcr[0]['total_ms']:.2f} ms ({evals / tcr[0]['total_ms'] * 1e3:.4g} evals/s; K1t {info['tc_ms']:.2f} ms, "
          f"rescoring {info['rescore_ms']:.2f} ms, {info['rescored']} rescored = {info['rescored'] / evals:.2e} of all, "
          f"max err {info['max_err']:.2e} of window {info['window']:.1e}) speed-up {d[0]['total_ms'] / tcr[0]['total_ms']:.2f}x "
          f"identical={same}", flush=True)


if __name__ == "__main__":
    ctx = nat.Context(0)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    run(ctx, 8, 520, 0.05, 180.0)
    run(ctx, 40, 520, 0.01, 180.0)
    if which == "all":
        run(ctx, 398, 520, 0.01, 180.0)
        run(ctx, 40, 510, 0.01, 180.0)
        run(ctx, 8, 2020, 0.05, 180.0)
        run(ctx, 1596, 510, 1.0, 180.0)
