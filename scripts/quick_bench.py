"""Quick device-side timing of the sweep kernel on config-2 / config-4 shaped batches."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "multimoda-rs_b200")]
from multimodars import _native as nat

def contour(rng, n, rot=0.0):
    phi = np.linspace(0, 2 * np.pi, n, endpoint=False)
    r = rng.uniform(1.5, 3) * (1 + rng.uniform(0.05, 0.35) * np.cos(2 * phi) + 0.03 * np.cos(3 * phi + 0.4))
    return np.stack([r * np.cos(phi + rot) + 4.5, r * np.sin(phi + rot) + 4.5], 1) + rng.normal(0, 0.005, (n, 2))

def run(ctx, U, N, step, rng_deg, reps=3):
    rng = np.random.default_rng(0)
    t = np.concatenate([contour(rng, N, rng.normal(0, .2)) for _ in range(U)])
    r = np.concatenate([contour(rng, N) for _ in range(U)])
    off = np.arange(U + 1) * N
    g = nat.make_grid(step, rng_deg)
    ctx.sweep_upload(t, off, r, off, np.full((U, 2), 4.5), [g], mode=0)
    best = None
    for _ in range(reps):
        ctx.sweep_run(); res = ctx.sweep_download(); tm = ctx.timings()
        best = tm if best is None or tm["sweep_ms"] < best["sweep_ms"] else best
    evals = U * g.n_cand
    F = 10.0 * N * N + 6 * N
    print(f"U={U} N={N} C={g.n_cand}: sweep {best['sweep_ms']:.2f} ms, shortlist {best['shortlist_ms']:.3f}, recheck {best['recheck_ms']:.3f}; "
          f"{evals / best['sweep_ms'] * 1e3:.4g} evals/s, {evals * F / best['sweep_ms'] * 1e3 / 1e12:.2f} algorithmic TFLOP/s; "
          f"mean shortlist {res['n_shortlist'].mean():.2f} max {res['n_shortlist'].max()}", flush=True)

if __name__ == "__main__":
    ctx = nat.Context(0)
    print("fp32 probe TFLOP/s:", ctx.fp32_probe(4096), flush=True)
    run(ctx, 40, 520, 0.01, 180.0)
    run(ctx, 398, 520, 0.01, 180.0)
    run(ctx, 40, 510, 0.01, 180.0)
    run(ctx, 8, 2020, 0.05, 180.0)
    run(ctx, 74, 2020, 0.05, 180.0)
    run(ctx, 1596, 510, 1.0, 180.0)
    print("fp32 probe TFLOP/s:", ctx.fp32_probe(4096), flush=True)
