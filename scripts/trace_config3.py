"""Phase trace (MMRS_TRACE=1) of BASELINE config 3 through mmrs_process_cases: double pair, 4 x 400 frames x 1 000 points,
coarse-to-fine 0.01 deg over +-180 — a host-dominated call (37 ms of sweeps in 160)."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "multimoda-rs_b200"), str(ROOT / "scripts")]
os.environ["MMRS_TRACE"] = "1"
import numpy as np
import config_bench as cb
import multimodars as mm
from multimodars import _native as nat

ctx = mm.get_context()
blobs = []
for k, dia in enumerate((True, False, True, False)):
    a, rp = cb.rows(300 + k, 400, 1000)
    blobs.append(nat.geometry_from_arrays(a, rp, diastole=dia, label=f"p{k}"))
nat.process_cases(ctx, 3, blobs, 0.01, 180.0, 500, True, False)
for rep in range(2):
    print(f"---- rep {rep}", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    nat.process_cases(ctx, 3, blobs, 0.01, 180.0, 500, True, False)
    print(f"python: process_cases {1e3*(time.perf_counter()-t0):.2f} ms; stats {ctx.process_stats()}", file=sys.stderr, flush=True)
