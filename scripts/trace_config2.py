"""Phase trace (MMRS_TRACE=1) of BASELINE config 2 through the public API: from_array_singlepair on 2 x 200 frames x
500 points, brute force 0.01 deg over +-180 (36 000 candidates) — where the wall time that is not sweep goes."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "multimoda-rs_b200")]
os.environ["MMRS_TRACE"] = "1"
import numpy as np
import bench
import multimodars as mm
from multimodars import _native as nat

step = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
t0 = time.perf_counter()
rows = [bench.pullback_rows(bench.SEED + k, bench.N_FRAMES, bench.N_POINTS) for k in range(2)]
t1 = time.perf_counter()
ins = [mm.numpy_to_inputdata(*rows[k], k == 0, label="dia" if k == 0 else "sys") for k in range(2)]
t2 = time.perf_counter()
print(f"python: synthetic rows {1e3*(t1-t0):.1f} ms, numpy_to_inputdata x2 {1e3*(t2-t1):.1f} ms", file=sys.stderr)
kw = dict(step_rotation_deg=step, range_rotation_deg=180.0, sample_size=500, write_obj=False, bruteforce=True, smooth=True,
          postprocessing=False)
mm.from_array_singlepair(*ins, **kw)
for rep in range(2):
    print(f"---- rep {rep}", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    mm.from_array_singlepair(*ins, **kw)
    print(f"python: whole call {1e3*(time.perf_counter()-t0):.2f} ms; stats {mm.get_context().process_stats()}", file=sys.stderr, flush=True)
