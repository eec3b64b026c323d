"""Phase trace (MMRS_TRACE=1) of a config-4-shaped full-mode call: from_array_full on 4 synthetic pullbacks of
`frames` x `points`, brute force at `step` degrees over +-180. Python-side laps (ingest / native call / result objects)
next to the library's own laps on stderr.   python scripts/trace_full_mode.py [frames] [points] [step_deg]"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "multimoda-rs_b200")]
os.environ["MMRS_TRACE"] = "1"
import numpy as np
import bench
import multimodars as mm
from multimodars import _native as nat, _processing as P

F = int(sys.argv[1]) if len(sys.argv) > 1 else 125
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
STEP = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05


def rows(seed):
    fr = bench.synthetic_pullback(F, NP, seed)
    z = 0.5 * (F - 1 - np.arange(F))
    a = np.concatenate([np.column_stack([np.full(NP, float(i)), f, np.full(NP, z[i])]) for i, f in enumerate(fr)])
    last = a[a[:, 0] == F - 1][0]
    return a, np.array([F - 1, last[1] + 0.1, last[2], last[3]])


arrs = [rows(20261018 + k) for k in range(4)]
w, rp = arrs[0][0][: 4 * NP], arrs[0][1].copy()
rp[0] = 3
mm.from_array_single(mm.numpy_to_inputdata(w, np.array([3, w[3 * NP][1] + 0.1, w[3 * NP][2], w[3 * NP][3]]), True, label="w"), sample_size=64)
for rep in range(2):
    print(f"---- rep {rep}: 4 x {F} frames x {NP} points, step {STEP}", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    ins = [mm.numpy_to_inputdata(a, r, k % 2 == 0, label=f"phase{k}") for k, (a, r) in enumerate(arrs)]
    t1 = time.perf_counter()
    blobs = [P._blob_from_input(i, (4.5, 4.5), 0.5, 20) for i in ins]
    t2 = time.perf_counter()
    out, logs, _ = nat.process_cases(mm.get_context(), 4, blobs, STEP, 180.0, NP, True, True, False)
    t3 = time.perf_counter()
    geoms = [mm.PyGeometry.from_blob(o, "x") for o in out]
    t4 = time.perf_counter()
    print(f"python: numpy_to_inputdata {1e3*(t1-t0):.1f} ms, geometry_from_arrays x4 {1e3*(t2-t1):.1f} ms, "
          f"process_cases {1e3*(t3-t2):.1f} ms, from_blob x{len(out)} {1e3*(t4-t3):.1f} ms; stats {mm.get_context().process_stats()}",
          file=sys.stderr, flush=True)
