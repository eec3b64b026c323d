"""One-off full-size parity check on the GPU box (too slow for the test-suite: the CPU oracle needs minutes).

config 2 (BASELINE configs[1]): from_array_singlepair on 2 x 200 frames x 500 points, brute force 0.01 deg over
+-180 deg -> the per-frame logs of the first K frames of both pullbacks must be bit-identical to the oracle's
frame chain on those frames (the chain loop of frame i depends on frames 0..i only).
config 4 shape: a 4-frame pullback with 2000-point contours, brute force 0.005 deg over +-180 deg (72 000 candidates,
N = M = 2020, 4 register chunks)."""
import json, os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "multimoda-rs_b200")]
import bench
import multimodars as mm
from multimodars import _native as nat
from oracle import oracle_py as ora

K = int(sys.argv[1]) if len(sys.argv) > 1 else 40
cores = os.cpu_count() or 1
out = {}

def rows(seed, n_frames, n_points):
    fr = bench.synthetic_pullback(n_frames, n_points, seed)
    z = 0.5 * (n_frames - 1 - np.arange(n_frames))
    a = np.concatenate([np.column_stack([np.full(n_points, float(i)), f, np.full(n_points, z[i])]) for i, f in enumerate(fr)])
    last = a[a[:, 0] == n_frames - 1][0]
    return a, np.array([n_frames - 1, last[1] + 0.1, last[2], last[3]])

def truncated(blob, k):
    fr = ora.decode_geometry(blob)[:k]
    return ora.encode_geometry(fr)

# ---- config 2 -------------------------------------------------------------------------------------------
ins, blobs = [], []
for s, dia in ((20261018, True), (20261019, False)):
    a, rp = rows(s, 200, 500)
    ins.append(mm.numpy_to_inputdata(a, rp, dia, label="dia" if dia else "sys"))
    blobs.append(nat.geometry_from_arrays(a, rp, diastole=dia, label="x"))
t0 = time.perf_counter()
pair, logs = mm.from_array_singlepair(*ins, step_rotation_deg=0.01, range_rotation_deg=180.0, sample_size=500,
                                      write_obj=False, bruteforce=True, smooth=True, postprocessing=False)
gpu_s = time.perf_counter() - t0
st = mm.get_context().process_stats()
ok = True
t0 = time.perf_counter()
for p in range(2):
    _, want, _ = ora.align_within(truncated(blobs[p], K), 0.01, 180.0, False, True, 500, threads=cores, post_steps=False)
    got = np.array(logs[p], dtype=np.float64)[:K - 1]
    same = np.array_equal(got, want)
    ok &= same
    print(f"config2 pullback {p}: first {K - 1} frame pairs bit-identical: {same}", flush=True)
out["config2"] = dict(frames_checked_per_pullback=K - 1, bit_identical=bool(ok), gpu_wall_s=gpu_s, oracle_wall_s=time.perf_counter() - t0,
                      stats=st)

# ---- config 4 shape -----------------------------------------------------------------------------------------
a, rp = rows(777, 4, 2000)
inp = mm.numpy_to_inputdata(a, rp, True, label="oct")
blob = nat.geometry_from_arrays(a, rp, diastole=True, label="x")
t0 = time.perf_counter()
g, lg = mm.from_array_single(inp, step_rotation_deg=0.005, range_rotation_deg=180.0, sample_size=2000, bruteforce=True, smooth=False)
gpu_s = time.perf_counter() - t0
t0 = time.perf_counter()
_, want, _ = ora.align_within(blob, 0.005, 180.0, False, True, 2000, threads=cores, post_steps=False)
same = np.array_equal(np.array(lg, dtype=np.float64), want)
print(f"config4 shape (N=M=2020, 72000 candidates, 3 frame pairs) bit-identical: {same}", flush=True)
out["config4_shape"] = dict(bit_identical=bool(same), gpu_wall_s=gpu_s, oracle_wall_s=time.perf_counter() - t0, cores=cores,
                            plan=mm.get_context().plan())
print(json.dumps(out))
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "validate_full_size.json").write_text(json.dumps(out, indent=1))
