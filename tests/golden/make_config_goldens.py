"""Generates tests/golden/config_inputs.npz and tests/golden/config_goldens.npz:

    python tests/golden/make_config_goldens.py            # ~20 min of CPU (config 4 is 72 000 candidates x 2 020^2 pairs)

Small-frame-count instances of BASELINE.json configs 2-5 at their FULL search settings (candidate counts and contour
sizes as named there), run through the CPU ORACLE (oracle/, the f64 restatement pinned on the reference's Rust KATs;
the reference itself cannot be built here). tests/test_configs_gpu.py replays the same inputs through the public entry
points on the GPU and requires bit-identical logs and output geometries, so the GPU box never spends minutes on the CPU.

Inputs are synthetic pullbacks (bench.synthetic_pullback) quantised to a 1e-4 mm grid and stored as
int32 so that every machine rebuilds exactly the same doubles (x = k / 10000.0)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT)]
import bench  # noqa: E402
from oracle import oracle_py as ora  # noqa: E402
from tests import golden_io as gio  # noqa: E402

OUT = Path(__file__).resolve().parent
Q = 10000.0

# name -> (mode, pullbacks, frames, points, seed0, oracle/product arguments)
CONFIGS = {
    # config 2: single pair, 500-point contours, brute force 0.01 deg over +-180 (36 000 candidates)
    "cfg2": dict(mode=2, frames=6, points=500, seed=20261018, step=0.01, rng=180.0, brute=True, smooth=True, sample=500),
    # config 3: double pair, 1 000-point contours, coarse-to-fine 0.01 deg over +-180, sample_size 500 (800-point clouds)
    "cfg3": dict(mode=3, frames=5, points=1000, seed=20261118, step=0.01, rng=180.0, brute=False, smooth=True, sample=500),
    # config 4: full mode, 2 000-point contours, brute force 0.005 deg over +-180 (72 000 candidates), sample_size 2000
    "cfg4": dict(mode=4, frames=3, points=2000, seed=20261218, step=0.005, rng=180.0, brute=True, smooth=True, sample=2000),
    # config 5: cohort of 3 patients in full mode, 500-point contours, brute force 0.05 deg over +-90 (3 601 candidates)
    "cfg5": dict(mode=4, frames=8, points=500, seed=20261318, step=0.05, rng=90.0, brute=True, smooth=False, sample=500,
                 patients=3),
}


def quantised_pullback(frames, points, seed):
    xy = bench.synthetic_pullback(frames, points, seed)
    return np.rint(xy * Q).astype(np.int32)          # (frames, points, 2)


def rows_from_ints(k):
    """(frames, points, 2) int32 -> the (N, 4) [frame, x, y, z] rows and the reference point the entry points take."""
    F, P, _ = k.shape
    xy = k.astype(np.float64) / Q
    z = 0.5 * (F - 1 - np.arange(F))
    rows = np.concatenate([np.column_stack([np.full(P, float(i)), xy[i], np.full(P, z[i])]) for i in range(F)])
    last = rows[rows[:, 0] == F - 1][0]
    return rows, np.array([F - 1, last[1] + 0.1, last[2], last[3]])


def n_in(mode):
    return 4 if mode >= 3 else mode


def main():
    ora.build()
    inputs, gold = {}, {}
    for name, c in CONFIGS.items():
        t0 = time.time()
        patients = c.get("patients", 1)
        all_logs, all_sha = [], []
        for p in range(patients):
            blobs = []
            for k in range(n_in(c["mode"])):
                ints = quantised_pullback(c["frames"], c["points"], c["seed"] + 10 * p + k)
                inputs[f"{name}_p{p}_k{k}"] = ints
                rows, rp = rows_from_ints(ints)
                blobs.append(ora.build_geometry_from_arrays(rows, rp, diastole=k % 2 == 0, label=f"{name}_{p}_{k}"))
            outs, logs = ora.process(c["mode"], blobs, c["step"], c["rng"], c["smooth"], c["brute"], c["sample"], threads=8,
                                     postprocessing=False)
            for i, l in enumerate(logs):
                gold[f"{name}_p{p}_logs_{i}"] = l
            gold[f"{name}_p{p}_out_sha"] = np.array([gio.sha(o) for o in outs])
        print(f"{name}: {time.time() - t0:.1f} s", flush=True)
        np.savez_compressed(OUT / "config_inputs.npz", **inputs)
        np.savez_compressed(OUT / "config_goldens.npz", **gold)


if __name__ == "__main__":
    main()
