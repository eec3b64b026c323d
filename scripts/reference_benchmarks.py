"""The reference's own two published benchmarks (docs/benchmark.rst, benchmarks/benchmark_bruteforce_stepsize.py and
benchmark_cpu_scaling.py), re-run against this package on the GPU box. Same calls, same arguments, median of 3.

1. `from_file_full` on the example IVUS rest/stress pullbacks (re-created as CSV directories from
   tests/golden/inputs.npz — /root/reference does not exist on the box), range 90 deg, step in
   {5, 2.5, 1, 0.5, 0.25, 0.1, 0.05} deg, brute force and coarse-to-fine, write_obj / smooth / postprocessing off.
   Published (16-thread Xeon Gold 6234): 64.4 s brute / 6.25 s coarse-to-fine at 0.05 deg.
2. `from_array_single` on an OCT-like pullback: 280 frames, step 0.01 deg, range 6 deg, sample_size 200,
   image_center (5, 5), n_points 40. The reference's OCT contour file is not in its checkout
   (.MISSING_LARGE_BLOBS), so the pullback is SYNTHETIC of the same shape (280 frames x 500 points, bench.py's
   generator) — comparable in work, not in data. Published: 14.15 s brute / 2.40 s coarse-to-fine at 16 threads.

Writes gpurun_out/reference_benchmarks.json."""
import json
import statistics
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "multimoda-rs_b200")]
import bench  # noqa: E402
import multimodars as mm  # noqa: E402
from tests import golden_io as gio  # noqa: E402

STEP_SIZES = [5.0, 2.5, 1.0, 0.5, 0.25, 0.1, 0.05]
RANGE_DEG = 90.0
REPEATS = 3
PUBLISHED = {"stepsize_0p05": {"bruteforce_s": 64.4, "optimized_s": 6.25},
             "oct_16_threads": {"bruteforce_s": 14.15, "optimized_s": 2.40}}


def median_of(f, n=REPEATS):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts)


def main():
    out = {"published_reference_cpu": PUBLISHED, "repeats": REPEATS}
    ctx = mm.get_context()
    with tempfile.TemporaryDirectory() as tmp:
        pack = gio.inputs()
        ab = str(gio.write_dir(pack, "rest", Path(tmp) / "ivus_rest"))
        cd = str(gio.write_dir(pack, "stress", Path(tmp) / "ivus_stress"))

        def full(step, brute):
            return mm.from_file_full(input_path_ab=ab, input_path_cd=cd, step_rotation_deg=step,
                                     range_rotation_deg=RANGE_DEG, write_obj=False, smooth=False, postprocessing=False,
                                     bruteforce=brute, interpolation_steps=0)

        full(5.0, False)  # CUDA context, module load
        rows = []
        for step in STEP_SIZES:
            bf = median_of(lambda: full(step, True))
            evals_bf = ctx.process_stats()["evals"]
            opt = median_of(lambda: full(step, False))
            evals_opt = ctx.process_stats()["evals"]
            rows.append({"step_deg": step, "n_steps": 2 * RANGE_DEG / step, "bruteforce_s": bf, "optimized_s": opt,
                         "bruteforce_evals": evals_bf, "optimized_evals": evals_opt})
            print(f"step={step:5.2f}  brute {bf * 1e3:8.1f} ms ({evals_bf} evals)   coarse-to-fine {opt * 1e3:8.1f} ms "
                  f"({evals_opt} evals)", flush=True)
        out["stepsize_from_file_full"] = rows
        last = rows[-1]
        out["stepsize_0p05_speedup_vs_published"] = {
            "bruteforce": PUBLISHED["stepsize_0p05"]["bruteforce_s"] / last["bruteforce_s"],
            "optimized": PUBLISHED["stepsize_0p05"]["optimized_s"] / last["optimized_s"]}

    # ---- benchmark 2: OCT-like single pullback
    n_frames, n_points = 280, 500
    fr = bench.synthetic_pullback(n_frames, n_points, 20261018)
    fr = fr + 0.5  # centred about (5, 5) like the OCT acquisition
    z = 0.2 * (n_frames - 1 - np.arange(n_frames))
    a = np.concatenate([np.column_stack([np.full(n_points, float(i)), f, np.full(n_points, z[i])])
                        for i, f in enumerate(fr)])
    first = a[a[:, 0] == n_frames - 1][0]
    ref = np.array([n_frames - 1, first[1] + 0.1, first[2], first[3]])
    oct_input = mm.numpy_to_inputdata(lumen_arr=a, ref_point=ref, record=None, diastole=True, label="oct")

    def single(brute):
        return mm.from_array_single(input_data=oct_input, step_rotation_deg=0.01, range_rotation_deg=6.0,
                                    sample_size=200, image_center=(5.0, 5.0), n_points=40, write_obj=False,
                                    smooth=False, bruteforce=brute)

    single(False)
    bf = median_of(lambda: single(True))
    st_bf = ctx.process_stats()
    opt = median_of(lambda: single(False))
    st_opt = ctx.process_stats()
    out["oct_like_from_array_single"] = {
        "data": "synthetic, 280 frames x 500 points (the reference's OCT file is not in its checkout)",
        "bruteforce_s": bf, "optimized_s": opt, "bruteforce_stats": st_bf, "optimized_stats": st_opt,
        "speedup_vs_published_16_threads": {"bruteforce": PUBLISHED["oct_16_threads"]["bruteforce_s"] / bf,
                                            "optimized": PUBLISHED["oct_16_threads"]["optimized_s"] / opt}}
    print(f"OCT-like: brute {bf * 1e3:.1f} ms, coarse-to-fine {opt * 1e3:.1f} ms", flush=True)

    dst = ROOT / "gpurun_out"
    dst.mkdir(exist_ok=True)
    (dst / "reference_benchmarks.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out["stepsize_0p05_speedup_vs_published"]))


if __name__ == "__main__":
    main()
