"""Wall time of the public entry points on BASELINE.json's other configurations (1, 3, 5-mini), warm context.
Not the bench line (bench.py measures configs[1]); these are the 'full-mode align wall time' figures of DESIGN.md §7."""
import json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "multimoda-rs_b200")]
import bench
import multimodars as mm
from multimodars import _native as nat
from tests import golden_io as gio


def rows(seed, n_frames, n_points):
    fr = bench.synthetic_pullback(n_frames, n_points, seed)
    z = 0.5 * (n_frames - 1 - np.arange(n_frames))
    a = np.concatenate([np.column_stack([np.full(n_points, float(i)), f, np.full(n_points, z[i])]) for i, f in enumerate(fr)])
    last = a[a[:, 0] == n_frames - 1][0]
    return a, np.array([n_frames - 1, last[1] + 0.1, last[2], last[3]])

def timed(f, reps=3):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter(); r = f(); dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return best, r


def main():
    out = {}
    ctx = mm.get_context()
    # warm-up (CUDA context, module load)
    a, rp = rows(1, 8, 100)
    mm.from_array_single(mm.numpy_to_inputdata(a, rp, True, label="w"), sample_size=100)

    # ---- config 1: examples ivus_rest + ivus_stress, from_array_full, defaults (hierarchical 0.5 deg, +-90) and brute 0.05
    pack = gio.inputs()
    ins = [gio.py_input(mm, pack, n, d, f"{n}{d}") for n, d in (("rest", True), ("rest", False), ("stress", True), ("stress", False))]
    for tag, kw in (("default_hier_0p5", dict(step_rotation_deg=0.5, bruteforce=False)),
                    ("brute_0p05", dict(step_rotation_deg=0.05, bruteforce=True)),
                    ("hier_0p05", dict(step_rotation_deg=0.05, bruteforce=False))):
        dt, _ = timed(lambda: mm.from_array_full(*ins, range_rotation_deg=90.0, sample_size=500, write_obj=False, smooth=False,
                                                 postprocessing=False, **kw))
        st = ctx.process_stats()
        out[f"config1_{tag}"] = dict(wall_s=dt, **st, evals_per_s=st["evals"] / dt)
        print(f"config1 {tag}: {dt*1e3:.1f} ms", st, flush=True)

    # ---- config 3: double pair, 4 x 400 frames x 1000 pts, hierarchical 0.01 deg, +-180, sample 500
    ins3 = []
    for k, dia in enumerate((True, False, True, False)):
        a, rp = rows(300 + k, 400, 1000)
        ins3.append(mm.numpy_to_inputdata(a, rp, dia, label=f"p{k}"))
    dt, _ = timed(lambda: mm.from_array_doublepair(*ins3, step_rotation_deg=0.01, range_rotation_deg=180.0, sample_size=500,
                                                   write_obj=False, smooth=True, postprocessing=False), reps=2)
    st = ctx.process_stats()
    out["config3_doublepair_hier_0p01"] = dict(wall_s=dt, **st, evals_per_s=st["evals"] / dt)
    print(f"config3: {dt*1e3:.1f} ms", st, flush=True)
    t0 = time.perf_counter()
    blobs3 = [nat.geometry_from_arrays(i._flat(i.lumen), np.array([i.ref_point.frame_index, i.ref_point.x, i.ref_point.y, i.ref_point.z]),
                                       diastole=i.diastole, label=i.label) for i in ins3]
    t_ing = time.perf_counter() - t0
    dt, _ = timed(lambda: nat.process_cases(ctx, 3, blobs3, 0.01, 180.0, 500, True, False), reps=2)
    out["config3_doublepair_hier_0p01"].update(ingest_s=t_ing, process_cases_s=dt)
    print(f"config3 ingest {t_ing*1e3:.1f} ms, process_cases {dt*1e3:.1f} ms", flush=True)

    # ---- config 5 (mini): cohort of P patients in full mode, 200 frames x 500 pts, brute 0.05 deg, +-90, ONE call
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    blobs = []
    for p in range(P):
        for k, dia in enumerate((True, False, True, False)):
            a, rp = rows(20261018 + 1000 * p + k, 200, 500)
            blobs.append(nat.geometry_from_arrays(a, rp, diastole=dia, label=f"pt{p}_{k}"))
    dt, _ = timed(lambda: nat.process_cases(ctx, 4, blobs, 0.05, 90.0, 500, False, True), reps=2)
    st = ctx.process_stats()
    out[f"config5_cohort_{P}_patients_full_brute_0p05"] = dict(wall_s=dt, **st, evals_per_s=st["evals"] / dt)
    print(f"config5 ({P} patients): {dt*1e3:.1f} ms", st, flush=True)

    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "config_bench.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
