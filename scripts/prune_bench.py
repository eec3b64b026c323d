"""Wall time of the public entry points with and without exact lower-bound pruning (mmrs_ctx_set_prune):
config 2 (from_array_singlepair, brute force 0.01 deg over +-180 deg), a config-4 shaped single pullback
(from_array_single, 60 frames x 2000 pts, 0.005 deg over +-180 deg) and a config-5 cohort of P patients.
Checks that logs and geometries are bit-identical."""
import json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "multimoda-rs_b200")]
import bench
import multimodars as mm
from multimodars import _native as nat
from scripts.config_bench import rows

ctx = mm.get_context()
out = {}


def run(name, f, same):
    res = {}
    for tag, on in (("dense", False), ("pruned", True)):
        ctx.set_prune(on)
        best, val = None, None
        for _ in range(2):
            t0 = time.perf_counter(); val = f(); dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        res[tag] = (best, val, ctx.process_stats())
    ctx.set_prune(False)
    ok = same(res["dense"][1], res["pruned"][1])
    st = res["dense"][2]
    out[name] = dict(dense_wall_s=res["dense"][0], pruned_wall_s=res["pruned"][0], speedup=res["dense"][0] / res["pruned"][0],
                     evals=st["evals"], units=st["units"], identical=bool(ok))
    print(name, json.dumps(out[name]), flush=True)


def logs_equal(a, b):
    la, lb = a[-1], b[-1]
    if isinstance(la, tuple):
        return all(np.array_equal(np.array(x), np.array(y)) for x, y in zip(la, lb))
    return np.array_equal(np.array(la), np.array(lb))


ins = []
for k, dia in enumerate((True, False)):
    a, rp = rows(20261018 + k, 200, 500)
    ins.append(mm.numpy_to_inputdata(a, rp, dia, label=f"p{k}"))
run("config2_from_array_singlepair_brute_0p01", lambda: mm.from_array_singlepair(
    *ins, step_rotation_deg=0.01, range_rotation_deg=180.0, sample_size=500, write_obj=False, bruteforce=True, smooth=True,
    postprocessing=False), lambda x, y: logs_equal(x, y) and np.array_equal(x[0].geom_b.to_blob(), y[0].geom_b.to_blob()))

a, rp = rows(77, 60, 2000)
one = mm.numpy_to_inputdata(a, rp, True, label="oct")
run("config4_shape_from_array_single_60x2000_brute_0p005", lambda: mm.from_array_single(
    one, step_rotation_deg=0.005, range_rotation_deg=180.0, sample_size=2000, bruteforce=True, smooth=True),
    lambda x, y: logs_equal(x, y) and np.array_equal(x[0].to_blob(), y[0].to_blob()))

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
blobs = []
for p in range(P):
    for k, dia in enumerate((True, False, True, False)):
        a, rp = rows(20261018 + 1000 * p + k, 200, 500)
        blobs.append(nat.geometry_from_arrays(a, rp, diastole=dia, label=f"pt{p}_{k}"))
run(f"config5_cohort_{P}_patients_full_brute_0p05", lambda: nat.process_cases(ctx, 4, blobs, 0.05, 90.0, 500, False, True),
    lambda x, y: all(np.array_equal(p, q) for p, q in zip(x[0], y[0])) and all(np.array_equal(p, q) for p, q in zip(x[1], y[1])))

(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "prune_bench.json").write_text(json.dumps(out, indent=1))
