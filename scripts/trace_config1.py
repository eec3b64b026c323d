"""Phase trace (MMRS_TRACE=1) of the reference's headline benchmark call: from_file_full on the example pullbacks,
coarse-to-fine at 0.05 deg, plus Python-side timings of ingest / native call / result decoding."""
import os, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "multimoda-rs_b200")]
os.environ["MMRS_TRACE"] = "1"
import multimodars as mm
from multimodars import _native as nat, _processing as P
from tests import golden_io as gio

tmp = tempfile.mkdtemp()
pack = gio.inputs()
ab = str(gio.write_dir(pack, "rest", Path(tmp) / "ivus_rest"))
cd = str(gio.write_dir(pack, "stress", Path(tmp) / "ivus_stress"))
kw = dict(step_rotation_deg=0.05, range_rotation_deg=90.0, write_obj=False, smooth=False, postprocessing=False,
          bruteforce=False, interpolation_steps=0)
mm.from_file_full(ab, cd, **kw)
for rep in range(3):
    print(f"---- rep {rep}", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    blobs, names = P._four_from_paths(ab, cd, None, (4.5, 4.5), 0.5, 20)
    t1 = time.perf_counter()
    out, logs, _ = nat.process_cases(mm.get_context(), 4, blobs, 0.05, 90.0, 500, False, False, False)
    t2 = time.perf_counter()
    geoms = [mm.PyGeometry.from_blob(o, "x") for o in out]
    t3 = time.perf_counter()
    print(f"python: ingest {1e3*(t1-t0):.2f} ms, process_cases {1e3*(t2-t1):.2f} ms, from_blob x{len(out)} {1e3*(t3-t2):.2f} ms",
          file=sys.stderr, flush=True)
    t0 = time.perf_counter(); mm.from_file_full(ab, cd, **kw); print(f"python: whole call {1e3*(time.perf_counter()-t0):.2f} ms", file=sys.stderr, flush=True)
