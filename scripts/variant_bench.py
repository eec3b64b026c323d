"""Times k_sweep for the shipped library and every multimoda-rs_b200/variants/libmmrs_*.so (experimental builds; not shipped)."""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path[:0] = [str(ROOT), str(ROOT / "multimoda-rs_b200")]
    import numpy as np
    from multimodars import _native as nat
    sys.path.insert(0, str(ROOT / "scripts"))
    import quick_bench as qb
    ctx = nat.Context(0)
    qb.run(ctx, 80, 520, 0.01, 180.0)
    qb.run(ctx, 80, 510, 0.01, 180.0)
    if os.environ.get("MMRS_VARIANT_FULL"):
        qb.run(ctx, 16, 2020, 0.05, 180.0)
else:
    libs = [ROOT / "multimoda-rs_b200" / "libmmrs_b200.so"] + sorted((ROOT / "multimoda-rs_b200" / "variants").glob("libmmrs_*.so"))
    for lib in libs:
        print("==", lib.name, flush=True)
        env = dict(os.environ, MMRS_B200_LIB=str(lib))
        subprocess.run([sys.executable, __file__, "--one"], env=env)
    print("== libmmrs_b200.so with MMRS_NO_TAILPASS=1 (padded TA=17)", flush=True)
    subprocess.run([sys.executable, __file__, "--one"], env=dict(os.environ, MMRS_NO_TAILPASS="1"))
