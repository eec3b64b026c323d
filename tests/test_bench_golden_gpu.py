"""tests/golden/bench_units.npz (the 1-GPU result of the bench workload that `bench.py --gpus N` compares every rank's
merged result with) is itself held against the CPU oracle: the first units of pullback pairs 0 and 7, all 36 000
candidates each, must give the golden's leftmost f64 arg-min and bit-equal f64 distance on the oracle AND on the GPU."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT)]
import bench  # noqa: E402
from multimodars import _native as nat  # noqa: E402
from oracle import oracle_py as ora  # noqa: E402

pytestmark = pytest.mark.gpu


def test_bench_golden_units_match_oracle_and_gpu():
    g = np.load(bench.GOLDEN)
    assert int(g["seed"]) == bench.SEED and int(g["n_pairs"]) == bench.N_PAIRS and len(g["best_idx"]) == 398 * bench.N_PAIRS
    txy, toff, rxy, roff, cen, U, n = bench.make_units()
    pick = [0, 1, 7 * 398, U - 1]
    sel = np.concatenate([np.arange(u * n, (u + 1) * n) for u in pick])
    off = np.arange(len(pick) + 1, dtype=np.int64) * n
    ctx = nat.Context(0)
    res = ctx.sweep_batched(txy[sel], off, rxy[sel], off, cen[pick], [nat.make_grid(bench.STEP_DEG, bench.RANGE_DEG)], mode=0)
    ctx.close()
    bi, bc = ora.sweep_batch(txy[sel], off, rxy[sel], off, cen[pick], 0, bench.STEP_DEG, bench.RANGE_DEG, bench.RANGE_DEG, threads=16)
    for k, u in enumerate(pick):
        assert int(bi[k]) == int(g["best_idx"][u]) == int(res["best_idx"][k]), u
        assert float(bc[k]) == float(g["best_dist"][u]) == float(res["best_dist"][k]), u
