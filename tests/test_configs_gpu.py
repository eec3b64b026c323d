"""BASELINE.json configs 2-5 at their full search settings through the public entry points on the GPU, against the
committed oracle goldens (tests/golden/config_goldens.npz, made by tests/golden/make_config_goldens.py from
tests/golden/config_inputs.npz): per-frame logs and every output geometry must be BIT-IDENTICAL.

  cfg2  from_array_singlepair  500-point contours, brute force 0.01 deg over +-180 (36 000 candidates)
  cfg3  from_array_doublepair  1 000-point contours, coarse-to-fine 0.01 deg over +-180, 800-point inter-pullback clouds
  cfg4  from_array_full        2 000-point contours (N = M = 2 020), brute force 0.005 deg over +-180 (72 000 candidates)
  cfg5  a 3-patient cohort, full mode, brute force 0.05 deg over +-90: one mmrs_process_cases call, and the pipelined
        cohort path

Config 1 (the example pullbacks) lives in tests/test_process_gpu.py."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "tests" / "golden")]

import multimodars as mm
from multimodars import _dist
from multimodars import _native as nat
from tests import golden_io as gio

import make_config_goldens as mk  # noqa: E402  (the generator: same CONFIGS table, same input decoding)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def packs():
    return np.load(gio.GOLD / "config_inputs.npz"), np.load(gio.GOLD / "config_goldens.npz")


def py_inputs(ins, name, p, mode):
    out = []
    for k in range(mk.n_in(mode)):
        rows, rp = mk.rows_from_ints(ins[f"{name}_p{p}_k{k}"])
        out.append(mm.numpy_to_inputdata(rows, rp, k % 2 == 0, label=f"{name}_{p}_{k}"))
    return out


def check(gold, name, p, logs, geoms):
    for i, l in enumerate(logs):
        assert np.array_equal(np.array(l, dtype=np.float64).reshape(-1, 7), gold[f"{name}_p{p}_logs_{i}"]), (name, p, i)
    want = [str(s) for s in gold[f"{name}_p{p}_out_sha"]]
    got = [gio.sha(g if isinstance(g, np.ndarray) else g.to_blob()) for g in geoms]
    assert got == want, (name, p)


def test_config2_singlepair_bruteforce_36000_candidates(packs):
    ins, gold = packs
    c = mk.CONFIGS["cfg2"]
    a, b = py_inputs(ins, "cfg2", 0, 2)
    pair, logs = mm.from_array_singlepair(a, b, step_rotation_deg=c["step"], range_rotation_deg=c["rng"], sample_size=c["sample"],
                                          write_obj=False, bruteforce=True, smooth=c["smooth"], postprocessing=False)
    check(gold, "cfg2", 0, logs, [pair.geom_a, pair.geom_b])
    st = mm.get_context().process_stats()
    assert st["evals"] >= 2 * (c["frames"] - 1) * 36000          # every candidate of every frame pair was scored
    assert mm.get_context().plan()["TA"] >= 2


def test_config3_doublepair_hierarchical_1000_point_contours(packs):
    ins, gold = packs
    c = mk.CONFIGS["cfg3"]
    four = py_inputs(ins, "cfg3", 0, 3)
    ab, cd, logs = mm.from_array_doublepair(*four, step_rotation_deg=c["step"], range_rotation_deg=c["rng"],
                                            sample_size=c["sample"], write_obj=False, bruteforce=False, smooth=c["smooth"],
                                            postprocessing=False)
    check(gold, "cfg3", 0, logs, [ab.geom_a, ab.geom_b, cd.geom_a, cd.geom_b])


def test_config4_full_mode_oct_resolution_72000_candidates(packs):
    ins, gold = packs
    c = mk.CONFIGS["cfg4"]
    four = py_inputs(ins, "cfg4", 0, 4)
    ab, cd, ac, bd, logs = mm.from_array_full(*four, step_rotation_deg=c["step"], range_rotation_deg=c["rng"],
                                              sample_size=c["sample"], write_obj=False, bruteforce=True, smooth=c["smooth"],
                                              postprocessing=False)
    check(gold, "cfg4", 0, logs, [ab.geom_a, ab.geom_b, cd.geom_a, cd.geom_b, ac.geom_a, ac.geom_b, bd.geom_a, bd.geom_b])
    st = mm.get_context().process_stats()
    assert st["evals"] >= 4 * (c["frames"] - 1) * 72000


def _cohort_blobs(ins):
    c = mk.CONFIGS["cfg5"]
    blobs = []
    for p in range(c["patients"]):
        for k in range(4):
            rows, rp = mk.rows_from_ints(ins[f"cfg5_p{p}_k{k}"])
            blobs.append(nat.geometry_from_arrays(rows, rp, diastole=k % 2 == 0, label=f"cfg5_{p}_{k}"))
    return c, blobs


def test_config5_cohort_in_one_call(packs):
    ins, gold = packs
    c, blobs = _cohort_blobs(ins)
    outs, logs, _ = nat.process_cases(mm.get_context(), 4, blobs, c["step"], c["rng"], c["sample"], c["smooth"], True, False)
    for p in range(c["patients"]):
        check(gold, "cfg5", p, logs[4 * p:4 * p + 4], outs[8 * p:8 * p + 8])


def test_config5_cohort_pipelined(packs):
    ins, gold = packs
    c, blobs = _cohort_blobs(ins)
    outs, logs, _, st = _dist.process_cases_pipelined(0, 4, blobs, c["step"], c["rng"], c["sample"], c["smooth"], True,
                                                      postprocessing=False, chunk_cases=1, workers=3)
    for p in range(c["patients"]):
        check(gold, "cfg5", p, logs[4 * p:4 * p + 4], outs[8 * p:8 * p + 8])
    assert st["evals"] >= c["patients"] * 4 * (c["frames"] - 1) * 3601
