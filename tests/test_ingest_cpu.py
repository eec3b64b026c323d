"""Ingest parity (host-only, no GPU): the product's mmrs_geometry_from_{dir,arrays} and the oracle's
restatement of io/build.rs:9-205 produce bit-identical geometry blobs, and both reproduce the
golden hashes made from the reference's example directories (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import multimodars as mm
from multimodars import _native as nat
from oracle import oracle_py as ora
from tests import golden_io as gio

CASES = [("rest", True), ("rest", False), ("stress", True), ("stress", False)]


@pytest.mark.parametrize("i,case", list(enumerate(CASES)))
def test_arrays_ingest_matches_oracle_and_golden(i, case):
    pack, gold = gio.inputs(), gio.oracle_outputs()
    name, dia = case
    a = gio.phase_arrays(pack, name, dia)
    want = ora.build_geometry_from_arrays(a["lumen"], a["ref_point"], a["eem"], a["calc"], a["side"], a["records"],
                                          dia, name)
    got = nat.geometry_from_arrays(a["lumen"], a["ref_point"], a["eem"], a["calc"], a["side"], a["records"], dia, name)
    assert np.array_equal(got, want)
    assert gio.sha(got) == str(gold[f"cfg1_in_sha_{i}"])   # == what the oracle built from the reference's CSV files


def test_dir_ingest_matches_oracle(tmp_path):
    pack = gio.inputs()
    for name in ("rest", "ideal"):
        d = gio.write_dir(pack, name, tmp_path / name)
        for dia in (True, False):
            assert np.array_equal(nat.geometry_from_dir(d, name, dia), ora.build_geometry_from_dir(d, name, dia))


def test_ingest_errors_match_reference_messages(tmp_path):
    with pytest.raises(nat.MmrsError, match="required contours file missing"):
        nat.geometry_from_dir(tmp_path, "x", True)
    pack = gio.inputs()
    a = gio.phase_arrays(pack, "rest", True)
    bad_ref = a["ref_point"].copy()
    bad_ref[0] = 99999  # frame absent from the lumen -> exactly-one-reference-point check fails (integrity_check.rs:107-118)
    with pytest.raises(nat.MmrsError, match="Expected exactly one reference point, found 0"):
        nat.geometry_from_arrays(a["lumen"], bad_ref, diastole=True)
    with pytest.raises(ora.OracleError, match="Expected exactly one reference point, found 0"):
        ora.build_geometry_from_arrays(a["lumen"], bad_ref, diastole=True)
    ragged = a["lumen"][:-7]  # unequal point counts per frame (integrity_check.rs:121-166)
    with pytest.raises(nat.MmrsError, match="point count mismatch"):
        nat.geometry_from_arrays(ragged, a["ref_point"], diastole=True)


def test_pyinputdata_round_trip_and_types():
    pack = gio.inputs()
    inp = gio.py_input(mm, pack, "rest", True, "rest_dia")
    assert isinstance(inp, mm.PyInputData) and len(inp.lumen) == 20 and len(inp.lumen[0]) == 501
    p = inp.lumen[0].points[3]
    assert isinstance(p, mm.PyContourPoint) and p.point_index == 3
    a = gio.phase_arrays(pack, "rest", True)
    blob = nat.geometry_from_arrays(inp._flat(inp.lumen), a["ref_point"], records=inp._records(), diastole=True,
                                    label="rest_dia")
    assert gio.sha(blob) == str(gio.oracle_outputs()["cfg1_in_sha_0"])
    g = mm.PyGeometry.from_blob(blob, "rest_dia")
    assert repr(g) == "Geometry(20 frames, label='rest_dia')"
    assert np.array_equal(g.to_blob(), blob)
    assert set(g.frames[0].extras) == {"Catheter"} and len(g.frames[0].extras["Catheter"]) == 20


def test_argument_errors_follow_the_reference():
    with pytest.raises(nat.MmrsError):      # missing input directory: anyhow error -> RuntimeError (functions.rs:228)
        mm.from_file_full("does/not/exist/a", "does/not/exist/b")
    with pytest.raises(TypeError):          # binding/align.rs:151
        mm.align_three_point(None, None, None, None, None)
