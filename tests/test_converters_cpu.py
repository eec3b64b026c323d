"""The reference's array <-> object converters (multimodars/_converters.py:19-686) and `to_obj`
(binding/functions.rs:1435-1501) as exposed by this package: shapes, grouping, error messages, file names."""
import os

import numpy as np
import pytest

import multimodars as mm
from tests.test_centerline_cpu import pullback
from tests.test_export_cpu import expected_obj


def test_to_array_shapes_and_round_trip():
    g = pullback(n_frames=3, n=12, extras=True, ref_frame=2, ref_index=4)
    d = mm.to_array(g)
    assert set(d) == {"lumen", "eem", "calcification", "sidebranch", "catheter", "wall", "reference"}
    assert d["lumen"].shape == (36, 4) and d["catheter"].shape == (9, 4) and d["wall"].shape == (0, 4)
    assert d["reference"].shape == (1, 4) and d["reference"][0, 0] == 2.0
    assert np.array_equal(d["lumen"][:12, 0], np.zeros(12)) and d["lumen"][12, 0] == 1.0
    assert np.array_equal(mm.to_array(g.frames[1].lumen), d["lumen"][12:24])
    f = mm.to_array(g.frames[0])
    assert set(f) == {"lumen", "catheter", "reference"} and f["reference"].shape == (0, 4)
    a, b = mm.to_array(mm.PyGeometryPair(g, g, "p"))
    assert np.array_equal(a["lumen"], b["lumen"])
    back = mm.numpy_to_geometry(d["lumen"], catheter_arr=d["catheter"], reference_arr=d["reference"], label="x")
    assert repr(back) == "Geometry(3 frames, label='x')"
    for fa, fb in zip(g.frames, back.frames):
        assert np.array_equal(fa.lumen.points_array()[:, 2:5], fb.lumen.points_array()[:, 2:5])
        assert list(fb.lumen.points_array()[:, 1]) == list(range(12))
        assert np.allclose(fb.centroid, fa.lumen.centroid) and set(fb.extras) == {"Catheter"}
        assert fb.reference_point is not None and fb.reference_point.frame_index == 2   # the same point on every frame
    with pytest.raises(ValueError, match="lumen_arr cannot be empty"):
        mm.numpy_to_geometry(np.zeros((0, 4)))
    with pytest.raises(TypeError, match="Unsupported type for to_array"):
        mm.to_array(3)


def test_to_array_of_input_data():
    lum = np.array([[0, 1.0, 2.0, 0.5], [0, 2.0, 2.0, 0.5], [1, 1.0, 3.0, 1.0]])
    inp = mm.numpy_to_inputdata(lum, np.array([1, 1.0, 3.0, 1.0]), True, label="lbl")
    d = mm.to_array(inp)
    assert d["label"] == "lbl" and d["diastole"] is True
    assert d["lumen"].shape == (3, 4) and d["reference"].tolist() == [[1.0, 1.0, 3.0, 1.0]] and d["eem"].shape == (0, 4)


def test_numpy_to_centerline_interpolates_and_validates():
    cl = mm.numpy_to_centerline(np.array([[0.0, np.nan, 3.0], [0.0, 1.0, 2.0], [np.nan, 2.0, 1.0], [0.0, np.nan, 0.0]]))
    arr = mm.to_array(cl)
    assert arr.shape == (4, 4) and arr[:, 0].tolist() == [0.0, 1.0, 2.0, 3.0]
    assert arr[:, 2].tolist() == [1.0, 1.0, 2.0, 2.0] and arr[2, 1] == 0.0       # edges repeat, gaps interpolate
    assert cl.points[-1].tangent == cl.points[-2].tangent
    with pytest.raises(ValueError, match=r"Input must be a \(N,3\) array"):
        mm.numpy_to_centerline(np.zeros((3, 4)))
    with pytest.raises(ValueError, match="at least one point"):
        mm.numpy_to_centerline(np.zeros((0, 3)))
    with pytest.raises(ValueError, match="All values are NaN for coordinate column 1"):
        mm.numpy_to_centerline(np.array([[0.0, np.nan, 1.0], [0.0, np.nan, 0.0]]))
    with pytest.raises(ValueError, match="at least two points"):
        mm.numpy_to_centerline(np.array([[0.0, 0.0, 1.0]]))


def test_to_obj_names_and_content(tmp_path):
    g = pullback(n_frames=3, n=12, extras=True)
    mm.to_obj(g, str(tmp_path), contour_types=[mm.PyContourType.Lumen, mm.PyContourType.Wall], filename_prefix="case")
    mm.to_obj(g, str(tmp_path / "plain"), watertight=False)
    assert sorted(os.listdir(tmp_path)) == ["case_lumen.mtl", "case_lumen.obj", "plain"]
    assert sorted(os.listdir(tmp_path / "plain")) == ["catheter.mtl", "catheter.obj", "lumen.mtl", "lumen.obj"]
    lum = [f.lumen for f in g.frames]
    assert open(tmp_path / "case_lumen.obj").read() == expected_obj(lum, [(0.0, 0.0)] * 36, str(tmp_path / "case_lumen.mtl"), True)
    assert open(tmp_path / "plain" / "lumen.obj").read() == expected_obj(lum, [(0.0, 0.0)] * 36,
                                                                       str(tmp_path / "plain" / "lumen.mtl"), False)
    one = mm.PyGeometry(g.frames[:1], "one")
    with pytest.raises(mm.MmrsError, match="Failed to write lumen OBJ: .*Need at least two contours"):
        mm.to_obj(one, str(tmp_path / "x"))


# ---- numpy_to_inputdata / array_to_pyinputdata (multimodars/_converters.py:204-437, 689-966) -------------------------
def test_numpy_to_inputdata_reference_semantics():
    from multimodars._converters import numpy_to_inputdata
    lum = np.array([[3, 1.0, 2.0, 3.0], [1, 0.0, 0.0, 1.0], [3, 2.0, 2.0, 3.0], [1, 1.0, 1.0, 1.0]])
    eem = np.array([[1, 9.0, 9.0, 1.0], [7, 5.0, 5.0, 5.0]])                 # frame 7 has no lumen: ignored
    rec = np.array([(1, "D", 1.5, np.nan), (3, "S", np.nan, 2.5)],
                   dtype=[("frame", "i4"), ("phase", "U1"), ("m1", "f8"), ("m2", "f8")])
    inp = numpy_to_inputdata(lum, np.array([[3, 7.0, 8.0, 9.0]]), 1, record=rec, eem_arr=eem,
                             calcification=np.zeros((0, 4)), label=None)
    assert [c.id for c in inp.lumen] == [1, 3] and [len(c) for c in inp.lumen] == [2, 2]
    assert [p.point_index for p in inp.lumen[1].points] == [0, 1] and inp.lumen[1].points[0].x == 1.0   # row order kept
    assert inp.lumen[0].centroid == (0.5, 0.5, 1.0)
    assert [c.id for c in inp.eem] == [1] and inp.calcification is None and inp.sidebranch is None
    assert [(r.frame, r.phase, r.measurement_1, r.measurement_2) for r in inp.record] == [(1, "D", 1.5, None),
                                                                                           (3, "S", None, 2.5)]
    assert (inp.ref_point.frame_index, inp.ref_point.x, inp.ref_point.z) == (3, 7.0, 9.0)
    assert inp.diastole is True and inp.label == ""
    # numeric phases: 0 -> "D", anything else -> "S"; short rows leave the measurements empty
    r = numpy_to_inputdata(lum, None, False, record=np.array([[1, 0], [3, 1]])).record
    assert [(x.phase, x.measurement_1) for x in r] == [("D", None), ("S", None)]
    # unusable reference point -> origin on frame 0; empty record array -> None
    inp = numpy_to_inputdata(lum, np.array([1.0, 2.0]), True, record=np.zeros((0, 4)))
    assert (inp.ref_point.frame_index, inp.ref_point.x) == (0, 0.0) and inp.record is None
    with pytest.raises(ValueError, match="lumen_arr cannot be empty"):
        numpy_to_inputdata(np.zeros((0, 4)), np.zeros(4), True)


def test_array_to_pyinputdata_inverts_to_array():
    from multimodars._converters import array_to_pyinputdata, numpy_to_inputdata, to_array
    lum = np.array([[0, 1.0, 2.0, 3.0], [0, 2.0, 2.0, 3.0], [1, 5.0, 5.0, 5.0]])
    side = np.array([[1, 4.0, 4.0, 5.0]])
    inp = numpy_to_inputdata(lum, np.array([0, 9.0, 9.0, 9.0]), False, sidebranch=side, label="x",
                             record=np.array([[0, "D", 1.0, None], [1, "S", None, 2.0]], dtype=object))
    d = to_array(inp)
    back = array_to_pyinputdata(d["lumen"], d["eem"], d["calcification"], d["sidebranch"], d["records"], d["reference"],
                                d["diastole"], d["label"])
    assert np.array_equal(to_array(back)["lumen"], lum) and np.array_equal(to_array(back)["sidebranch"], side)
    assert back.eem is None and back.calcification is None and back.diastole is False and back.label == "x"
    assert [(r.frame, r.phase, r.measurement_1, r.measurement_2) for r in back.record] == [(0, "D", 1.0, None),
                                                                                            (1, "S", None, 2.0)]
    assert (back.ref_point.frame_index, back.ref_point.x) == (0, 9.0)
    # existing objects pass through untouched; row lists and structured records work; the first NON-ZERO reference row
    same = array_to_pyinputdata(lumen=inp.lumen, records=[inp.record[0], (5, "S", 3.0)],
                                reference=np.array([[0, 0, 0, 0], [2, 1.0, 1.0, 1.0]]))
    assert same.lumen[0] is inp.lumen[0] and same.record[0] is inp.record[0]
    assert (same.record[1].frame, same.record[1].measurement_1, same.record[1].measurement_2) == (5, 3.0, None)
    assert same.ref_point.frame_index == 2 and same.diastole is True
    st = np.array([(4, "D", 0.5)], dtype=[("Frame", "i4"), ("PHASE", "U1"), ("m1", "f8")])
    r = array_to_pyinputdata(lumen=lum, records=st).record
    assert (r[0].frame, r[0].phase, r[0].measurement_1, r[0].measurement_2) == (4, "D", 0.5, None)
    assert array_to_pyinputdata(lumen=lum).ref_point.x == 0.0
    for bad, msg in ((np.array([1.0, 2.0, 3.0]), "1D array must have length 4"), (np.zeros((2, 3)), r"must be \(N,4\)-like")):
        with pytest.raises(ValueError, match=msg):
            array_to_pyinputdata(lumen=bad)
    with pytest.raises(ValueError, match="reference must be length 4"):
        array_to_pyinputdata(lumen=lum, reference=np.zeros(3))
    with pytest.raises(ValueError, match="Unsupported records format"):
        array_to_pyinputdata(lumen=lum, records="nope")
