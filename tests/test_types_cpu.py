"""The reference's Python-level type tests (tests/test_core.py:11-101 with the fixtures of tests/conftest.py:90-242),
run against this package's value types, plus checks of the helpers against the oracle's arithmetic."""
import math

import numpy as np
import pytest

from multimodars import PyContour, PyContourPoint, PyFrame, PyGeometry


def _centroid(points):
    n = len(points)
    return (sum(p.x for p in points) / n, sum(p.y for p in points) / n, sum(p.z for p in points) / n)


def _ring(cid, center, rx, ry, n=20, aortic=False):
    pts = [PyContourPoint(cid, i, center[0] + rx * math.cos(t), center[1] + ry * math.sin(t), center[2], aortic)
           for i, t in enumerate(np.linspace(0, 2 * np.pi, n, endpoint=False))]
    return PyContour(cid, cid, pts, _centroid(pts), None, None, "Lumen")


@pytest.fixture
def sample_contour():
    pts = [PyContourPoint(0, 0, 0.0, 4.0, 3.0, False), PyContourPoint(0, 1, 0.0, 0.0, 3.0, False),
           PyContourPoint(0, 2, 4.0, 0.0, 3.0, True), PyContourPoint(0, 3, 4.0, 4.0, 3.0, True)]
    return PyContour(0, 0, pts, _centroid(pts), None, None, "Lumen")


@pytest.fixture
def sample_contour_round():
    return _ring(1, (0.0, 0.0, 3.0), 4.0, 4.0, aortic=True)


@pytest.fixture
def sample_contour_elliptic():
    return _ring(2, (2.0, 2.0, 3.0), 5.0, 3.0)


@pytest.fixture
def sample_geometry(sample_contour):
    return PyGeometry([PyFrame(0, sample_contour.centroid, sample_contour, {}, None)], "test")


def test_contour_point_distance():  # test_core.py:11-14
    assert PyContourPoint(0, 0, 0.0, 0.0, 0.0, False).distance(PyContourPoint(0, 0, 3.0, 4.0, 0.0, False)) == 5.0


def test_contour_area(sample_contour, sample_contour_round, sample_contour_elliptic):  # :22-30
    assert sample_contour.get_area() == pytest.approx(16.0)
    assert sample_contour_round.get_area() > 0 and sample_contour_elliptic.get_area() > 0


def test_contour_centroid(sample_contour):  # :33-48
    assert sample_contour.centroid == pytest.approx((2.0, 2.0, 3.0), abs=1e-6)


def test_find_farthest_points(sample_contour, sample_contour_elliptic):  # :51-58
    (p1, p2), d = sample_contour.find_farthest_points()
    assert d == pytest.approx(math.sqrt(32.0)) and (p1.point_index, p2.point_index) == (0, 2)
    assert sample_contour_elliptic.find_farthest_points()[1] == pytest.approx(10.0)
    assert sample_contour_elliptic.get_elliptic_ratio() == pytest.approx(10.0 / 6.0)


def test_rotation_and_translation_contour(sample_contour, sample_contour_elliptic):  # :61-79
    r = sample_contour.rotate(90)
    assert isinstance(r, PyContour) and len(r.points) == 4
    assert (r.points[0].x, r.points[0].y) == pytest.approx((0.0, 0.0), abs=1e-12)   # (0,4) about (2,2) by +90 deg
    t = sample_contour_elliptic.translate(-2.0, 3.0, 0.0)
    assert isinstance(t, PyContour) and t.points[0].x == pytest.approx(sample_contour_elliptic.points[0].x - 2.0)
    assert sample_contour.points[0].y == 4.0                                          # originals untouched


def test_geometry_creation_rotation_translation(sample_geometry):  # :82-101
    assert isinstance(sample_geometry, PyGeometry) and len(sample_geometry.frames) == 1 and sample_geometry.frames[0].id == 0
    r = sample_geometry.rotate(90)
    assert isinstance(r, PyGeometry) and len(r.frames) == 1
    t = sample_geometry.translate(1.0, 2.0, 3.0)
    assert t.frames[0].centroid == pytest.approx((3.0, 4.0, 6.0))
    assert t.frames[0].lumen.centroid == pytest.approx((3.0, 4.0, 6.0))


def test_helpers_follow_the_reference_arithmetic():
    """sort / rotate / smooth on a PyGeometry equal the oracle's restatement (contour.rs:368-405, frame.rs:40-63,
    geometry.rs:165-250) bit for bit: the oracle's within-alignment with a 0-candidate search is not needed —
    compare against its align_within post-step helpers through a blob round trip instead."""
    rng = np.random.default_rng(0)
    frames = []
    for i in range(4):
        n = 24
        th = np.sort(rng.uniform(0, 2 * np.pi, n))
        rows = np.stack([np.full(n, i), np.arange(n), 4 + 2 * np.cos(th), 5 + 1.5 * np.sin(th), np.full(n, float(i)), np.zeros(n)], 1)
        c = PyContour(i, i, rows, None, None, None, "Lumen")
        c.compute_centroid()
        frames.append(PyFrame(i, c.centroid, c, {}, PyContourPoint(i, 0, 9.0, 5.0, float(i), False) if i == 0 else None))
    g = PyGeometry(frames, "g")
    assert np.array_equal(PyGeometry.from_blob(g.to_blob(), "g").to_blob(), g.to_blob())
    s = g.frames[1].sort_frame_points().lumen
    rows = s.points_array()
    ang = np.arctan2(rows[:, 3] - rows[:, 3].mean(), rows[:, 2] - rows[:, 2].mean())
    k = int(np.argmax(rows[:, 3]))
    assert k == 0 and (np.diff(np.unwrap(ang)) > 0).all() and list(rows[:, 1]) == list(range(len(rows)))
    sm = g.smooth_frames()
    want = (g.frames[0].lumen.points_array()[:, 2] * 2 + g.frames[1].lumen.points_array()[:, 2])
    assert np.allclose(sm.frames[0].lumen.points_array()[:, 2], want / 3.0)
    mla, sten, length = g.get_summary()
    assert mla > 0 and 0 <= sten < 1 and length >= 0
    d = g.downsample(7)
    assert len(d.frames[0].lumen) == 7 and list(d.frames[0].lumen.points_array()[:, 1]) == [float(int(i * 24 / 7)) for i in range(7)]
    assert g.get_frame_at_z(2.2).id == 2 and g.get_frame_at_index(3).id == 3
    with pytest.raises(IndexError):
        g.get_frame_at_index(9)


def test_geometry_helpers_added_for_source_compatibility(capsys):
    """py_geometry.rs:98-100 (get_contours), :152-156 (sort_frame_points -> sort_frame_points_by_z), :267-273
    (center_to_contour); py_geometry_pair.rs:48-55 (repr) and :70-201 (get_summary + printed table)."""
    from multimodars import PyContourType, PyGeometryPair

    def geom(label, shift):
        frames = []
        for i in range(3):
            c = _ring(i, (1.0 + shift * i, 2.0 - shift * i, float(i)), 2.0 + 0.1 * i, 2.0 + 0.1 * i, n=12)
            frames.append(PyFrame(i, c.centroid, c, {}, None))
        return PyGeometry(frames, label)

    g = geom("dia", 0.3)
    assert [c.id for c in g.get_contours("Lumen")] == [0, 1, 2]
    centred = g.center_to_contour(PyContourType.Lumen)
    for f in centred.frames:
        assert abs(f.lumen.centroid[0] - 1.0) < 1e-12 and abs(f.lumen.centroid[1] - 2.0) < 1e-12
        assert abs(f.centroid[0] - 1.0) < 1e-12
    assert centred.frames[2].centroid[2] == 2.0 and g.frames[2].centroid[0] != 1.0      # a copy; z untouched
    # sort_frame_points_by_z: frame 0's highest-z lumen point (the LAST maximum) becomes index 0 everywhere
    g.frames[0].lumen.points[5].z = 9.0
    s = g.sort_frame_points()
    assert s.frames[0].lumen.points[0].z == 9.0
    assert [p.point_index for p in s.frames[1].lumen.points] == list(range(12))
    assert s.frames[1].lumen.points[0].x == g.frames[1].lumen.points[5].x
    pair = PyGeometryPair(g, geom("sys", 0.1), "dia - sys")
    assert repr(pair) == "GeometryPair dia - sys (diastolic: 3 frames, systolic: 3 frames)"
    (dia, sys_), table = pair.get_summary()
    assert dia == g.get_summary() and len(table) == 3 and len(table[0]) == 6
    assert table[1][0] == 1.0 and table[1][1] == g.frames[1].lumen.get_area() and table[2][5] == 2.0
    out = capsys.readouterr().out.splitlines()
    assert out[1].replace(" ", "") == "|id|area_dia|ellip_dia|area_sys|ellip_sys|z|" and out[0].startswith("+----+")
