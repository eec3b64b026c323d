"""Branch bookkeeping of PyCenterline (multimodars/_centerline.py) against the reference's own pins:
the Rust unit tests of src/types/native/centerline.rs:1023-1524 (same constructions, same expected values) and the
Python tests of tests/test_intravascular.py:249-343 on the example RCA centerline (tests/golden/
centerline_rca_short.npz, written by tests/golden/make_centerline_golden.py)."""
import math
from pathlib import Path

import numpy as np
import pytest

from multimodars import PyCenterline, PyCenterlinePoint, PyContourPoint, numpy_to_centerline

GOLDEN = Path(__file__).resolve().parent / "golden"


def multi(*branches):
    """centerline.rs:944-970 make_multi_branch: explicit branch ids, zero tangents."""
    pts, starts = [], []
    for bid, coords in enumerate(branches):
        starts.append(len(pts))
        for x, y, z in coords:
            i = len(pts)
            pts.append(PyCenterlinePoint(PyContourPoint(i, i, float(x), float(y), float(z), False), (0.0, 0.0, 0.0), bid))
    cl = PyCenterline(pts)
    cl.branch_start_indices = starts
    return cl


def line(coords):
    """centerline.rs:972-986 cl_from_coords."""
    return PyCenterline.from_contour_points([PyContourPoint(i, i, float(x), float(y), float(z), False)
                                             for i, (x, y, z) in enumerate(coords)])


def branches(cl):
    s = cl.branch_start_indices
    return [cl.points[a:b] for a, b in zip(s, s[1:] + [len(cl.points)])]


def xs(points):
    return [p.contour_point.x for p in points]


def well_formed(cl):
    for i, p in enumerate(cl.points):
        assert p.contour_point.point_index == i
    for bid, br in enumerate(branches(cl)):
        assert all(p.branch_id == bid for p in br)


# ---- find_sharp_angles (centerline.rs:1023-1066) -------------------------------------------------------------------
def test_sharp_angles():
    assert line([(i, 0, 0) for i in range(5)]).find_sharp_angles(0, 0.0) == []
    v = line([(0, 0, 0), (1, 0, 0), (2, 0, 0), (3, 0, 0), (2.5, 0.5, 0), (2, 1, 0)])
    assert v.find_sharp_angles(0, 0.0) == [3]
    assert v.find_sharp_angles(0, 0.8) == []
    assert v.find_sharp_angles(5, 0.0) == []
    cl = multi([(0, 0, 0), (0, 0, 1), (0, 0, 2)],
               [(0, 0, 0), (1, 0, 0), (2, 0, 0), (3, 0, 0), (2.5, 0.5, 0), (2, 1, 0)])
    assert cl.find_sharp_angles(1, 0.0) == [6]          # global index, not the position inside the branch


# ---- split / merge (centerline.rs:1069-1208) -----------------------------------------------------------------------
def test_split_longer_piece_becomes_main():
    src = line([(i, 0, 0) for i in range(9)])
    cl = src.split_branch(0, 3)
    assert len(src.points) == 9 and src.branch_start_indices == [0]      # input untouched
    assert cl.branch_start_indices == [0, 6] and len(cl.points) == 10
    assert [p.branch_id for p in cl.points] == [0] * 6 + [1] * 4
    well_formed(cl)
    assert xs(branches(cl)[0]) == [3, 4, 5, 6, 7, 8] and xs(branches(cl)[1]) == [0, 1, 2, 3]


def test_split_equal_pieces_keep_order():
    assert line([(i, 0, 0) for i in range(5)]).split_branch(0, 2).branch_start_indices == [0, 3]


def test_split_takes_a_global_index():
    cl = multi([(i, 0, 0) for i in range(5)], [(10 + i, 0, 0) for i in range(5)]).split_branch(1, 7)
    b = branches(cl)
    assert [len(x) for x in b] == [5, 3, 3]
    assert b[1][0].contour_point.x == 10.0 and b[2][0].contour_point.x == 12.0


def test_split_resorts_side_branches():
    cl = multi([(i, 0, 0) for i in range(10)], [(i, 1, 0) for i in range(6)], [(i, 2, 0) for i in range(2)])
    assert [len(x) for x in branches(cl.split_branch(1, 13))] == [10, 4, 3, 2]


@pytest.mark.parametrize("bid,idx", [(3, 1), (0, 0), (0, 4), (0, 99), (1, 2)])
def test_split_out_of_range_is_a_no_op(bid, idx):
    src = multi([(i, 0, 0) for i in range(5)], [(i, 1, 0) for i in range(3)])
    cl = src.split_branch(bid, idx) if not (bid == 1 and idx == 2) else src.split_branch(1, 2)
    assert cl.branch_start_indices == src.branch_start_indices
    assert [p.contour_point.x for p in cl.points] == [p.contour_point.x for p in src.points]


def test_merge_back_after_split():
    cl = line([(i, 0, 0) for i in range(5)]).split_branch(0, 2).merge_branches(0, 1)
    assert cl.branch_start_indices == [0] and len(cl.points) == 6
    well_formed(cl)
    assert xs(cl.points) == [0, 1, 2, 2, 3, 4]


def test_merged_side_branches_can_become_main():
    cl = multi([(i, 0, 0) for i in range(5)], [(0, 1, 0), (1, 1, 0), (2, 1, 0), (3, 1, 0)],
               [(3, 1, 0), (4, 1, 0), (5, 1, 0), (6, 1, 0)]).merge_branches(1, 2)
    assert [len(x) for x in branches(cl)] == [8, 5]
    assert xs(branches(cl)[0]) == [0, 1, 2, 3, 3, 4, 5, 6]


def test_merge_orientations():
    # last-last: the second branch is appended reversed; first-first: the second branch reversed goes in front
    a = multi([(0, 0, 0), (1, 0, 0), (2, 0, 0)], [(5, 0, 0), (4, 0, 0), (2.5, 0, 0)]).merge_branches(0, 1)
    assert xs(a.points) == [0, 1, 2, 2.5, 4, 5]
    b = multi([(2, 0, 0), (1, 0, 0), (0, 0, 0)], [(2.5, 0, 0), (4, 0, 0), (5, 0, 0)]).merge_branches(1, 0)
    assert xs(b.points) == [5, 4, 2.5, 2, 1, 0]
    c = multi([(2, 0, 0), (1, 0, 0), (0, 0, 0)], [(5, 0, 0), (4, 0, 0), (2.5, 0, 0)]).merge_branches(0, 1)
    assert xs(c.points) == [5, 4, 2.5, 2, 1, 0]
    same = multi([(0, 0, 0), (1, 0, 0)], [(5, 0, 0), (6, 0, 0)])
    assert xs(same.merge_branches(0, 0).points) == xs(same.points) and xs(same.merge_branches(0, 7).points) == xs(same.points)


def test_get_branch():
    cl = multi([(i, 0, 0) for i in range(4)], [(i, 1, 0) for i in range(3)])
    b = cl.get_branch(1)
    assert len(b.points) == 3 and b.branch_start_indices == [0] and all(p.branch_id == 0 for p in b.points)
    assert cl.points[5].branch_id == 1
    with pytest.raises(ValueError, match="branch_id 4 not found in centerline"):
        cl.get_branch(4)


# ---- tangents (centerline.rs:1211-1242) ----------------------------------------------------------------------------
def test_tangents():
    cl = line([(0, 0, 0), (1, 0, 0), (2, 0, 0)])
    assert [p.tangent for p in cl.points] == [(1.0, 0.0, 0.0)] * 3
    # inside branches only: a branch's last point repeats its predecessor, a lone point gets zeros
    re = multi([(0, 0, 0), (0, 3, 0), (0, 3, 4)], [(9, 9, 9)]).orient_to_reference(line([(0, 0, 0), (0, -1, 0)]))
    assert [p.tangent for p in re.points] == [(0.0, 1.0, 0.0), (0.0, 0.0, 1.0), (0.0, 0.0, 1.0), (0.0, 0.0, 0.0)]


# ---- overlap / trimming (centerline.rs:1245-1320) ------------------------------------------------------------------
def test_overlap_prefix_is_trimmed_to_the_junction():
    cl = multi([(i, 0, 0) for i in range(5)], [(0, 0, 0), (1, 0, 0), (2, 0, 0), (2, 1.5, 0), (2, 3, 0)])
    b = branches(cl.remove_branch_overlap())
    assert [len(x) for x in b] == [5, 3]
    assert (b[1][0].contour_point.x, b[1][0].contour_point.y) == (2.0, 0.0)


def test_fully_overlapping_branch_is_dropped():
    cl = multi([(0, 0, 0), (1, 0, 0), (2, 0, 0)], [(0, 0, 0), (1, 0, 0)]).remove_branch_overlap()
    assert cl.branch_start_indices == [0]


def test_no_overlap_leaves_branch():
    cl = multi([(0, 0, 0), (1, 0, 0), (2, 0, 0)], [(0, 5, 0), (0, 6, 0), (0, 7, 0)]).remove_branch_overlap()
    assert [len(x) for x in branches(cl)] == [3, 3]


def test_overlap_against_an_earlier_side_branch():
    # branch 2 starts on branch 1 (not on the main vessel): the growing set of known points catches it
    cl = multi([(i, 0, 0) for i in range(6)], [(2, 0, 0), (2, 2, 0), (2, 3, 0), (2, 4, 0)],
               [(2, 3, 0), (2, 4, 0), (4, 4, 0), (6, 4, 0)]).remove_branch_overlap()
    b = branches(cl)
    assert [len(x) for x in b] == [6, 4, 3]
    assert (b[2][0].contour_point.x, b[2][0].contour_point.y) == (2.0, 4.0)


def test_trim_start():
    cl = multi([(i, 0, 0) for i in range(6)]).trim_start(3.0)
    assert cl.branch_start_indices == [0] and xs(cl.points) == [3, 4, 5]
    well_formed(cl)
    assert xs(multi([(i, 0, 0) for i in range(6)]).trim_start(0.0).points) == [0, 1, 2, 3, 4, 5]
    assert xs(multi([(i, 0, 0) for i in range(6)]).trim_start(0.5).points) == [0, 1, 2, 3, 4, 5]


# ---- smoothing (centerline.rs:1323-1377) ---------------------------------------------------------------------------
def test_smooth():
    straight = line([(i, 0, 0) for i in range(20)])
    sm = straight.smooth(3.0)
    for a, b in zip(straight.points, sm.points):
        assert abs(a.contour_point.x - b.contour_point.x) < 1e-10 and b.contour_point.y == 0.0 and b.contour_point.z == 0.0
    pts = [(i, 0, 0) for i in range(15)]
    pts[7] = (7, 5, 0)
    y = line(pts).smooth(2.0).points[7].contour_point.y
    # window of +-6 points, weights exp(-d^2/8): 5 / sum
    w = sum(math.exp(-0.5 * d * d / 4.0) for d in range(-6, 7))
    assert 0.0 < y < 5.0 and y == pytest.approx(5.0 / w, rel=1e-14)
    pts = [(i, 0, 0) for i in range(20)]
    pts[10] = (10, 3, 0)
    for p in line(pts).smooth(2.0).points:
        n = math.sqrt(sum(c * c for c in p.tangent))
        assert abs(n - 1.0) < 1e-10 or n < 1e-12
    src = line([(i, 0, 0) for i in range(10)])
    same = src.smooth(0.0)
    assert xs(same.points) == xs(src.points) and [p.tangent for p in same.points] == [p.tangent for p in src.points]


def test_smooth_does_not_bleed_across_branches():
    cl = multi([(i, 0, 0) for i in range(8)], [(i, 10, 0) for i in range(8)]).smooth(2.0)
    assert all(p.contour_point.y == 0.0 for p in cl.points[:8]) and all(abs(p.contour_point.y - 10.0) < 1e-12 for p in cl.points[8:])


# ---- resampling (centerline.rs:1380-1403) --------------------------------------------------------------------------
def test_resample():
    cl = line([(0, 0, 0), (10, 0, 0)]).resample(2.5)
    assert xs(cl.points) == [0.0, 2.5, 5.0, 7.5, 10.0]
    assert [p.contour_point.frame_index for p in cl.points] == [0, 1, 2, 3, 4]
    two = multi([(0, 0, 0), (10, 0, 0)], [(10, 0, 0), (10, 5, 0)]).resample(2.0)
    b = branches(two)
    assert len(b) == 2 and [len(x) for x in b] == [6, 4]
    well_formed(two)
    assert b[1][0].contour_point.y == 0.0 and b[1][-1].contour_point.y == 5.0
    assert [p.contour_point.frame_index for p in b[1]] == [0, 1, 2, 3]    # per branch, like the reference
    r = multi([(0, 0, 0), (4, 0, 0)])
    r.points[0].radius, r.points[1].radius = 1.0, 3.0
    assert [p.radius for p in r.resample(1.0).points] == [1.0, 1.5, 2.0, 2.5, 3.0]
    assert xs(line([(0, 0, 0), (10, 0, 0)]).resample(0.0).points) == [0.0, 10.0]


# ---- orientation (centerline.rs:1406-1523) -------------------------------------------------------------------------
def test_orient_by_max_z():
    cl = multi([(0, 0, 0), (0, 0, 1), (0, 0, 2)], [(0, 0, 2), (5, 0, 2)]).orient_by_max_z()
    b = branches(cl)
    assert [p.contour_point.z for p in b[0]] == [2.0, 1.0, 0.0] and xs(b[1]) == [0.0, 5.0]
    well_formed(cl)
    assert line([(0, 0, 2), (0, 0, 1), (0, 0, 0)]).orient_by_max_z().points[0].contour_point.z == 2.0
    b = branches(multi([(0, 0, 0), (0, 0, 1), (0, 0, 2)], [(5, 0, 2), (0, 0, 2)]).orient_by_max_z())
    assert b[0][0].contour_point.z == 2.0 and xs(b[1]) == [0.0, 5.0]
    # equal maxima: Iterator::max_by keeps the last one, so a flat branch IS reversed
    assert xs(line([(0, 0, 1), (1, 0, 1), (2, 0, 1)]).orient_by_max_z().points) == [2.0, 1.0, 0.0]


def test_orient_to_reference():
    ref = line([(10, 1, 0), (20, 1, 0)])
    b = branches(multi([(0, 0, 0), (5, 0, 0), (10, 0, 0)], [(0, 0, 0), (0, 5, 0)]).orient_to_reference(ref))
    assert xs(b[0]) == [10.0, 5.0, 0.0] and xs(b[1]) == [0.0, 0.0] and b[1][1].contour_point.y == 5.0
    assert line([(10, 0, 0), (5, 0, 0), (0, 0, 0)]).orient_to_reference(ref).points[0].contour_point.x == 10.0
    # only the reference's branch 0 counts
    ref2 = multi([(0, 1, 0), (1, 1, 0)], [(10, 1, 0), (11, 1, 0)])
    assert line([(0, 0, 0), (5, 0, 0), (10, 0, 0)]).orient_to_reference(ref2).points[0].contour_point.x == 0.0
    b = branches(multi([(0, 0, 0), (5, 0, 0), (10, 0, 0)], [(10, 5, 0), (0, 5, 0)])
                 .orient_to_reference(line([(0, 1, 0), (-10, 1, 0)])))
    assert b[0][0].contour_point.x == 0.0 and xs(b[1]) == [0.0, 10.0]
    ref3 = multi([(0, 1, 0), (1, 1, 0)], [(10, 5, 0), (11, 5, 0)])
    b = branches(multi([(0, 0, 0), (5, 0, 0), (10, 0, 0)], [(0, 0, 0), (10, 5, 0)]).orient_to_reference(ref3))
    assert b[1][0].contour_point.x == 0.0


# ---- calculate_branches ---------------------------------------------------------------------------------------------
def test_calculate_branches_two_segments():
    # main 0..49 along x, a 12-point side branch hanging off x = 20 stored AFTER it with a jump, plus a 2-point
    # artefact far away (2 jumps in 63 gaps, so the 95th-percentile spacing is still 1.0)
    main = [(float(i), 0.0, 0.0) for i in range(50)]
    side = [(20.0, float(j), 0.0) for j in range(12, 0, -1)]
    noise = [(500.0, 500.0, 0.0), (500.5, 500.0, 0.0)]
    cl = numpy_to_centerline(np.array(main + side + noise)).calculate_branches(1.0)
    b = branches(cl)
    assert [len(x) for x in b] == [50, 12]                          # the artefact is dropped
    well_formed(cl)
    assert xs(b[0]) == [float(i) for i in range(50)]                # second BFS's far end (x = 0) traced back to x = 49
    assert [p.contour_point.y for p in b[1]] == [float(j) for j in range(12, 0, -1)]
    assert b[1][-1].tangent == b[1][-2].tangent
    empty = PyCenterline([]).calculate_branches(1.0)
    assert empty.points == [] and empty.branch_start_indices == []


@pytest.fixture(scope="module")
def rca():
    return numpy_to_centerline(np.load(GOLDEN / "centerline_rca_short.npz")["xyz"])


def test_rca_branches(rca):
    """tests/test_intravascular.py:275-343 (the docstring there says 510 main points; the code drops the 2-point
    artefact, centerline.rs:133-136, so 508 remain — the tests only ask for the branch structure)."""
    before = [p.branch_id for p in rca.points]
    cl = rca.calculate_branches(2.0)
    assert [p.branch_id for p in rca.points] == before
    assert len(cl.branch_start_indices) == 4
    sizes = [len(b) for b in branches(cl)]
    assert sizes == [508, 131, 116, 31] and len(cl.points) == 786
    well_formed(cl)
    main = {p.contour_point.frame_index for p in cl.points if p.branch_id == 0}
    assert set(range(463, 639)) <= main and set(range(132, 463)) <= main
    assert all(p.branch_id != 0 for p in cl.points if p.contour_point.frame_index <= 130)
    assert {p.branch_id for p in cl.points if 639 <= p.contour_point.frame_index <= 669} == {3}
    assert {p.branch_id for p in cl.points if 672 <= p.contour_point.frame_index <= 787} == {2}
    assert not any(p.contour_point.frame_index in (670, 671) for p in cl.points)
    assert repr(cl).startswith("Centerline(len=786, spacing=") and repr(cl).endswith(" mm, branches=4)")


def test_rca_pipeline_keeps_invariants(rca):
    cl = rca.calculate_branches(2.0).orient_by_max_z().remove_branch_overlap().resample(1.0).smooth(2.0)
    well_formed(cl)
    for br in branches(cl):
        d = np.diff(np.array([(p.contour_point.x, p.contour_point.y, p.contour_point.z) for p in br]), axis=0)
        assert len(br) >= 2 and np.all(np.linalg.norm(d, axis=1) < 1.0 + 1e-9)


def test_load_and_prepare(rca, tmp_path, capsys):
    """multimodars/ccta/centerline_prep.py:10-140."""
    import multimodars as mm
    xyz = np.load(GOLDEN / "centerline_rca_short.npz")["xyz"]
    assert mm.load_centerline(rca, "RCA") is rca
    from_arr = mm.load_centerline(xyz, "RCA")
    p = tmp_path / "cl.csv"
    np.savetxt(p, xyz, delimiter=",", fmt="%.17g")
    from_csv = mm.load_centerline(p, "RCA")
    assert from_csv.points_as_tuples() == from_arr.points_as_tuples() == rca.points_as_tuples()
    out = capsys.readouterr().out.splitlines()
    assert out == ["Using provided RCA centerline: 788 points", "Using provided RCA centerline: 788 points",
                   "Loaded RCA centerline: 788 points"]
    with pytest.raises(Exception):
        mm.load_centerline(tmp_path / "missing.csv", "LCA")
    assert "Error reading LCA centerline from" in capsys.readouterr().out

    aorta = line([(0, 0, 100 + 0.5 * i) for i in range(40)])
    ao = mm.prepare_centerline(aorta)                        # no reference: no branch search, highest z first
    assert ao.branch_start_indices == [0] and ao.points[0].contour_point.z > ao.points[-1].contour_point.z
    want = rca.calculate_branches(2.0).remove_branch_overlap().trim_start(3.0).resample(0.5).orient_to_reference(ao).smooth(2.5)
    got = mm.prepare_centerline(rca, ref_centerline=ao, spacing_mm=0.5, rm_start_mm=3.0)
    assert got.points_as_tuples() == want.points_as_tuples() and got.branch_start_indices == want.branch_start_indices
    assert [p.tangent for p in got.points] == [p.tangent for p in want.points]
    # an already branched centerline is not searched again; sigma 0 and no spacing skip those steps
    pre = rca.calculate_branches(2.0)
    kept = mm.prepare_centerline(pre, ref_centerline=ao, smooth_sigma=0.0)
    assert kept.points_as_tuples() == pre.remove_branch_overlap().orient_to_reference(ao).points_as_tuples()


def test_reference_import_paths():
    import multimodars as mm
    from multimodars.ccta.centerline_prep import load_centerline, prepare_centerline
    from multimodars.ccta import load_centerline as lc2
    assert load_centerline is mm.load_centerline is lc2 and prepare_centerline is mm.prepare_centerline
    import multimodars.ccta as ccta
    with pytest.raises(AttributeError, match="not part of this build"):
        ccta.stitching
