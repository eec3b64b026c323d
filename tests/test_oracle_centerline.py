"""Pins oracle/centerline_py.py (the restatement of the centerline-alignment path) on the
reference's own unit tests: centerline_align/preprocessing.rs:291-605 and align_algorithms.rs:573-935, value for value."""
import math

import numpy as np

from oracle import centerline_py as oc


def pts(rows):
    """[(point_index, x, y, z)] -> (n, 6) rows [frame_index, point_index, x, y, z, aortic]."""
    return np.array([[0.0, float(i), float(x), float(y), float(z), 0.0] for i, x, y, z in rows])


def contour(rows, centroid=None):
    return dict(kind=0, id=0, original_frame=0, centroid=centroid, aortic_thickness=None, pulmonary_thickness=None,
                points=pts(rows))


def clp(x, y, z, tangent=(0.0, 0.0, 1.0)):   # create_test_centerline_point, align_algorithms.rs:595-610
    return dict(p=[x, y, z], t=list(tangent), branch=0, radius=0.0)


# ---- preprocessing.rs:291-605 ---------------------------------------------------------------------------------------
def test_ensure_descending_z_via_preprocess():   # :296-360
    frames = [dict(centroid=(0.0, 0.0, float(k))) for k in range(2)]
    up = [clp(0, 0, 0.0), clp(0, 0, 1.0)]
    out, _ = oc.preprocess_centerline(up, frames)
    assert out[0]["p"][2] == 1.0 and out[-1]["p"][2] == 0.0
    down = [clp(0, 0, 1.0), clp(0, 0, 0.0)]
    out, _ = oc.preprocess_centerline(down, frames)
    assert out[0]["p"][2] == 1.0 and out[-1]["p"][2] == 0.0


def test_calculate_mean_spacing():   # :362-460: centroids (0,0,0), (3,4,0), (6,8,0) -> Some(5.0); one frame -> None
    frames = [dict(centroid=(0.0, 0.0, 0.0)), dict(centroid=(3.0, 4.0, 0.0)), dict(centroid=(6.0, 8.0, 0.0))]
    assert oc.calculate_mean_spacing(frames) == 5.0
    assert oc.calculate_mean_spacing([dict(centroid=(1.0, 2.0, 3.0))]) is None


def test_cumulative_arc_length_and_decide_spacing():   # :463-528
    cl = [clp(0, 0, float(k)) for k in range(4)]
    cum = oc.cumulative_arc_length(cl)
    assert cum == [0.0, 1.0, 2.0, 3.0]
    assert oc.decide_spacing(None, cum[-1], len(cl) - 1) == 1.0


def test_build_samples_and_interpolate():   # :530-605
    cl = [clp(0, 0, float(k)) for k in range(4)]
    cum = oc.cumulative_arc_length(cl)
    s = oc.build_samples(3.0, 0.75)
    assert len(s) >= 2 and s[0] == 0.0 and s[-1] == 3.0 and s == [0.0, 0.75, 1.5, 2.25, 3.0]
    p = oc.interpolate_centerline_at_s(cl, cum, 1.5)
    assert abs(p["p"][2] - 1.5) < 1e-12 and abs(p["t"][2] - 1.0) < 1e-12 and abs(p["radius"]) < 1e-12


# ---- align_algorithms.rs:573-935 --------------------------------------------------------------------------------------
def test_frame_transformation_apply_to_point():   # :598-627: translate, then rotate 90 deg about z through the pivot
    tr = dict(translation=[1.0, 2.0, 3.0], rotation=oc.IDENTITY, pivot=[0.0, 0.0, 0.0])
    assert oc.apply_to_xyz(tr, 1.0, 1.0, 1.0) == (2.0, 3.0, 4.0)
    rot = oc.rot_from_axis_angle([0.0, 0.0, 1.0], math.pi / 2)   # Rotation3::from_axis_angle(&Vector3::z_axis(), FRAC_PI_2)
    tr = dict(translation=[0.0, 0.0, 0.0], rotation=rot, pivot=[0.0, 0.0, 0.0])
    x, y, z = oc.apply_to_xyz(tr, 1.0, 0.0, 0.0)
    assert abs(x) < 1e-12 and abs(y - 1.0) < 1e-12 and abs(z) < 1e-12


def test_align_frame():   # :628-680: centroid (0,0,0) of the square goes to (10,10,10); pivot = the centerline point
    c = contour([(0, -1, -1, 0), (1, 1, -1, 0), (2, 1, 1, 0), (3, -1, 1, 0)])
    tr = oc.align_frame(c, clp(10.0, 10.0, 10.0))
    assert all(abs(t - 10.0) < 1e-12 for t in tr["translation"]) and all(abs(p - 10.0) < 1e-12 for p in tr["pivot"])
    assert tr["rotation"] == oc.IDENTITY     # the square's Newell normal is +z, parallel to the default tangent


def test_apply_transformation_to_contour():   # :681-725
    c = contour([(0, 0, 0, 0), (1, 1, 0, 0)], centroid=(0.5, 0.0, 0.0))
    oc.apply_transformation_to_contour(c, dict(translation=[2.0, 3.0, 4.0], rotation=oc.IDENTITY, pivot=[0.0, 0.0, 0.0]))
    assert c["points"][0][2:5].tolist() == [2.0, 3.0, 4.0] and c["points"][1][2:5].tolist() == [3.0, 3.0, 4.0]
    assert c["centroid"] == (2.5, 3.0, 4.0)


def test_calculate_normal():   # :726-766: unit length, along z for a triangle in the x-y plane
    n = oc.calculate_normal(pts([(0, 0, 0, 0), (1, 1, 0, 0), (2, 0, 1, 0)]), (0.0, 0.0, 0.0))
    assert abs(oc.v_norm(n) - 1.0) < 1e-12 and n[:2] == [0.0, 0.0] and n[2] == 1.0
    assert oc.calculate_normal(pts([(0, 0, 0, 0), (1, 1, 0, 0)]), (0.0, 0.0, 0.0)) == [0.0, 0.0, 1.0]   # < 3 points


def test_rotate_contour_around_centroid():   # :767-821: (1,0,0) -> (0,1,0) under 90 degrees about the normal (+z)
    c = contour([(0, 1, 0, 0), (1, 0, 1, 0), (2, -1, 0, 0), (3, 0, -1, 0)], centroid=(0.0, 0.0, 0.0))
    oc.rotate_contour_around_centroid(c, math.pi / 2)
    p = c["points"][0]
    assert abs(p[2]) < 1e-6 and abs(p[3] - 1.0) < 1e-6 and abs(p[4]) < 1e-6


def test_get_transformations():   # :822-876: one frame, two centerline points, reference = the first -> one transformation
    frame = dict(id=0, centroid=(0.5, 0.0, 0.0), reference_point=None,
                 contours={0: contour([(0, 0, 0, 0), (1, 1, 0, 0)])})
    trs = oc.get_transformations([frame], [clp(10.0, 10.0, 10.0), clp(11.0, 10.0, 10.0)], (10.0, 10.0, 10.0))
    assert len(trs) == 1 and trs[0]["pivot"] == [10.0, 10.0, 10.0]


def test_best_rotation_three_point_simple_case():   # :877-935: targets at the current positions -> |best| < one step
    rows = [(i, math.cos(i * math.pi / 4), math.sin(i * math.pi / 4), 0.0) for i in range(8)]
    c = contour(rows, centroid=(0.0, 0.0, 0.0))
    ref = np.array([0.0, 0.0, 1.0, 0.0, 0.0, 0.0])
    step = math.pi / 8
    best = oc.best_rotation_three_point(c, ref, (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (-1.0, 0.0, 0.0), step,
                                        clp(0.0, 0.0, 0.0))
    assert abs(best) < step + 1e-6
    # worked by hand: the main and the counter-clockwise landmark are BOTH point 0 here (reference index 0), so the sum of
    # squared errors is 2 at angle 0, 2 (2 - 2 cos 22.5) + (2 - 2 sin 22.5) = 1.539 at one step, 1.757 at two: one step wins
    assert best == step


# ---- pieces shared with the C++ oracle: the two restatements must agree bit for bit ----------------------------------------
def test_shared_pieces_agree_with_the_cpp_oracle():
    from oracle import oracle_py as ora

    rng = np.random.default_rng(3)
    a = rng.normal(0, 1, (40, 2))
    b = rng.normal(0, 1, (55, 2))
    assert oc.hausdorff_xy([tuple(p) for p in a], [tuple(p) for p in b]) == ora.hausdorff(a, b)
    assert [int(i) for i in ora.downsample_indices(501, 37)] == [int(r) for r in oc.downsample(list(range(501)), 37)]
