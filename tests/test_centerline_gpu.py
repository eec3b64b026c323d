"""align_combined = three-point start + refine_alignment_hausdorff (align_algorithms.rs:339-451): every
(centerline index, angle) candidate is scored by the symmetric Hausdorff distance on the GPU as one batch of
sweep units. Checked (1) by recovering a planted (rotation, centerline index) whose cost is exactly 0 and
(2) against the ORACLE's align_combined (oracle/centerline_py.py: the pure-Python restatement of align.rs:166-283 and
align_algorithms.rs:339-451, pinned on the reference's Rust tests by tests/test_oracle_centerline.py)."""
import math

import numpy as np
import pytest

from multimodars import PyCenterline, PyCenterlinePoint, PyGeometry, align_combined, align_manual, align_three_point, get_context
from multimodars import _native as nat
from tests.test_centerline_cpu import centerline, newell, place, pullback
from scipy.spatial.transform import Rotation

pytestmark = pytest.mark.gpu


def hausdorff_xy(a, b):
    d2 = ((a[:, None, 0] - b[None, :, 0]) ** 2 + (a[:, None, 1] - b[None, :, 1]) ** 2)
    return max(math.sqrt(d2.min(1).max()), math.sqrt(d2.min(0).max()))


def lumen_cloud(g):
    return np.concatenate([f.lumen.points_array()[:, 2:5] for f in g.frames])


def targets_for(g, theta_deg, p0, d, ref_index, n):
    lum = g.frames[0].lumen
    pts = lum.points_array()[:, 2:5]
    c = np.array(lum.centroid)
    R = Rotation.from_rotvec(newell(pts, c) * math.radians(theta_deg)).as_matrix()
    placed = place(c + (pts - c) @ R.T, c, p0, d)
    return tuple(placed[ref_index]), tuple(placed[0]), tuple(placed[n // 2])


def test_combined_recovers_planted_rotation_and_index():
    # a tilted centerline: the cost reads (x, y) only (process_utils.rs:78-121), so a centerline along z could not
    # tell the indices apart. The candidates rotate the ALREADY PLACED frames in the x-y plane
    # (rotate_by_best_rotation on the aligned target, align_algorithms.rs:398-402), which for tilted frames is close
    # to, not exactly, a rotation about the tangent: the planted 6 degrees come back to within one grid step.
    # The stored tangents point against the direction of travel (towards +z, like the lumen normals): with tangents
    # along -z the placement turns every frame over (angle(normal, tangent) ~ 174 degrees) and the candidates'
    # re-sort (Geometry::rotate_geometry, geometry.rs:241-250) makes all non-zero angles cost the same.
    n, direction = 48, (0.1, -0.05, -1.0)
    g = pullback(n_frames=5, n=n, dz=1.0, ref_frame=0, ref_index=7, ry=1.5)
    p0 = np.array([3.0, -2.0, 25.0])
    cl = centerline(p0, direction, 60, 0.5)
    d = np.asarray(direction) / np.linalg.norm(direction)
    cl = PyCenterline([PyCenterlinePoint(q.contour_point, tuple(-d)) for q in cl.points])
    theta0 = 30.0
    s = float(np.mean(np.linalg.norm(np.diff(np.array([f.centroid for f in g.frames]), axis=0), axis=1)))
    main, ccw, cw = targets_for(g, theta0, p0 + 4.0 * s * d, -d, 7, n)   # three-point start: resampled index 4
    start, spacing, rot0 = align_three_point(cl, g, main, ccw, cw)
    assert abs(rot0 - theta0) < 1e-9 and abs(spacing - s) < 1e-12
    # plant: the cloud is the geometry placed one centerline point further with 6 more degrees
    planted, _, _ = align_manual(cl, g, theta0 + 6.0, tuple(p0 + 5.0 * s * d))
    cloud = lumen_cloud(planted)
    res, _, rot = align_combined(cl, g, main, ccw, cw, [tuple(p) for p in cloud], angle_step_deg=1.0,
                                 angle_range_deg=10.0, index_range=2)
    assert abs(rot - (theta0 + 6.0)) <= 1.0 + 1e-9, rot
    assert np.allclose(res.frames[0].centroid, p0 + 5.0 * s * d, atol=1e-9)      # the planted centerline index
    assert np.abs(lumen_cloud(res) - cloud).max() < 0.06                          # within one degree at r = 2.25
    st = get_context().process_stats()
    assert st["units"] > 0 and st["launches"] > 0   # the candidates went through the sweep kernels


def test_combined_matches_the_oracle():
    """mmrs_align_centerline method 2 (three-point start + refine_alignment_hausdorff with every candidate scored on the
    GPU) against oracle/centerline_py.py::align_combined, the pure-Python restatement of align.rs:166-283 /
    align_algorithms.rs:339-451 pinned on the reference's Rust tests: the same number of candidates scored, the same
    refined (angle, centerline index) — i.e. the same total rotation to 1e-12 — the same minimal Hausdorff distance and
    the same final geometry to 1e-9 (float tolerance: oracle/centerline_py.py header)."""
    from oracle import centerline_py as oc
    from oracle import oracle_py as ora
    from tests.test_centerline_cpu import _assert_same_geometry

    n, direction = 40, (0.0, 0.0, -1.0)
    g = pullback(n_frames=4, n=n, dz=1.0, ref_frame=0, ref_index=3, ry=1.4)
    p0 = np.array([-1.0, 4.0, 30.0])
    cl = centerline(p0, direction, 50, 0.25)
    d = np.asarray(direction)
    theta0 = 100.0
    s = float(np.mean(np.linalg.norm(np.diff(np.array([f.centroid for f in g.frames]), axis=0), axis=1)))
    main, ccw, cw = targets_for(g, theta0, p0 + 3.0 * s * d, d, 3, n)
    rng = np.random.default_rng(5)
    truth, _, _ = align_manual(cl, g, theta0 - 3.3, tuple(p0 + 2.0 * s * d))
    cloud = lumen_cloud(truth) + rng.normal(0, 0.02, (4 * n, 3))
    cloud = np.concatenate([cloud, rng.uniform(-30, 30, (50, 3)) + p0])   # outliers, mostly outside the 5 mm box
    step, rng_deg, idx_range = 1.0, 5.0, 2
    ctx = get_context()
    blobs, spacing, rot_rad, (best_h, n_cand) = nat.align_centerline(
        ctx, 2, cl._rows(), [g.to_blob()], main_ref_pt=main, ccw_ref_pt=ccw, cw_ref_pt=cw,
        angle_step_rad=math.radians(step), points=cloud, angle_range_rad=math.radians(rng_deg), index_range=idx_range)
    want, wspacing, wtotal, info = oc.align_combined(oc.centerline_from_rows(cl._rows()), [ora.decode_geometry(g.to_blob())],
                                                     main, ccw, cw, cloud, math.radians(step), math.radians(rng_deg), idx_range)
    assert spacing == wspacing
    assert n_cand == info["scored"] and info["scored"] == (2 * idx_range + 1) * 11
    assert abs(best_h - info["min_hausdorff"]) < 1e-9, (best_h, info)
    assert abs(rot_rad - wtotal) < 1e-12, (rot_rad, wtotal, info)
    # the planted -3.3 degrees come back to the grid; the centerline runs along z and the cost reads (x, y) only, so the
    # index candidates tie and the FIRST one (strict `<`, align_algorithms.rs:433) wins — the final geometry's z tells which
    assert abs(info["delta"] - math.radians(-3.0)) < 1e-12 and info["refined_idx"] == 1
    _assert_same_geometry(ora.decode_geometry(blobs[0]), want[0])
