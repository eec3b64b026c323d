"""align_combined = three-point start + refine_alignment_hausdorff (align_algorithms.rs:339-451): every
(centerline index, angle) candidate is scored by the symmetric Hausdorff distance on the GPU as one batch of
sweep units. Checked (1) by recovering a planted (rotation, centerline index) whose cost is exactly 0 and
(2) against a numpy brute force over the same candidate loop, each candidate re-placed with align_manual
(the same rotate_by_best_rotation + apply_transformations the reference's loop calls) and scored with the
oracle-style f64 max-min in (x, y) (process_utils.rs:78-121 ignores z)."""
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


def test_combined_matches_numpy_brute_force():
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
    # numpy restatement of the candidate loop (align_algorithms.rs:373-441) on the three-point start
    start, _, rot0 = align_three_point(cl, g, main, ccw, cw)
    assert abs(rot0 - theta0) < 1e-9
    best = (float("inf"), None, None)
    count = 0
    for delta in range(-idx_range, idx_range + 1):
        cur = 3 + delta
        ref_pt = p0 + cur * spacing * d
        seg_end = p0 + (cur + 4 - 1) * spacing * d
        lo, hi = np.minimum(ref_pt, seg_end) - 5.0, np.maximum(ref_pt, seg_end) + 5.0
        filt = cloud[((cloud >= lo) & (cloud <= hi)).all(1)]
        angle = -math.radians(rng_deg)
        while angle <= math.radians(rng_deg):
            cand, _, _ = align_manual(cl, start, math.degrees(angle), tuple(ref_pt))
            ratio = len(filt) / (n * 4)
            nd = min(max(math.ceil(ratio * n), 1), n)
            pts = []
            for f in cand.frames:
                rows = f.lumen.points_array()[:, 2:5]
                pick = range(n) if nd >= n else [int(i * (n / nd)) for i in range(nd)]
                pts.append(rows[list(pick)])
            h = hausdorff_xy(filt, np.concatenate(pts))
            count += 1
            if h < best[0]:
                best = (h, angle, cur)
            angle += math.radians(step)
    assert n_cand == count
    assert abs(best_h - best[0]) < 1e-9, (best_h, best)
    assert abs(rot_rad - (math.radians(theta0) + best[1])) < 1e-9
    res = PyGeometry.from_blob(blobs[0], "g")
    want, _, _ = align_manual(cl, g, math.degrees(rot_rad), tuple(p0 + best[2] * spacing * d))
    assert np.allclose(lumen_cloud(res), lumen_cloud(want), atol=1e-9)
