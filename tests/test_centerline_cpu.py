"""Centerline alignment (align_three_point / align_manual, mmrs_align_centerline methods 0 and 1) against an
independent numpy / scipy restatement of centerline_align/{preprocessing,align_algorithms,align}.rs and the
known answers of the reference's own unit tests (preprocessing.rs:291-605: mean spacing 5.0 for centroids
(0,0,0),(3,4,0),(6,8,0); resampling of a straight centerline; align_algorithms.rs:634-935: translation onto the
centerline point, identity rotation for parallel normals). Host-only (f64, no Hausdorff scoring): no GPU needed;
the Hausdorff-scored align_combined is covered by tests/test_centerline_gpu.py."""
import math

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from multimodars import (PyCenterline, PyCenterlinePoint, PyContour, PyContourPoint, PyFrame, PyGeometry, PyGeometryPair,
                         align_manual, align_three_point)
from multimodars import _native as nat


def ring_frame(fid, centre, r, n, ry=None, ref_index=None, extras=False):
    ry = r if ry is None else ry
    # sort_contour_points order (contour.rs:368-405): ascending atan2 about the mean, highest-y point first
    ang = [math.pi / 2 + 2 * math.pi * i / n for i in range(n)]
    pts = [PyContourPoint(fid, i, centre[0] + r * math.cos(a), centre[1] + ry * math.sin(a), centre[2], False)
           for i, a in enumerate(ang)]
    c = (sum(p.x for p in pts) / n, sum(p.y for p in pts) / n, sum(p.z for p in pts) / n)
    lum = PyContour(fid, fid, pts, c, None, None, "Lumen")
    ex = {}
    if extras:
        cp = [PyContourPoint(fid, i, c[0] + 0.4 * math.cos(a), c[1] + 0.4 * math.sin(a), centre[2], False)
              for i, a in enumerate(ang[::4])]
        ex["Catheter"] = PyContour(fid, fid, cp, c, None, None, "Catheter")
    ref = None
    if ref_index is not None:
        p = pts[ref_index]
        ref = PyContourPoint(fid, ref_index, p.x, p.y, p.z, False)
    return PyFrame(fid, c, lum, ex, ref)


def pullback(n_frames=6, n=48, dz=1.0, label="g", ref_frame=0, ref_index=5, extras=False, ry=None):
    frames = [ring_frame(f, (4.5 + 0.05 * f, 4.5 - 0.03 * f, f * dz), 2.0 + 0.05 * f, n, ry=ry,
                         ref_index=ref_index if f == ref_frame else None, extras=extras) for f in range(n_frames)]
    return PyGeometry(frames, label)


def centerline(p0, direction, n, step):
    d = np.asarray(direction, float) / np.linalg.norm(direction)
    pts = [PyContourPoint(i, i, *(np.asarray(p0) + i * step * d), False) for i in range(n)]
    return PyCenterline.from_contour_points(pts)


def newell(points, c):
    nn = np.zeros(3)
    m = len(points)
    for i in range(m):
        a, b = points[i] - c, points[(i + 1) % m] - c
        nn += np.cross(a, b)
    return nn / np.linalg.norm(nn)


def place(points, centroid, cl_pt, tangent):
    """align_frame + apply_to_point (align_algorithms.rs:66-181) with scipy's rotation."""
    nrm = newell(points, centroid)
    ang = math.acos(max(-1.0, min(1.0, float(nrm @ tangent))))
    axis = np.cross(nrm, tangent)
    if abs(ang) < 1e-6 or np.linalg.norm(axis) < 1e-6:
        R = np.eye(3)
    else:
        R = Rotation.from_rotvec(axis / np.linalg.norm(axis) * ang).as_matrix()
    moved = points + (cl_pt - centroid)
    return cl_pt + (moved - cl_pt) @ R.T


def test_types_and_argument_errors():
    pts = [PyContourPoint(i, i, 0.0, 0.0, 10.0 - i, False) for i in range(4)]
    cl = PyCenterline.from_contour_points(pts)
    assert len(cl) == 4 and cl.branch_start_indices == [0]
    assert cl.points[0].tangent == (0.0, 0.0, -1.0) and cl.points[3].tangent == (0.0, 0.0, -1.0)  # centerline.rs:14-43
    assert cl.points_as_tuples()[1] == (0.0, 0.0, 9.0)
    p = PyCenterlinePoint(pts[0], (0, 0, 1))
    assert p.branch_id == 0 and p.radius == 0.0 and "CenterlinePoint(point=Point(" in repr(p)
    g = pullback()
    with pytest.raises(TypeError, match="geometry must be a PyGeometry or PyGeometryPair"):  # binding/align.rs:151
        align_three_point(cl, "nope", (0, 0, 0), (0, 0, 0), (0, 0, 0))
    no_ref = pullback(ref_index=None)
    with pytest.raises(nat.MmrsError, match="No reference point found in any frame"):
        align_three_point(cl, no_ref, (0, 0, 0), (0, 0, 0), (0, 0, 0))
    side = PyCenterline([PyCenterlinePoint(q, (0, 0, -1), 1) for q in pts])
    with pytest.raises(nat.MmrsError, match="Centerline has no branch-0 points"):  # preprocessing.rs:14-20
        align_manual(side, g, 0.0, (0, 0, 0))


def test_spacing_is_the_mean_centroid_distance():
    # preprocessing.rs:81-154: centroids (0,0,0), (3,4,0), (6,8,0) -> Some(5.0)
    frames = [ring_frame(f, (3.0 * f, 4.0 * f, 0.0), 1.0, 12, ref_index=0 if f == 0 else None) for f in range(3)]
    g = PyGeometry(frames, "k")
    for f, fr in enumerate(g.frames):
        fr.centroid = (3.0 * f, 4.0 * f, 0.0)
    cl = centerline((0, 0, 50), (0, 0, -1), 40, 1.0)
    _, spacing, rot = align_manual(cl, g, 0.0, (0, 0, 50))
    assert spacing == 5.0 and rot == 0.0


def test_manual_alignment_places_frames_on_the_resampled_centerline():
    g = pullback(n_frames=6, dz=1.25, extras=True)
    cl = centerline((10.0, -3.0, 40.0), (0.3, -0.2, -1.0), 60, 0.4)   # ascending-z input is reversed first
    rev = PyCenterline(list(reversed(cl.points)))
    for use in (cl, rev):
        res, spacing, rot = align_manual(use, g, 25.0, (10.0, -3.0, 40.0))
        cents = np.array([f.centroid for f in g.frames])
        want_spacing = np.mean(np.linalg.norm(np.diff(cents, axis=0), axis=1))
        assert abs(spacing - want_spacing) < 1e-12 and abs(rot - 25.0) < 1e-12
        d = np.array([0.3, -0.2, -1.0]) / np.linalg.norm([0.3, -0.2, -1.0])
        for i, f in enumerate(res.frames):
            want_c = np.array([10.0, -3.0, 40.0]) + i * spacing * d   # resampled point i (preprocessing.rs:133-245)
            assert np.allclose(f.centroid, want_c, atol=1e-9)
            assert np.allclose(f.lumen.centroid, want_c, atol=1e-9)
            assert np.allclose(f.extras["Catheter"].centroid, want_c, atol=1e-9)
            # rigid: point-to-centroid distances are preserved, and the lumen plane is normal to the tangent
            src = g.frames[i].lumen.points_array()[:, 2:5]
            dst = f.lumen.points_array()[:, 2:5]
            assert np.allclose(np.sort(np.linalg.norm(dst - want_c, axis=1)),
                               np.sort(np.linalg.norm(src - np.array(g.frames[i].lumen.centroid), axis=1)), atol=1e-9)
            tangent = d if use is cl else d   # the reversed input is flipped back; tangents keep their stored values
            if use is cl:
                assert np.allclose((dst - want_c) @ tangent, 0.0, atol=1e-9)


def test_manual_alignment_matches_numpy_restatement():
    g = pullback(n_frames=5, dz=1.0)
    cl = centerline((1.0, 2.0, 30.0), (0.5, 0.1, -1.0), 50, 0.5)
    rot_deg = 40.0
    res, spacing, _ = align_manual(cl, g, rot_deg, (1.0, 2.0, 30.0))
    d = np.array([0.5, 0.1, -1.0]) / np.linalg.norm([0.5, 0.1, -1.0])
    a = math.radians(rot_deg)
    for i, f in enumerate(g.frames):
        pts = f.lumen.points_array()[:, 2:5].copy()
        c = np.array(f.centroid)
        # Geometry::rotate_geometry (geometry.rs:241-250): in-plane rotation about the frame centroid, then re-sort
        x, y = pts[:, 0] - c[0], pts[:, 1] - c[1]
        rp = np.stack([x * math.cos(a) - y * math.sin(a) + c[0], x * math.sin(a) + y * math.cos(a) + c[1], pts[:, 2]], 1)
        order = np.argsort(np.arctan2(rp[:, 1] - rp[:, 1].mean(), rp[:, 0] - rp[:, 0].mean()), kind="stable")
        rp = rp[order]
        top = max(range(len(rp)), key=lambda k: (rp[k, 1], k))
        rp = np.roll(rp, -top, axis=0)
        want = place(rp, np.array(f.lumen.centroid), np.array([1.0, 2.0, 30.0]) + i * spacing * d, d)
        got = res.frames[i].lumen.points_array()[:, 2:5]
        assert np.allclose(got, want, atol=1e-9)
        assert list(res.frames[i].lumen.points_array()[:, 1]) == list(range(len(rp)))


@pytest.mark.parametrize("direction", [(0.0, 0.0, -1.0), (0.4, -0.3, -1.0)])
@pytest.mark.parametrize("theta_deg", [0.0, 37.0, 211.0])
def test_three_point_recovers_a_known_rotation(direction, theta_deg):
    """Targets = where the reference frame's ref point / point 0 / point n/2 land after rotating the lumen by
    theta about its normal and placing it on the centerline (align_algorithms.rs:264-337); the 1-degree sweep
    must return exactly that step."""
    n = 48
    g = pullback(n_frames=4, n=n, ref_frame=0, ref_index=7, ry=1.6)
    pair = PyGeometryPair(g, pullback(n_frames=4, n=n, ref_frame=0, ref_index=7, label="b", ry=1.6), "g - b")
    p0 = np.array([3.0, -2.0, 25.0])
    cl = centerline(p0, direction, 40, 0.5)
    d = np.asarray(direction) / np.linalg.norm(direction)
    lum = g.frames[0].lumen
    pts = lum.points_array()[:, 2:5]
    c = np.array(lum.centroid)
    nrm = newell(pts, c)
    R = Rotation.from_rotvec(nrm * math.radians(theta_deg)).as_matrix()
    placed = place(c + (pts - c) @ R.T, c, p0, d)
    main, ccw, cw = placed[7], placed[0], placed[n // 2]
    for target in (g, pair):
        res, spacing, rot = align_three_point(cl, target, tuple(main), tuple(ccw), tuple(cw), angle_step_deg=1.0)
        assert abs(rot - theta_deg) < 1e-9, rot
        assert type(res) is type(target)
        first = res.geom_a if isinstance(res, PyGeometryPair) else res
        assert np.allclose(first.frames[0].centroid, p0, atol=1e-9)
        if isinstance(res, PyGeometryPair):
            assert res.label == "g - b" and np.allclose(res.geom_b.frames[0].centroid, p0, atol=1e-9)
    # the placed geometry equals the manual alignment by the same angle (align.rs:61-164 share the tail)
    man, _, _ = align_manual(cl, g, rot, tuple(p0))
    auto, _, _ = align_three_point(cl, g, tuple(main), tuple(ccw), tuple(cw), angle_step_deg=1.0)
    for fa, fm in zip(auto.frames, man.frames):
        assert np.allclose(fa.lumen.points_array(), fm.lumen.points_array(), atol=1e-9)


def test_wall_alignment_flag_only_touches_walls():
    g = pullback(n_frames=4, n=24, ry=1.5)
    for f in g.frames:
        w = ring_frame(f.id, f.centroid, 2.6, 24, ry=2.0).lumen
        w.kind = "Wall"
        rows = w.points_array()
        rows[:6, 5] = 1.0   # aortic half
        f.extras["Wall"] = w
    # twist the wall of frame 2 so align_walls (align.rs:443-520) has something to undo
    w2 = g.frames[2].extras["Wall"].points_array()
    c = np.array(g.frames[2].centroid)
    a = math.radians(30)
    x, y = w2[:, 2] - c[0], w2[:, 3] - c[1]
    w2[:, 2], w2[:, 3] = x * math.cos(a) - y * math.sin(a) + c[0], x * math.sin(a) + y * math.cos(a) + c[1]
    cl = centerline((0.0, 0.0, 30.0), (0.0, 0.0, -1.0), 40, 0.5)
    plain, _, _ = align_manual(cl, g, 0.0, (0.0, 0.0, 30.0))
    fixed, _, _ = align_manual(cl, g, 0.0, (0.0, 0.0, 30.0), align_wall_anomalous=True)
    for k in range(4):
        assert np.array_equal(plain.frames[k].lumen.points_array(), fixed.frames[k].lumen.points_array())

    def aortic_dir(fr):
        r = fr.extras["Wall"].points_array()
        m = r[r[:, 5] != 0][:, 2:5].mean(0) - np.array(fr.centroid)
        return m / np.linalg.norm(m)

    assert aortic_dir(plain.frames[2]) @ aortic_dir(plain.frames[1]) < 0.9
    for k in (1, 2, 3):  # after the fix every wall's aortic direction is the parallel transport of frame 0's
        assert aortic_dir(fixed.frames[k]) @ aortic_dir(fixed.frames[0]) > 1 - 1e-9


# ---- against the oracle restatement (oracle/centerline_py.py, pinned on the reference's Rust tests) --------------------
def _oracle_frames(g):
    from oracle import oracle_py as ora

    return ora.decode_geometry(g.to_blob())


def _assert_same_geometry(got_frames, want_frames, atol=1e-9):
    assert len(got_frames) == len(want_frames)
    for a, b in zip(got_frames, want_frames):
        assert a["id"] == b["id"] and np.allclose(a["centroid"], b["centroid"], atol=atol, rtol=0)
        assert (a["reference_point"] is None) == (b["reference_point"] is None)
        if a["reference_point"] is not None:
            assert np.allclose(a["reference_point"], b["reference_point"], atol=atol, rtol=0)
        assert sorted(a["contours"]) == sorted(b["contours"])
        for k in a["contours"]:
            ca, cb = a["contours"][k], b["contours"][k]
            assert np.array_equal(ca["points"][:, :2], cb["points"][:, :2])          # frame / point indices: exact
            assert np.allclose(ca["points"][:, 2:5], cb["points"][:, 2:5], atol=atol, rtol=0)
            assert (ca["centroid"] is None) == (cb["centroid"] is None)
            if ca["centroid"] is not None:
                assert np.allclose(ca["centroid"], cb["centroid"], atol=atol, rtol=0)


@pytest.mark.parametrize("pair", [False, True])
def test_manual_and_three_point_match_the_oracle(pair):
    """The product's host-f64 methods against the independent pure-Python restatement: spacing and the selected
    three-point angle EXACTLY, every coordinate within 1e-9 (nalgebra's association order cannot be confirmed from the
    reference checkout, oracle/centerline_py.py header)."""
    from oracle import centerline_py as oc

    g = pullback(n_frames=6, n=48, dz=1.25, extras=True, ref_index=7, ry=1.6)
    g2 = pullback(n_frames=6, n=48, dz=1.25, extras=True, ref_index=7, ry=1.3, label="h")
    target = PyGeometryPair(g, g2, "pair") if pair else g
    cl = centerline((10.0, -3.0, 40.0), (0.3, -0.2, -1.0), 60, 0.4)
    ocl = oc.centerline_from_rows(cl._rows())
    ogeoms = [_oracle_frames(g), _oracle_frames(g2)] if pair else [_oracle_frames(g)]
    ref = (10.0, -3.0, 40.0)
    # align_manual
    res, spacing, rot = align_manual(cl, target, 25.0, ref)
    want, wspacing, wrot = oc.align_manual(ocl, ogeoms, 25.0, ref)
    assert spacing == wspacing and math.radians(rot) == pytest.approx(wrot, abs=1e-15)
    got = [res.geom_a, res.geom_b] if pair else [res]
    for gg, ww in zip(got, want):
        _assert_same_geometry(_oracle_frames(gg), ww)
    # align_three_point: landmarks taken from a planted 40-degree placement
    planted, _, _ = oc.align_manual(ocl, ogeoms[:1], 40.0, ref)
    lum = planted[0][0]["contours"][0]["points"]
    main, ccw, cw = tuple(lum[7][2:5]), tuple(lum[0][2:5]), tuple(lum[24][2:5])
    res, spacing, rot = align_three_point(cl, target, main, ccw, cw, angle_step_deg=1.0)
    want, wspacing, wrot = oc.align_three_point(ocl, ogeoms, main, ccw, cw, math.radians(1.0))
    assert spacing == wspacing
    assert abs(math.radians(rot) - wrot) < 1e-12      # the same candidate of the serial `while angle < TAU` loop
    got = [res.geom_a, res.geom_b] if pair else [res]
    for gg, ww in zip(got, want):
        _assert_same_geometry(_oracle_frames(gg), ww)
