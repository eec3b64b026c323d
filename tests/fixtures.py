"""Python restatement of the reference's Rust test fixtures
(src/intravascular/utils/test_utils.rs:8-418) as geometry-blob frame dicts,
plus the seeded synthetic pullback generator of SURVEY.md §8(d).

Plain Python floats are IEEE f64 and math.sin/cos are glibc's, so the fixture
coordinates are bit-identical to what the Rust constructors produce."""
import math

import numpy as np

HEX = [(1.0, 3.0), (0.0, 2.0), (0.0, 0.0), (1.0, 0.0), (2.0, 0.0), (2.0, 2.0)]


def _pts(frame_index, xy, z):
    return np.array([[frame_index, i, x, y, z, 0.0] for i, (x, y) in enumerate(xy)], dtype=np.float64)


def _centroid(pts):
    sx = sy = sz = 0.0
    for p in pts:
        sx += p[2]
        sy += p[3]
        sz += p[4]
    n = float(len(pts))
    return (sx / n, sy / n, sz / n)


def _translate(pts, dx, dy, dz):
    q = pts.copy()
    for p in q:
        p[2] = p[2] + dx
        p[3] = p[3] + dy
        p[4] = p[4] + dz
    return q


def _rotate(pts, angle, cx, cy):
    if angle == 0.0:
        return pts.copy()
    q = pts.copy()
    c, s = math.cos(angle), math.sin(angle)
    for p in q:
        x, y = p[2] - cx, p[3] - cy
        p[2] = x * c - y * s + cx
        p[3] = x * s + y * c + cy
    return q


def _frame(fid, pts, original_frame, centroid, ref_point=None, lumen_centroid=None):
    return dict(id=fid, centroid=centroid, reference_point=ref_point,
                contours={0: dict(kind=0, id=fid, original_frame=original_frame,
                                  centroid=lumen_centroid if lumen_centroid is not None else centroid,
                                  aortic_thickness=None, pulmonary_thickness=None, points=pts)})


def dummy_geometry():
    """test_utils.rs:111-335."""
    rot = math.radians(15.0) if False else 15.0 * (math.pi / 180.0)  # f64::to_radians
    a = _pts(1, HEX, 0.0)
    b = _translate(_pts(2, HEX, 1.0), 1.0, 1.0, 0.0)
    cb = _centroid(b)
    b = _rotate(b, rot, cb[0], cb[1])
    c = _translate(_pts(3, HEX, 2.0), 2.0, 2.0, 0.0)
    cc = _centroid(c)
    c = _rotate(c, rot * 2.0, cc[0], cc[1])
    ca = _centroid(a)
    ref = np.array([1, 0, 3.0, 1.0, 0.0, 0.0])
    return [_frame(0, a, 1, ca, ref), _frame(1, b, 2, cb), _frame(2, c, 3, cc)]


def frame_translate(f, dx, dy, dz):
    """Frame::translate (frame.rs:18-38) on a frame dict."""
    g = dict(f)
    g["contours"] = {}
    for k, c in f["contours"].items():
        c2 = dict(c)
        c2["points"] = _translate(c["points"], dx, dy, dz)
        c2["centroid"] = _centroid(c2["points"])
        g["contours"][k] = c2
    if f["reference_point"] is not None:
        rp = f["reference_point"].copy()
        rp[2] += dx
        rp[3] += dy
        rp[4] += dz
        g["reference_point"] = rp
    cx, cy, cz = f["centroid"]
    g["centroid"] = (cx + dx, cy + dy, cz + dz)
    return g


def frame_rotate(f, angle, cx, cy):
    """Frame::rotate (frame.rs:40-63)."""
    if angle == 0.0:
        return f
    g = dict(f)
    g["contours"] = {}
    for k, c in f["contours"].items():
        c2 = dict(c)
        c2["points"] = _rotate(c["points"], angle, cx, cy)
        g["contours"][k] = c2
    if f["reference_point"] is not None:
        g["reference_point"] = _rotate(f["reference_point"].reshape(1, 6), angle, cx, cy)[0]
    x, y = f["centroid"][0] - cx, f["centroid"][1] - cy
    c, s = math.cos(angle), math.sin(angle)
    g["centroid"] = (x * c - y * s + cx, x * s + y * c + cy, f["centroid"][2])
    return g


def geometry_rotate_plain(frames, angle):
    """Geometry::rotate_geometry WITHOUT the point re-sort is not what the
    reference does; use oracle-side rotate for that. This helper only rotates."""
    return [frame_rotate(f, angle, f["centroid"][0], f["centroid"][1]) for f in frames]


def dummy_geometry_aligned_long():
    """test_utils.rs:353-383."""
    g1 = dummy_geometry()
    rot = -15.0 * (math.pi / 180.0)
    g1[1] = frame_translate(g1[1], -1.0, -1.0, 0.0)
    g1[2] = frame_translate(g1[2], -2.0, -2.0, 0.0)
    g1[1] = frame_rotate(g1[1], rot, g1[1]["centroid"][0], g1[1]["centroid"][1])
    g1[2] = frame_rotate(g1[2], rot * 2.0, g1[2]["centroid"][0], g1[2]["centroid"][1])
    g2 = []
    for i, f in enumerate(g1):
        idx = i + 3
        f2 = frame_translate(f, 0.0, 0.0, 4.0)
        f2 = _set_value(f2, idx, f2["contours"][0]["centroid"], float(idx))
        g2.append(f2)
    frames = g1 + g2
    frames[3]["reference_point"] = None
    return frames


def _set_value(f, new_id, centroid, z):
    """Frame::set_value(Some(id), None, centroid, Some(z)) (frame.rs:69-118)."""
    g = dict(f)
    g["id"] = new_id
    g["contours"] = {}
    for k, c in f["contours"].items():
        c2 = dict(c)
        c2["id"] = new_id
        c2["centroid"] = (centroid[0], centroid[1], z)
        pts = c["points"].copy()
        pts[:, 4] = z
        c2["points"] = pts
        g["contours"][k] = c2
    if f["reference_point"] is not None:
        rp = f["reference_point"].copy()
        rp[4] = z
        g["reference_point"] = rp
    g["centroid"] = (centroid[0], centroid[1], z)
    return g


def dummy_geometry_center_reference():
    """test_utils.rs:385-418."""
    g1 = dummy_geometry()
    g2 = []
    for i, f in enumerate(dummy_geometry()):
        idx = i + 3
        f2 = frame_translate(f, 0.0, 0.0, 4.0)
        g2.append(_set_value(f2, idx, f2["contours"][0]["centroid"], float(idx)))
    frames = g1 + g2
    mid = len(frames) // 2
    ref = np.array([frames[mid]["contours"][0]["original_frame"], 0, 3.0, 1.0, frames[mid]["centroid"][2], 0.0])
    frames[0]["reference_point"] = None
    for f in frames:
        f["reference_point"] = None
    frames[mid]["reference_point"] = ref
    return frames


# ---- synthetic pullbacks (SURVEY.md §8(d) "Synthetic inputs") -----------------
def synthetic_pullback(n_frames, n_points, seed, z_step=0.5, rot_sigma_deg=4.0):
    """(N,4) [frame, x, y, z] lumen array + reference point [frame, x, y, z].

    r(phi) = r0 (1 + e cos 2(phi-psi) + sum_k a_k cos(k phi + phi_k)); shape
    parameters random-walk along the pullback; cumulative truth rotation
    N(0, rot_sigma^2) per frame; centroid jitter N(0, 0.3^2) about (4.5, 4.5);
    5 um point noise. Frame ids are consecutive; the reference point sits on the
    proximal (highest-id) frame."""
    rng = np.random.default_rng(seed)
    phi = np.linspace(0.0, 2.0 * np.pi, n_points, endpoint=False)
    r0 = rng.uniform(1.5, 3.0)
    e = rng.uniform(0.05, 0.35)
    psi = rng.uniform(0, np.pi)
    ak = rng.uniform(0, 0.04, size=4)
    pk = rng.uniform(0, 2 * np.pi, size=4)
    cum = 0.0
    rows = []
    for f in range(n_frames):
        r0 = float(np.clip(r0 + rng.normal(0, 0.02), 1.2, 3.2))
        e = float(np.clip(e + rng.normal(0, 0.01), 0.03, 0.4))
        ak = np.clip(ak + rng.normal(0, 0.002, size=4), 0, 0.05)
        cum += np.deg2rad(rng.normal(0, rot_sigma_deg))
        r = r0 * (1 + e * np.cos(2 * (phi - psi)) + sum(ak[k] * np.cos((k + 3) * phi + pk[k]) for k in range(4)))
        x = r * np.cos(phi + cum) + 4.5 + rng.normal(0, 0.3) + rng.normal(0, 0.005, size=n_points)
        y = r * np.sin(phi + cum) + 4.5 + rng.normal(0, 0.3) + rng.normal(0, 0.005, size=n_points)
        z = np.full(n_points, z_step * (n_frames - 1 - f))  # highest frame id = proximal = lowest z
        rows.append(np.stack([np.full(n_points, float(f)), x, y, z], axis=1))
    lumen = np.concatenate(rows, axis=0)
    last = rows[-1]
    ref_point = np.array([float(n_frames - 1), last[0, 1] + 0.1, last[0, 2], last[0, 3]])
    return lumen, ref_point
