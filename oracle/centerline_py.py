"""ORACLE (test infrastructure only — nothing under multimoda-rs_b200/ imports this file).

Pure-Python f64 restatement of the reference's centerline alignment, the third caller of the Hausdorff metric
(SURVEY.md §8 row f3). Every function cites the Rust it follows; plain Python floats are IEEE doubles, `math.sin /
cos / sqrt / acos / atan2` are the glibc routines Rust's f64 methods call, and nothing here is fused or vectorised,
so the operation ORDER of the reference is the order below.

    src/intravascular/centerline_align/preprocessing.rs:16-289   preprocess_centerline (+ helpers)
    src/intravascular/centerline_align/align_algorithms.rs:66-520 FrameTransformation, align_frame, calculate_normal,
                                                                  best_rotation_three_point, refine_alignment_hausdorff,
                                                                  filter_points_in_region, apply_transformations
    src/intravascular/centerline_align/align.rs:63-283            align_three_point_rs / align_manual_rs / align_combined_rs
                                                                  (without align_walls and the OBJ `write` step)
    src/types/native/centerline.rs:14-62                          Centerline::from_contour_points, find_reference_cl_point_idx
    src/types/native/geometry.rs:62-69, :241-250                  find_ref_frame_idx, rotate_geometry
    src/types/native/frame.rs:40-63, :123-129                     Frame::rotate, sort_frame_points
    src/types/native/contour.rs:47-58, :368-405                   downsample_contour_points, sort_contour_points
    src/intravascular/processing/process_utils.rs:78-121          hausdorff_distance (x, y only)

Third-party arithmetic: the reference uses nalgebra 0.35.0 (Cargo.lock; NOT vendored under /root/reference) for
Vector3 / Point3 / Rotation3. Its published algorithms are restated here: `Rotation3::from_axis_angle` is the
Rodrigues matrix of nalgebra's `Rotation3::from_scaled_axis`/`from_axis_angle` (src/geometry/rotation_specialization.rs),
`Vector::angle` is acos(clamp(dot / (|a| |b|), -1, 1)) with 0 for a zero vector (src/base/norm.rs / matrix.rs),
`norm` is sqrt(x^2 + y^2 + z^2), matrix x vector is the column-by-column accumulation of nalgebra's gemv. Because the
exact association order inside nalgebra cannot be confirmed from this checkout, floating-point outputs of this file are
compared at 1e-9 absolute (tests/test_oracle_centerline.py, tests/test_centerline_gpu.py); the DISCRETE results (best
three-point angle, refined angle and centerline index) are compared exactly.

Parity pins: the reference's own unit tests of this path — preprocessing.rs:291-605 and align_algorithms.rs:573-935 —
are restated as known-answer tests in tests/test_oracle_centerline.py.

Data model (plain Python, small cases only): a geometry is the list of frame dicts `oracle_py.decode_geometry`
returns (contours keyed by kind, points as (n, 6) rows [frame_index, point_index, x, y, z, aortic]); a centerline is
a list of dicts {p: [x, y, z], t: [tx, ty, tz], branch: int, radius: float}."""
from __future__ import annotations

import copy
import math

import numpy as np

TAU = 6.283185307179586  # core::f64::consts::TAU


# ---- nalgebra 0.35 pieces -------------------------------------------------------------------------------------------
def v_norm(v):
    return math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])


def v_dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def v_cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def v_angle(a, b):
    """nalgebra Matrix::angle: 0 if either vector is zero, else acos of the clamped cosine."""
    prod = v_dot(a, b)
    n1, n2 = v_norm(a), v_norm(b)
    if n1 == 0.0 or n2 == 0.0:
        return 0.0
    c = prod / (n1 * n2)
    return math.acos(-1.0 if c < -1.0 else 1.0 if c > 1.0 else c)


def unit(v):
    n = v_norm(v)
    return [v[0] / n, v[1] / n, v[2] / n]


def rot_from_axis_angle(u, angle):
    """nalgebra Rotation3::from_axis_angle (axis already a unit vector): identity for a zero angle, else Rodrigues."""
    if angle == 0.0:
        return [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
    ux, uy, uz = u
    sqx, sqy, sqz = ux * ux, uy * uy, uz * uz
    s, c = math.sin(angle), math.cos(angle)
    omc = 1.0 - c
    return [[sqx + (1.0 - sqx) * c, ux * uy * omc - uz * s, ux * uz * omc + uy * s],
            [ux * uy * omc + uz * s, sqy + (1.0 - sqy) * c, uy * uz * omc - ux * s],
            [ux * uz * omc - uy * s, uy * uz * omc + ux * s, sqz + (1.0 - sqz) * c]]


IDENTITY = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]


def mat_vec(m, v):
    # gemv: y = col0 * v0, then y += col_j * v_j
    return [m[0][0] * v[0] + m[0][1] * v[1] + m[0][2] * v[2],
            m[1][0] * v[0] + m[1][1] * v[1] + m[1][2] * v[2],
            m[2][0] * v[0] + m[2][1] * v[1] + m[2][2] * v[2]]


# ---- centerline ---------------------------------------------------------------------------------------------------------
def centerline_from_rows(rows):
    """rows: (n, 8) [x, y, z, tx, ty, tz, branch_id, radius] (the layout mmrs_align_centerline takes)."""
    return [dict(p=[float(r[0]), float(r[1]), float(r[2])], t=[float(r[3]), float(r[4]), float(r[5])],
                 branch=int(r[6]), radius=float(r[7])) for r in np.asarray(rows, dtype=np.float64).reshape(-1, 8)]


def centerline_rows(cl):
    return np.array([[*q["p"], *q["t"], float(q["branch"]), q["radius"]] for q in cl], dtype=np.float64).reshape(-1, 8)


def find_reference_cl_point_idx(cl, ref):  # centerline.rs:51-62 (ContourPoint::distance_to: sqrt of the squared sum)
    best_idx, best = 0, math.inf
    for i, q in enumerate(cl):
        dx, dy, dz = q["p"][0] - ref[0], q["p"][1] - ref[1], q["p"][2] - ref[2]
        d = math.sqrt(dx * dx + dy * dy + dz * dz)
        if d < best:
            best, best_idx = d, i
    return best_idx


# ---- preprocessing.rs ---------------------------------------------------------------------------------------------------
def calculate_mean_spacing(frames):  # :245-289
    c = [f["centroid"] for f in frames]
    d = []
    for i in range(1, len(c)):
        dx, dy, dz = c[i][0] - c[i - 1][0], c[i][1] - c[i - 1][1], c[i][2] - c[i - 1][2]
        d.append(math.sqrt(dx * dx + dy * dy + dz * dz))
    if not d:
        return None
    s = 0.0
    for x in d:
        s += x
    mean = s / float(len(d))
    return mean if math.isfinite(mean) and mean > 1e-12 else None


def cumulative_arc_length(cl):  # :105-121
    cum = []
    if not cl:
        return cum
    cum.append(0.0)
    for i in range(1, len(cl)):
        p0, p1 = cl[i - 1]["p"], cl[i]["p"]
        dx, dy, dz = p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]
        cum.append(cum[-1] + math.sqrt(dx * dx + dy * dy + dz * dz))
    return cum


def decide_spacing(mean_opt, total_length, n_segments):  # :123-138
    if mean_opt is not None and math.isfinite(mean_opt) and mean_opt > 1e-12:
        return mean_opt
    if n_segments >= 1:
        fb = total_length / float(n_segments)
        if math.isfinite(fb) and fb > 1e-12:
            return fb
    return None


def build_samples(total_length, spacing):  # :140-155
    s_new, s = [], 0.0
    while s <= total_length + 1e-9:
        s_new.append(s)
        s += spacing
    if s_new and s_new[-1] > total_length + 1e-6:
        s_new.pop()
        s_new.append(total_length)
    return s_new


def interpolate_centerline_at_s(cl, cum, target_s):  # :157-243
    # binary_search_by(partial_cmp): exact match -> that index (any of equal ones; cum is increasing), else insertion - 1
    lo, hi, idx = 0, len(cum), None
    while lo < hi:
        mid = (lo + hi) // 2
        if cum[mid] == target_s:
            idx = mid
            break
        if cum[mid] < target_s:
            lo = mid + 1
        else:
            hi = mid
    if idx is None:
        idx = 0 if lo == 0 else lo - 1
    if idx >= max(len(cl) - 1, 0):
        q = cl[-1]
        return dict(p=list(q["p"]), t=list(q["t"]), branch=0, radius=q["radius"])
    p0, p1 = cl[idx]["p"], cl[idx + 1]["p"]
    s0, s1 = cum[idx], cum[idx + 1]
    denom = s1 - s0
    t = 0.0 if abs(denom) < 1e-12 else (target_s - s0) / denom
    p = [p0[0] + t * (p1[0] - p0[0]), p0[1] + t * (p1[1] - p0[1]), p0[2] + t * (p1[2] - p0[2])]
    t0, t1 = cl[idx]["t"], cl[idx + 1]["t"]
    tan = [0.0, 0.0, 0.0]
    if v_norm(t0) > 0.0 or v_norm(t1) > 0.0:
        tan = [t0[k] * (1.0 - t) + t1[k] * t for k in range(3)]
        n = v_norm(tan)
        tan = [tan[0] / n, tan[1] / n, tan[2] / n] if n > 1e-12 else [0.0, 0.0, 0.0]
    radius = cl[idx]["radius"] * (1.0 - t) + cl[idx + 1]["radius"] * t
    return dict(p=p, t=tan, branch=0, radius=radius)


def preprocess_centerline(cl, frames):  # :16-103
    pts = [q for q in cl if q["branch"] == 0]
    if not pts:
        raise ValueError("Centerline has no branch-0 points")
    pts = copy.deepcopy(pts)
    if pts[0]["p"][2] < pts[-1]["p"][2]:  # ensure_descending_z, :38-46
        pts.reverse()
    if not frames:
        raise ValueError("Reference mesh has no frames")
    mean_opt = calculate_mean_spacing(frames)
    cum = cumulative_arc_length(pts)
    total = cum[-1] if cum else 0.0
    spacing = decide_spacing(mean_opt, total, max(len(pts) - 1, 0))
    if spacing is None:
        return pts, 0.0
    return [interpolate_centerline_at_s(pts, cum, s) for s in build_samples(total, spacing)], spacing


# ---- geometry pieces ----------------------------------------------------------------------------------------------------
LUMEN = 0


def find_ref_frame_idx(frames):  # geometry.rs:62-69 (returns frame.id as the index)
    for f in frames:
        if f["reference_point"] is not None:
            return int(f["id"])
    raise ValueError("No reference point found in any frame")


def _rotate_xy(x, y, angle, cx, cy):  # contour_point.rs:38-52
    if angle == 0.0:
        return x, y
    dx, dy = x - cx, y - cy
    c, s = math.cos(angle), math.sin(angle)
    return dx * c - dy * s + cx, dx * s + dy * c + cy


def sort_contour_points(pts):  # contour.rs:368-405 (stable sort by atan2 about the mean, highest y first, re-index)
    n = len(pts)
    if n == 0:
        return pts
    sx = sy = 0.0
    for r in pts:
        sx += r[2]
        sy += r[3]
    cx, cy = sx / float(n), sy / float(n)
    order = sorted(range(n), key=lambda i: math.atan2(pts[i][3] - cy, pts[i][2] - cx))   # Python's sort is stable, like sort_by
    pts = pts[order]
    start, best = 0, None
    for i in range(n):           # Iterator::max_by keeps the LAST maximum
        if best is None or pts[i][3] >= best:
            best, start = pts[i][3], i
    pts = np.concatenate([pts[start:], pts[:start]])
    pts[:, 1] = np.arange(n, dtype=np.float64)
    return pts


def rotate_geometry(frames, angle):  # geometry.rs:241-250 (Frame::rotate about the frame centroid, then re-sort)
    if angle == 0.0:
        return frames
    for f in frames:
        cx, cy = f["centroid"][0], f["centroid"][1]
        for c in f["contours"].values():
            p = c["points"]
            for i in range(len(p)):
                p[i][2], p[i][3] = _rotate_xy(p[i][2], p[i][3], angle, cx, cy)
        if f["reference_point"] is not None:
            rp = f["reference_point"]
            rp[2], rp[3] = _rotate_xy(rp[2], rp[3], angle, cx, cy)
        x, y = f["centroid"][0] - cx, f["centroid"][1] - cy            # frame.rs:55-61
        ca, sa = math.cos(angle), math.sin(angle)
        f["centroid"] = (x * ca - y * sa + cx, x * sa + y * ca + cy, f["centroid"][2])
        for c in f["contours"].values():
            c["points"] = sort_contour_points(c["points"])
    return frames


# ---- align_algorithms.rs -------------------------------------------------------------------------------------------------
def calculate_normal(pts, centroid):  # :186-220 (Newell)
    if len(pts) < 3:
        return [0.0, 0.0, 1.0]
    nx = ny = nz = 0.0
    n = len(pts)
    for i in range(n):
        cur, nxt = pts[i], pts[(i + 1) % n]
        nx += (cur[3] - centroid[1]) * (nxt[4] - centroid[2]) - (cur[4] - centroid[2]) * (nxt[3] - centroid[1])
        ny += (cur[4] - centroid[2]) * (nxt[2] - centroid[0]) - (cur[2] - centroid[0]) * (nxt[4] - centroid[2])
        nz += (cur[2] - centroid[0]) * (nxt[3] - centroid[1]) - (cur[3] - centroid[1]) * (nxt[2] - centroid[0])
    norm = v_norm([nx, ny, nz])
    if norm > 1e-12:
        return [nx / norm, ny / norm, nz / norm]
    return [0.0, 0.0, 1.0]


def _contour_centroid(c):
    if c["centroid"] is not None:
        return c["centroid"]
    p, n = c["points"], float(len(c["points"]))
    return (sum(r[2] for r in p) / n, sum(r[3] for r in p) / n, sum(r[4] for r in p) / n)


def align_frame(contour, clp):  # :128-173
    cen = _contour_centroid(contour)
    tr = [clp["p"][0] - cen[0], clp["p"][1] - cen[1], clp["p"][2] - cen[2]]
    cur = calculate_normal(contour["points"], cen)
    des = clp["t"]
    ang = v_angle(cur, des)
    if abs(ang) < 1e-6:
        rot = IDENTITY
    else:
        axis = v_cross(cur, des)
        rot = IDENTITY if v_norm(axis) < 1e-6 else rot_from_axis_angle(unit(axis), ang)
    return dict(translation=tr, rotation=rot, pivot=list(clp["p"]))


def apply_to_xyz(tr, x, y, z):  # FrameTransformation::apply_to_point, :73-94
    t = [x + tr["translation"][0], y + tr["translation"][1], z + tr["translation"][2]]
    rel = [t[0] - tr["pivot"][0], t[1] - tr["pivot"][1], t[2] - tr["pivot"][2]]
    r = mat_vec(tr["rotation"], rel)
    return tr["pivot"][0] + r[0], tr["pivot"][1] + r[1], tr["pivot"][2] + r[2]


def apply_transformation_to_contour(c, tr):  # :176-201
    p = c["points"]
    for i in range(len(p)):
        p[i][2], p[i][3], p[i][4] = apply_to_xyz(tr, p[i][2], p[i][3], p[i][4])
    if c["centroid"] is not None:
        c["centroid"] = apply_to_xyz(tr, *c["centroid"])


def rotate_contour_around_centroid(c, angle):  # :238-259
    cen = _contour_centroid(c)
    axis = calculate_normal(c["points"], cen)
    rot = rot_from_axis_angle(unit(axis), angle)
    p = c["points"]
    for i in range(len(p)):
        rel = [p[i][2] - cen[0], p[i][3] - cen[1], p[i][4] - cen[2]]
        r = mat_vec(rot, rel)
        p[i][2], p[i][3], p[i][4] = cen[0] + r[0], cen[1] + r[1], cen[2] + r[2]


def best_rotation_three_point(contour, ref_point, main, ccw, cw, angle_step, clp):  # :263-336
    idx_ref = int(ref_point[1])
    best_angle, min_err, angle = 0.0, 1.7976931348623157e308, 0.0
    while angle < TAU:
        tmp = copy.deepcopy(contour)
        rotate_contour_around_centroid(tmp, angle)
        apply_transformation_to_contour(tmp, align_frame(tmp, clp))
        pts = tmp["points"]
        n = len(pts)
        pick = lambda k: next(r for r in pts if int(r[1]) == k)   # noqa: E731  (Iterator::find)
        pm, pc, pw = pick(idx_ref), pick(0), pick(n // 2)

        def dist(r, t):
            dx, dy, dz = r[2] - t[0], r[3] - t[1], r[4] - t[2]
            return math.sqrt(dx * dx + dy * dy + dz * dz)

        dm, dc, dw = dist(pm, main), dist(pc, ccw), dist(pw, cw)
        err = dm * dm + dc * dc + dw * dw          # powi(2)
        if err < min_err:
            min_err, best_angle = err, angle
        angle += angle_step
    return best_angle


def get_transformations(frames, cl, ref_pt):  # :97-126
    ref_idx = find_reference_cl_point_idx(cl, ref_pt)
    out = []
    for i, f in enumerate(frames):
        k = ref_idx + i
        if 0 <= k < len(cl):
            out.append(align_frame(f["contours"][LUMEN], cl[k]))
    return out


def apply_transforms_to_geometry(frames, trs):  # :503-519
    for i, f in enumerate(frames):
        if i < len(trs):
            tr = trs[i]
            for c in f["contours"].values():
                apply_transformation_to_contour(c, tr)
            if f["reference_point"] is not None:
                rp = f["reference_point"]
                rp[2], rp[3], rp[4] = apply_to_xyz(tr, rp[2], rp[3], rp[4])
            lc = f["contours"][LUMEN]["centroid"]
            f["centroid"] = tuple(lc) if lc is not None else (0.0, 0.0, 0.0)


def apply_transformations(geoms, cl, ref_pt):  # :492-501 (the transformations come from the PRIMARY geometry)
    trs = get_transformations(copy.deepcopy(geoms[0]), cl, ref_pt)
    for g in geoms:
        apply_transforms_to_geometry(g, trs)
    return geoms


def rotate_by_best_rotation(geoms, angle):  # :488-490
    for g in geoms:
        rotate_geometry(g, angle)
    return geoms


def filter_points_in_region(points, a, b, margin=5.0):  # :454-486
    lo = [min(a["p"][k], b["p"][k]) - margin for k in range(3)]
    hi = [max(a["p"][k], b["p"][k]) + margin for k in range(3)]
    return [p for p in points if all(lo[k] <= p[k] <= hi[k] for k in range(3))]


def downsample(pts, n):  # contour.rs:47-58
    if len(pts) <= n:
        return [r for r in pts]
    step = float(len(pts)) / float(n)
    return [pts[int(float(i) * step)] for i in range(n)]


def hausdorff_xy(a, b):  # process_utils.rs:78-121 (x, y only; non-finite minima skipped)
    def directed(p, q):
        if not p or not q:
            return 0.0
        mx = 0.0
        for u in p:
            mn = math.inf
            for v in q:
                dx, dy = u[0] - v[0], u[1] - v[1]
                d2 = dx * dx + dy * dy
                if d2 < mn:
                    mn = d2
            if math.isfinite(mn) and mn > mx:
                mx = mn
        return math.sqrt(mx)

    return max(directed(a, b), directed(b, a))


def refine_alignment_hausdorff(geoms, cl, initial_idx, initial_rotation, points, angle_range, angle_step, index_range):
    """:339-451. geoms: the already aligned target (list of 1 or 2 geometries, the first is primary);
    points: [(x, y, z)]. Returns (best_angle, best_cl_ref_idx, min_hausdorff, candidates scored)."""
    len_frames = len(geoms[0])
    best_angle, best_idx, min_h, scored = initial_rotation, initial_idx, 1.7976931348623157e308, 0
    deltas = [0] if index_range == 0 else range(-index_range, index_range + 1)
    for d in deltas:
        signed = initial_idx + d
        if signed < 0:
            continue
        cur = signed
        if cur + len_frames >= len(cl):
            continue
        end = cur + len_frames
        seg = copy.deepcopy(cl[cur:end])
        angle = initial_rotation - angle_range
        while angle <= initial_rotation + angle_range:
            ref_pt = tuple(cl[cur]["p"])
            tr = apply_transformations(rotate_by_best_rotation(copy.deepcopy(geoms), angle), seg, ref_pt)
            filt = filter_points_in_region(points, cl[cur], cl[end - 1])
            if not filt:
                angle += angle_step
                continue
            frames = tr[0]
            npf = len(frames[0]["contours"][LUMEN]["points"])
            ratio = float(len(filt)) / (float(npf) * float(len_frames))
            nd = int(math.ceil(ratio * float(npf)))
            nd = min(max(nd, 1), npf)
            flat = []
            for f in frames:
                lp = f["contours"][LUMEN]["points"]
                flat.extend((r[2], r[3]) for r in (downsample(lp, nd) if nd < npf else lp))
            h = hausdorff_xy([(p[0], p[1]) for p in filt], flat)
            scored += 1
            if h < min_h:
                min_h, best_angle, best_idx = h, angle, cur
            angle += angle_step
    return best_angle, best_idx, min_h, scored


# ---- align.rs drivers ---------------------------------------------------------------------------------------------------
def _targets(geoms):
    return [copy.deepcopy(g) for g in geoms]


def align_manual(cl, geoms, rotation_deg, ref_pt):  # align.rs:123-164
    geoms = _targets(geoms)
    rcl, spacing = preprocess_centerline(cl, geoms[0])
    total = rotation_deg * (math.pi / 180.0)   # f64::to_radians
    geoms = apply_transformations(rotate_by_best_rotation(geoms, total), rcl, ref_pt)
    return geoms, spacing, total


def _three_point_start(cl, geoms, main, ccw, cw, angle_step):
    rcl, spacing = preprocess_centerline(cl, geoms[0])
    ref_idx = find_ref_frame_idx(geoms[0])
    ref_frame = geoms[0][ref_idx]
    cl_ref_idx = find_reference_cl_point_idx(rcl, main)
    rot = best_rotation_three_point(ref_frame["contours"][LUMEN], ref_frame["reference_point"], main, ccw, cw, angle_step,
                                    rcl[cl_ref_idx])
    return rcl, spacing, cl_ref_idx, rot


def align_three_point(cl, geoms, main, ccw, cw, angle_step):  # align.rs:63-121
    geoms = _targets(geoms)
    rcl, spacing, _, rot = _three_point_start(cl, geoms, main, ccw, cw, angle_step)
    geoms = apply_transformations(rotate_by_best_rotation(geoms, rot), rcl, main)
    return geoms, spacing, rot


def align_combined(cl, geoms, main, ccw, cw, points, angle_step, refine_angle_range, refine_index_range):  # align.rs:166-283
    original = _targets(geoms)
    rcl, spacing, idx0, rot0 = _three_point_start(cl, original, main, ccw, cw, angle_step)
    aligned = apply_transformations(rotate_by_best_rotation(_targets(original), rot0), rcl, main)
    delta, idx, min_h, scored = refine_alignment_hausdorff(aligned, rcl, idx0, 0.0, [tuple(p) for p in points],
                                                           refine_angle_range, angle_step, refine_index_range)
    total = rot0 + delta
    refined_ref = tuple(rcl[idx]["p"])
    final = apply_transformations(rotate_by_best_rotation(_targets(geoms), total), rcl, refined_ref)
    return final, spacing, total, dict(initial_rotation=rot0, initial_idx=idx0, refined_idx=idx, delta=delta,
                                       min_hausdorff=min_h, scored=scored)
