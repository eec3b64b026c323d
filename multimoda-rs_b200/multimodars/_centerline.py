"""Branch bookkeeping of `PyCenterline` — the methods a user runs on a raw centerline before handing it to
`align_three_point` / `align_manual` / `align_combined` (src/types/binding/py_centerline.rs:62-330 over
src/types/native/centerline.rs:78-938). Host-side f64 list/array work; every function returns a NEW centerline and
leaves its input untouched, like the reference's `to_rust_centerline()` round trip.

Distances are `sqrt(dx*dx + dy*dy + dz*dz)` in that order (types/native.rs:27-32), tangents are nalgebra's
`v / v.norm()` (a zero step gives NaN components, as there)."""
from __future__ import annotations

import math
from collections import deque

import numpy as np

MIN_BRANCH_SIZE = 5  # centerline.rs:79


def _types():
    from . import _types as t
    return t


def _clone_point(p, branch_id=None, point_index=None, tangent=None):
    t = _types()
    c = p.contour_point
    q = t.PyCenterlinePoint(t.PyContourPoint(c.frame_index, c.point_index if point_index is None else point_index,
                                             c.x, c.y, c.z, c.aortic),
                            p.tangent if tangent is None else tangent,
                            p.branch_id if branch_id is None else branch_id)
    q.radius = p.radius
    return q


def _xyz(points) -> np.ndarray:
    return np.array([(p.contour_point.x, p.contour_point.y, p.contour_point.z) for p in points],
                    dtype=np.float64).reshape(-1, 3)


def _dist(a, b) -> float:
    dx, dy, dz = a[0] - b[0], a[1] - b[1], a[2] - b[2]
    return math.sqrt(dx * dx + dy * dy + dz * dz)


def _consecutive(xyz: np.ndarray) -> np.ndarray:
    d = xyz[:-1] - xyz[1:]
    return np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])


def _make(points, starts):
    cl = _types().PyCenterline(points)
    cl.branch_start_indices = list(starts)
    return cl


def _recompute_tangents(points):
    """centerline.rs:377-391: forward difference inside a branch, a branch's last point repeats its predecessor."""
    n = len(points)
    for i, p in enumerate(points):
        if i + 1 < n and p.branch_id == points[i + 1].branch_id:
            a, b = p.contour_point, points[i + 1].contour_point
            dx, dy, dz = b.x - a.x, b.y - a.y, b.z - a.z
            nrm = math.sqrt(dx * dx + dy * dy + dz * dz)
            p.tangent = (dx / nrm, dy / nrm, dz / nrm) if nrm != 0.0 else (math.nan, math.nan, math.nan)
        elif i > 0 and points[i - 1].branch_id == p.branch_id:
            p.tangent = points[i - 1].tangent
        else:
            p.tangent = (0.0, 0.0, 0.0)


def _branches(cl):
    """centerline.rs:394-407."""
    s = list(cl.branch_start_indices)
    ends = s[1:] + [len(cl.points)]
    return [[_clone_point(p) for p in cl.points[a:b]] for a, b in zip(s, ends)]


def _rebuild(branches):
    """centerline.rs:411-430: flat list, sequential branch ids and point indices, fresh tangents."""
    pts, starts, g = [], [], 0
    for bid, br in enumerate(branches):
        starts.append(len(pts))
        for p in br:
            p.branch_id = bid
            p.contour_point.point_index = g
            g += 1
            pts.append(p)
    _recompute_tangents(pts)
    return _make(pts, starts)


def _branch_range(cl, idx):
    s = cl.branch_start_indices
    return s[idx], (s[idx + 1] if idx + 1 < len(s) else len(cl.points))


def mean_spacing(cl) -> float:
    """centerline.rs:304-320: mean consecutive spacing of branch 0 (1.0 with fewer than two points)."""
    s = cl.branch_start_indices
    end = s[1] if len(s) > 1 else len(cl.points)
    if end < 2:
        return 1.0
    d = _consecutive(_xyz(cl.points[:end]))
    total = 0.0
    for v in d.tolist():  # sequential f64 sum like Iterator::sum
        total += v
    return total / (end - 1)


def calculate_branches(cl, spacing_tolerance=1.0):
    """centerline.rs:78-155: segments = runs of consecutive points closer than p95 spacing x tolerance; a sparse tree
    joins consecutive points and the closest pair of every two segments; the tree diameter (double BFS by arc length)
    is branch 0, the other components with >= 5 points follow by descending size, smaller ones are dropped."""
    pts = cl.points
    n = len(pts)
    if n == 0:
        return _make([], [])
    xyz = _xyz(pts)
    gaps = _consecutive(xyz) if n > 1 else np.zeros(0)
    if n < 2:
        p95 = 1.0
    else:
        p95 = float(np.sort(gaps, kind="stable")[(len(gaps) * 95) // 100])
    thr = p95 * spacing_tolerance

    seg = [0] + [i for i in range(1, n) if gaps[i - 1] > thr] + [n]
    adj = [[] for _ in range(n)]
    for i in range(1, n):
        if gaps[i - 1] <= thr:
            adj[i - 1].append(i)
            adj[i].append(i - 1)
    for si in range(len(seg) - 1):
        a = xyz[seg[si]:seg[si + 1]]
        for sj in range(si + 1, len(seg) - 1):
            b = xyz[seg[sj]:seg[sj + 1]]
            d = a[:, None, :] - b[None, :, :]
            dd = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2])
            k = int(np.argmin(dd))  # first minimum in (pi, pj) row-major order == the strict `<` scan
            pi, pj = divmod(k, dd.shape[1])
            if dd[pi, pj] <= thr:
                gi, gj = seg[si] + pi, seg[sj] + pj
                adj[gi].append(gj)
                adj[gj].append(gi)

    def bfs_farthest(start):
        dist = [math.inf] * n
        prev = [None] * n
        dist[start] = 0.0
        q = deque([start])
        far = start
        while q:
            u = q.popleft()
            for v in adj[u]:
                if math.isinf(dist[v]):
                    dist[v] = dist[u] + _dist(xyz[u], xyz[v])
                    prev[v] = u
                    q.append(v)
                    if dist[v] > dist[far]:
                        far = v
        return far, prev

    a, _ = bfs_farthest(0)
    b, prev = bfs_farthest(a)
    main, cur = [], b
    while True:
        main.append(cur)
        if cur == a or prev[cur] is None:
            break
        cur = prev[cur]

    visited = [False] * n
    for i in main:
        visited[i] = True
    comps = []
    for s in range(n):
        if visited[s]:
            continue
        comp, q = [], deque([s])
        visited[s] = True
        while q:
            u = q.popleft()
            comp.append(u)
            for v in adj[u]:
                if not visited[v]:
                    visited[v] = True
                    q.append(v)
        comps.append(comp)
    real = sorted((c for c in comps if len(c) >= MIN_BRANCH_SIZE), key=lambda c: -len(c))  # stable, like sort_by_key

    def order_chain(comp):  # centerline.rs:342-371
        inside = set(comp)
        start = next((i for i in comp if sum(1 for nb in adj[i] if nb in inside) <= 1), comp[0])
        out, seen, c = [], set(), start
        while True:
            out.append(c)
            seen.add(c)
            nxt = next((nb for nb in adj[c] if nb in inside and nb not in seen), None)
            if nxt is None:
                break
            c = nxt
        out.extend(i for i in comp if i not in seen)
        return out

    new, starts, g = [], [0], 0
    for i in main:
        new.append(_clone_point(pts[i], branch_id=0, point_index=g))
        g += 1
    for k, comp in enumerate(real):
        starts.append(len(new))
        for i in order_chain(comp):
            new.append(_clone_point(pts[i], branch_id=k + 1, point_index=g))
            g += 1
    _recompute_tangents(new)
    return _make(new, starts)


def find_sharp_angles(cl, branch_id, cos_threshold):
    """centerline.rs:436-466: global indices of the interior points of a branch whose opening angle has
    cos > `cos_threshold`."""
    idx = int(branch_id)
    if idx < 0:
        raise OverflowError("can't convert negative int to unsigned")
    if idx >= len(cl.branch_start_indices):
        return []
    a, b = _branch_range(cl, idx)
    xyz = _xyz(cl.points[a:b])
    out = []
    for i in range(1, max(len(xyz) - 1, 0)):
        v1, v2 = xyz[i - 1] - xyz[i], xyz[i + 1] - xyz[i]
        n1 = math.sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2])
        n2 = math.sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2])
        if n1 < 1e-10 or n2 < 1e-10:
            continue
        if (v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2]) / (n1 * n2) > cos_threshold:
            out.append(a + i)
    return out


def _sorted_by_length(branches):
    return sorted(branches, key=lambda b: -len(b))  # stable: ties keep their order (centerline.rs:558-560)


def split_branch(cl, branch_id, point_index):
    """centerline.rs:474-505: both halves keep the split point; branches re-sorted by descending length."""
    idx, point_index = int(branch_id), int(point_index)
    if idx >= len(cl.branch_start_indices):
        return _copy(cl)
    a, b = _branch_range(cl, idx)
    if point_index < a or point_index >= b:
        return _copy(cl)
    local = point_index - a
    branches = _branches(cl)
    br = branches.pop(idx)
    if local == 0 or local >= max(len(br) - 1, 0):
        return _copy(cl)
    branches.append(br[:local + 1])
    branches.append([_clone_point(p) for p in br[local:]])
    return _rebuild(_sorted_by_length(branches))


def merge_branches(cl, branch_id_a, branch_id_b):
    """centerline.rs:512-553: concatenated at the closest pair of end points."""
    ia, ib = int(branch_id_a), int(branch_id_b)
    branches = _branches(cl)
    if ia == ib or ia >= len(branches) or ib >= len(branches):
        return _copy(cl)
    low, high = (ia, ib) if ia < ib else (ib, ia)
    bh = branches.pop(high)
    bl = branches.pop(low)
    c = lambda p: (p.contour_point.x, p.contour_point.y, p.contour_point.z)  # noqa: E731
    lf, ll, hf, hl = c(bl[0]), c(bl[-1]), c(bh[0]), c(bh[-1])
    d_ll_hf, d_ll_hl, d_lf_hf, d_lf_hl = _dist(ll, hf), _dist(ll, hl), _dist(lf, hf), _dist(lf, hl)
    m = min(d_ll_hf, d_ll_hl, d_lf_hf, d_lf_hl)
    if abs(m - d_ll_hf) < 1e-12:
        merged = bl + bh
    elif abs(m - d_ll_hl) < 1e-12:
        merged = bl + bh[::-1]
    elif abs(m - d_lf_hf) < 1e-12:
        merged = bh[::-1] + bl
    else:
        merged = bh + bl
    branches.append(merged)
    return _rebuild(_sorted_by_length(branches))


def get_branch(cl, branch_id):
    """py_centerline.rs:216-236."""
    pts = [_clone_point(p, branch_id=0) for p in cl.points if p.branch_id == int(branch_id)]
    if not pts:
        raise ValueError(f"branch_id {branch_id} not found in centerline")
    return _make(pts, [0])


def _should_reverse_relative_to(points, reference):
    """centerline.rs:648-670: is the LAST point nearer to the reference polyline than the first?"""
    if not points or not reference:
        return False
    ref = _xyz(reference)

    def nearest(p):
        d = ref - np.array([p.contour_point.x, p.contour_point.y, p.contour_point.z])
        return float(np.min(np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])))

    return nearest(points[-1]) < nearest(points[0])


def orient_by_max_z(cl):
    """centerline.rs:570-586: branch 0 starts at its highest z; side branches start at the end nearer branch 0."""
    if not cl.branch_start_indices:
        return _copy(cl)
    br = _branches(cl)
    if br[0]:
        zs = [p.contour_point.z for p in br[0]]
        # Iterator::max_by keeps the LAST of equal maxima
        best = 0
        for i, z in enumerate(zs):
            if not (z < zs[best]):
                best = i
        if best != 0:
            br[0].reverse()
    for b in br[1:]:
        if _should_reverse_relative_to(b, br[0]):
            b.reverse()
    return _rebuild(br)


def orient_to_reference(cl, reference):
    """centerline.rs:599-615: every branch starts at the end nearer to the reference's branch 0."""
    if not cl.branch_start_indices:
        return _copy(cl)
    s = reference.branch_start_indices
    ref0 = reference.points[:(s[1] if len(s) > 1 else len(reference.points))]
    br = _branches(cl)
    for b in br:
        if _should_reverse_relative_to(b, ref0):
            b.reverse()
    return _rebuild(br)


def remove_branch_overlap(cl):
    """centerline.rs:681-693, 877-913: drop the prefix a side branch shares (within one mean spacing) with the
    branches before it, keeping the last shared point as the junction; fully shared branches disappear."""
    if not cl.branch_start_indices:
        return _copy(cl)
    buf = mean_spacing(cl)
    buf_sq = buf * buf
    br = _branches(cl)
    if len(br) > 1:
        known = _xyz(br[0])
        for b in br[1:]:
            xyz = _xyz(b)
            first_outside = None
            for j in range(len(b)):
                d = known - xyz[j]
                if not np.any(d[:, 0] ** 2 + d[:, 1] ** 2 + d[:, 2] ** 2 <= buf_sq):
                    first_outside = j
                    break
            if first_outside is None:
                del b[:]
            elif first_outside > 0:
                del b[:first_outside - 1]
            if b:
                known = np.vstack([known, _xyz(b)])
        br = [b for b in br if b]
    return _rebuild(br)


def trim_start(cl, mm):
    """centerline.rs:698-710, 917-936: remove `mm` of arc length from the start of branch 0."""
    if mm <= 0.0 or not cl.branch_start_indices:
        return _copy(cl)
    br = _branches(cl)
    if len(br[0]) > 1:
        d = _consecutive(_xyz(br[0]))
        arc, trim = 0.0, 0
        for i in range(1, len(br[0])):
            arc += float(d[i - 1])
            if arc <= mm:
                trim = i
            else:
                break
        if trim > 0:
            del br[0][:trim]
    return _rebuild(br)


def _resample_branch(points, spacing):
    """centerline.rs:730-790."""
    t = _types()
    if len(points) < 2:
        return points
    d = _consecutive(_xyz(points))
    cum = [0.0]
    for v in d.tolist():
        cum.append(cum[-1] + v)
    total = cum[-1]
    if total < 1e-12:
        return points
    targets, s = [], 0.0
    while s < total:
        targets.append(s)
        s += spacing
    targets.append(total)
    out, seg = [], 0
    for k, tt in enumerate(targets):
        while seg < len(points) - 2 and cum[seg + 1] < tt:
            seg += 1
        s0, s1 = cum[seg], cum[seg + 1]
        frac = 0.0 if abs(s1 - s0) < 1e-12 else (tt - s0) / (s1 - s0)
        p0, p1 = points[seg].contour_point, points[seg + 1].contour_point
        q = t.PyCenterlinePoint(t.PyContourPoint(k, k, p0.x + frac * (p1.x - p0.x), p0.y + frac * (p1.y - p0.y),
                                                 p0.z + frac * (p1.z - p0.z), p0.aortic), (0.0, 0.0, 0.0),
                                points[seg].branch_id)
        q.radius = points[seg].radius + frac * (points[seg + 1].radius - points[seg].radius)
        out.append(q)
    return out


def resample(cl, spacing_mm):
    """centerline.rs:717-727: every branch to even arc-length spacing (linear interpolation, end point kept)."""
    if not cl.points or spacing_mm <= 1e-12:
        return _copy(cl)
    return _rebuild([_resample_branch(b, spacing_mm) for b in _branches(cl)])


def smooth(cl, sigma):
    """centerline.rs:798-866: Gaussian over the point index, per branch, window cut symmetrically at 3 sigma and at
    the branch ends (so straight lines stay put)."""
    if not cl.points or sigma < 1e-12:
        return _copy(cl)
    pts = [_clone_point(p) for p in cl.points]
    xyz = _xyz(pts)
    new = xyz.copy()
    radius = int(math.ceil(3.0 * sigma))
    ids = [p.branch_id for p in pts]
    for bid in range(max(ids) + 1):
        idx = [i for i, b in enumerate(ids) if b == bid]
        m = len(idx)
        for li, gi in enumerate(idx):
            r = min(li, radius, m - 1 - li)
            wx = wy = wz = wt = 0.0
            for j in range(li - r, li + r + 1):
                diff = float(li) - float(j)
                w = math.exp(-0.5 * diff * diff / (sigma * sigma))
                x, y, z = xyz[idx[j]]
                wx += w * x
                wy += w * y
                wz += w * z
                wt += w
            if wt > 1e-12:
                new[gi] = (wx / wt, wy / wt, wz / wt)
    for p, (x, y, z) in zip(pts, new.tolist()):
        p.contour_point.x, p.contour_point.y, p.contour_point.z = x, y, z
    _recompute_tangents(pts)
    return _make(pts, cl.branch_start_indices)


def _copy(cl):
    return _make([_clone_point(p) for p in cl.points], cl.branch_start_indices)


# ---- loading and the standard preparation chain (multimodars/ccta/centerline_prep.py:10-140) ------------------------
def load_centerline(source, name):
    """A PyCenterline (returned as it is), an (N, 3) array, a `.vtp` path or a comma-separated x,y,z file ->
    PyCenterline, not yet prepared. `name` only labels the progress line, as in the reference."""
    from ._converters import numpy_to_centerline
    from ._vtp import read_centerline_vtp
    t = _types()
    if isinstance(source, t.PyCenterline):
        cl, how = source, f"Using provided {name} centerline"
    elif isinstance(source, np.ndarray):
        cl, how = numpy_to_centerline(source), f"Using provided {name} centerline"
    else:
        is_vtp = str(source).lower().endswith(".vtp")
        try:
            cl = read_centerline_vtp(str(source)) if is_vtp else numpy_to_centerline(np.genfromtxt(source, delimiter=","))
        except Exception as e:
            print(f"Error reading {name} centerline from {source}: {e}")
            raise
        how = f"Loaded {name} centerline from VTP" if is_vtp else f"Loaded {name} centerline"
    print(f"{how}: {len(cl.points)} points")
    return cl


def prepare_centerline(centerline, ref_centerline=None, spacing_mm=None, branch_spacing_tolerance=2.0, rm_start_mm=0.0,
                       smooth_sigma=2.5):
    """Branch detection (only for a coronary, i.e. when a reference is given, that has no branch structure yet) ->
    overlap removal -> optional inlet trim -> optional resampling -> orientation (towards the reference's branch 0, or
    highest z first without one) -> optional smoothing."""
    cl = centerline
    if ref_centerline is not None and len(cl.branch_start_indices) <= 1:
        cl = calculate_branches(cl, branch_spacing_tolerance)
    cl = remove_branch_overlap(cl)
    if rm_start_mm > 0:
        cl = trim_start(cl, rm_start_mm)
    if spacing_mm:
        cl = resample(cl, spacing_mm)
    cl = orient_to_reference(cl, ref_centerline) if ref_centerline is not None else orient_by_max_z(cl)
    if smooth_sigma > 0:
        cl = smooth(cl, smooth_sigma)
    return cl
