"""Known answers of the reference's Rust unit tests for its value types, held against the Python value types of this
package (multimodars/_types.py): types/native/contour.rs:606-1040, types/native/frame.rs:214-640 (through the
PyFrame surface: degrees, about the frame centroid — py_frame.rs:90-116)."""
import math

from multimodars import PyContour, PyContourPoint, PyFrame, PyGeometry


def contour(xy, cid=1, centroid=None, kind="Lumen", z=0.0):
    pts = [PyContourPoint(cid, i, float(x), float(y), z, False) for i, (x, y) in enumerate(xy)]
    return PyContour(cid, cid, pts, centroid, None, None, kind)


def xy(c):
    return [(p.x, p.y) for p in c.points]


def test_compute_centroid():  # contour.rs:606-654
    c = contour([(0, 0), (2, 0), (2, 2), (0, 2)])
    c.compute_centroid()
    assert c.centroid == (1.0, 1.0, 0.0)


def test_farthest_points():  # contour.rs:657-707
    (a, b), d = contour([(0, 0), (2, 0), (2, 2), (0, 2)], centroid=(1.0, 1.0, 0.0)).find_farthest_points()
    assert abs(d - math.sqrt(8.0)) < 1e-6
    assert {(a.x, a.y), (b.x, b.y)} == {(0.0, 0.0), (2.0, 2.0)}


def test_closest_opposite():  # contour.rs:710-761
    c = contour([(0, 1), (1, 0), (0, -0.5), (-1, 0)], centroid=(0.0, 0.125, 0.0))
    (a, b), d = c.find_closest_opposite()
    assert abs(d - 1.5) < 1e-6 and {a.point_index, b.point_index} == {0, 2}


def test_sort_contour_points():  # contour.rs:764-831: highest y first, then counter-clockwise
    s = contour([(-2, 0), (0, 2), (2, 0), (0, -2)], centroid=(0.0, 0.0, 0.0)).sort_contour_points()
    assert xy(s) == [(0.0, 2.0), (-2.0, 0.0), (0.0, -2.0), (2.0, 0.0)]
    assert [p.point_index for p in s.points] == [0, 1, 2, 3]


def test_areas():  # contour.rs:834-972
    assert abs(contour([(0, 0), (3, 0), (0, 4)]).get_area() - 6.0) < 1e-12
    assert abs(contour([(0, 0), (2, 0), (2, 2), (0, 2)]).get_area() - 4.0) < 1e-12          # counter-clockwise
    assert abs(contour([(0, 0), (0, 2), (2, 2), (2, 0)]).get_area() - 4.0) < 1e-12          # clockwise: same
    assert contour([(0, 0), (1, 1)]).get_area() == 0.0 and contour([(0, 0)]).get_area() == 0.0


def test_elliptic_ratio_and_area():  # contour.rs:983-1040
    c = contour([(1, 0), (0, 2), (1, 4), (2, 2)], centroid=(1.0, 2.0, 0.0))
    assert abs(c.get_elliptic_ratio() - 2.0) < 1e-6 and abs(c.get_area() - 4.0) < 1e-6


def test_contour_rotate_translate_leave_the_original():  # py_contour.rs:216-262
    c = contour([(1, 0), (0, 1), (-1, 0), (0, -1)], centroid=(0.0, 0.0, 0.0))
    r = c.rotate(90.0)
    assert xy(c) == [(1.0, 0.0), (0.0, 1.0), (-1.0, 0.0), (0.0, -1.0)]
    for (x, y), (wx, wy) in zip(xy(r), [(0, 1), (-1, 0), (0, -1), (1, 0)]):
        assert abs(x - wx) < 1e-12 and abs(y - wy) < 1e-12
    t = c.translate(1.0, 2.0, 3.0)
    assert [(p.x, p.y, p.z) for p in t.points] == [(2.0, 2.0, 3.0), (1.0, 3.0, 3.0), (0.0, 2.0, 3.0), (1.0, 1.0, 3.0)]
    assert t.centroid == (0.0, 0.0, 0.0)          # Contour::translate moves the points only (contour.rs:60-66)


def _frame():
    lumen = contour([(0, 2), (2, 4), (4, 2), (2, 0)], 1, (2.0, 2.0, 0.0))
    eem = contour([(-1, 2), (2, 5), (5, 2), (0, -1)], 2, None, "Eem")
    return PyFrame(1, (1.0, 1.0, 0.0), lumen, {"Eem": eem}, PyContourPoint(1, 0, 0.0, 4.0, 0.0, False))


def _extra(frame, kind):
    return next(c for k, c in frame.extras.items() if str(k) == kind)


def test_frame_rotate_with_eem_90deg():  # frame.rs:214-445: lumen, extras and reference point turn about frame.centroid
    f = _frame()
    r = f.rotate(90.0)
    for got, want in ((xy(r.lumen), [(0, 0), (-2, 2), (0, 4), (2, 2)]),
                      (xy(_extra(r, "Eem")), [(0, -1), (-3, 2), (0, 5), (3, 0)]),
                      ([(r.reference_point.x, r.reference_point.y)], [(-2, 0)])):
        for (x, y), (wx, wy) in zip(got, want):
            assert abs(x - wx) < 1e-6 and abs(y - wy) < 1e-6
    back = r.rotate(-90.0)
    for got, want in ((xy(back.lumen), xy(f.lumen)), (xy(_extra(back, "Eem")), xy(_extra(f, "Eem"))),
                      ([(back.reference_point.x, back.reference_point.y)], [(0.0, 4.0)])):
        for (x, y), (wx, wy) in zip(got, want):
            assert abs(x - wx) < 1e-6 and abs(y - wy) < 1e-6
    assert xy(f.lumen) == [(0.0, 2.0), (2.0, 4.0), (4.0, 2.0), (2.0, 0.0)]                   # the input is untouched


def test_frame_translate_with_eem_and_reference():  # frame.rs:555-695
    lumen = contour([(0, 0), (2, 0), (2, 2), (0, 2)], 1, (1.0, 1.0, 0.0))
    eem = contour([(-1, 2), (2, 5), (5, 2), (0, -1)], 2, None, "Eem")
    f = PyFrame(1, (1.0, 1.0, 0.0), lumen, {"Eem": eem}, PyContourPoint(1, 0, 0.5, -0.5, 0.0, False))
    t = f.translate(1.0, 2.0, 3.0)
    assert t.centroid == (2.0, 3.0, 3.0)
    assert [(p.x, p.y, p.z) for p in t.lumen.points] == [(1.0, 2.0, 3.0), (3.0, 2.0, 3.0), (3.0, 4.0, 3.0), (1.0, 4.0, 3.0)]
    assert [(p.x, p.y, p.z) for p in _extra(t, "Eem").points] == [(0.0, 4.0, 3.0), (3.0, 7.0, 3.0), (6.0, 4.0, 3.0),
                                                                   (1.0, 1.0, 3.0)]
    rp = t.reference_point
    assert (rp.x, rp.y, rp.z) == (1.5, 1.5, 3.0)
    assert t.lumen.centroid == (2.0, 3.0, 3.0)                                               # recomputed, frame.rs:20


def test_geometry_rotate_is_per_frame_and_a_copy():
    g = PyGeometry([_frame()], "g")
    r = g.rotate(90.0)
    assert abs(r.frames[0].lumen.points[1].x + 2.0) < 1e-9 and g.frames[0].lumen.points[1].x == 2.0
    assert r.label == "g" and len(r) == 1


def test_farthest_points_and_area_equal_the_reference_loops():
    """find_farthest_points / get_area / get_elliptic_ratio run in the library (mmrs_contour_metrics) with vectorised
    numpy twins; both must return exactly what the reference's loops do
    (contour.rs:227-243: strict `>` on the rooted distances in (i, j > i) order, starting from (p0, p0), 0.0;
    contour.rs:345-363: sequential sums of the cross-product terms) — including exact ties and degenerate sets."""
    import numpy as np
    rng = np.random.default_rng(3)

    def loops(p):
        best, md = (0, 0), 0.0
        for i in range(len(p)):
            for j in range(i + 1, len(p)):
                d = math.sqrt((p[i][0] - p[j][0]) ** 2 + (p[i][1] - p[j][1]) ** 2 + (p[i][2] - p[j][2]) ** 2)
                if d > md:
                    md, best = d, (i, j)
        cx = cy = cz = 0.0
        for i in range(len(p)):
            a, b = p[i], p[(i + 1) % len(p)]
            cx += a[1] * b[2] - a[2] * b[1]
            cy += a[2] * b[0] - a[0] * b[2]
            cz += a[0] * b[1] - a[1] * b[0]
        return best, md, (0.5 * math.sqrt(cx * cx + cy * cy + cz * cz) if len(p) >= 3 else 0.0)

    for trial in range(40):
        n = int(rng.integers(1, 150))
        if trial % 4 == 0:
            ang = 2 * np.pi * np.arange(n) / n
            xyz = np.stack([2 * np.cos(ang), 2 * np.sin(ang), np.zeros(n)], 1)       # many (nearly) tied diameters
        elif trial % 4 == 1:
            xyz = np.round(rng.normal(0, 2, (n, 3)))                                  # integer grid: exact ties
        elif trial % 4 == 2:
            xyz = np.tile(rng.normal(0, 1, (1, 3)), (n, 1))                           # all points identical
        else:
            xyz = rng.normal(0, 2, (n, 3))
        c = PyContour(0, 0, [PyContourPoint(0, i, *map(float, p), False) for i, p in enumerate(xyz)], (0.0, 0.0, 0.0))
        (i, j), md, area = loops(xyz.tolist())
        for far, ar in ((c.find_farthest_points, c.get_area), (c._farthest_numpy, c._area_numpy)):   # library, numpy twin
            (a, b), d = far()
            assert (a.point_index, b.point_index, d) == (i, j, md), (trial, n, far.__name__)
            assert ar() == area, (trial, n)
        from multimodars import _native as nat
        assert nat.contour_metrics(xyz)[0] == area       # the library's area, the third implementation
        if n > 2:                                        # elliptic ratio: farthest over the shortest (i, i + n/2) chord
            minor = min(math.sqrt(sum((xyz[k][q] - xyz[(k + n // 2) % n][q]) ** 2 for q in range(3))) for k in range(n))
            assert c._minor_numpy() == minor
            want = math.nan if (md == 0.0 and minor == 0.0) else (minor / md if md < minor else (md / minor if minor else math.inf))
            got = c.get_elliptic_ratio()
            assert got == want or (math.isnan(got) and math.isnan(want)), (trial, n, got, want)


def test_closest_opposite_equals_the_reference_loops():
    """contour.rs:247-296, literally: partner = first j != i whose angle about the centroid is nearest to opposite,
    chord = sqrt(dx*dx + dy*dy), first strictly shortest chord wins; without a stored centroid the mean of the points."""
    import numpy as np
    rng = np.random.default_rng(4)

    def loops(p, cx, cy):
        n, th = len(p), []
        for q in p:
            t = math.atan2(q[1] - cy, q[0] - cx)
            th.append(t + 2 * math.pi if t < 0 else t)
        md, best = float("inf"), (0, 1)
        for i in range(n):
            bd, bj = float("inf"), i
            for j in range(n):
                if j == i:
                    continue
                d = abs(th[j] - th[i])
                d = 2 * math.pi - d if d > math.pi else d
                if abs(d - math.pi) < bd:
                    bd, bj = abs(d - math.pi), j
            dist = math.sqrt((p[i][0] - p[bj][0]) ** 2 + (p[i][1] - p[bj][1]) ** 2)
            if dist < md:
                md, best = dist, (i, bj)
        return best, md

    for t in range(45):
        n = int(rng.integers(3, 150))
        if t % 3 == 0:
            ang = 2 * np.pi * np.arange(n) / n
            xy = np.stack([2 * np.cos(ang), 1.5 * np.sin(ang)], 1)
        elif t % 3 == 1:
            xy = np.round(rng.normal(0, 2, (n, 2)))
        else:
            xy = rng.normal(0, 2, (n, 2))
        pts = [PyContourPoint(0, i, float(a), float(b), 0.0, False) for i, (a, b) in enumerate(xy)]
        for centroid in ((0.1, -0.2, 0.0), None):
            c = PyContour(0, 0, pts, centroid)
            cx, cy = (centroid[0], centroid[1]) if centroid else (sum(q[0] for q in xy.tolist()) / n, sum(q[1] for q in xy.tolist()) / n)
            (i, j), md = loops(xy.tolist(), cx, cy)
            for opp in (c.find_closest_opposite, c._opposite_numpy):                          # library, numpy twin
                (a, b), d = opp()
                assert (a.point_index, b.point_index, d) == (i, j, md), (t, n, centroid, opp.__name__)
