"""Value types mirroring the reference's PyO3 classes (src/types/binding/*.rs):
PyContourPoint, PyContour, PyContourType, PyFrame, PyGeometry, PyGeometryPair,
PyInputData, PyRecord — same constructor arguments, attributes and reprs.

Contours keep their points in one (n, 6) float64 array
[frame_index, point_index, x, y, z, aortic] (the row layout of the geometry
blob, include/mmrs_b200.h); the `points` list of PyContourPoint objects the
reference exposes is materialised lazily, so array-heavy callers never pay for
one Python object per point."""
from __future__ import annotations

import enum
import math

import numpy as np

KIND_NAMES = ["Lumen", "Eem", "Calcification", "Sidebranch", "Catheter", "Wall"]
KIND_ID = {n: i for i, n in enumerate(KIND_NAMES)}


class PyContourType(enum.Enum):
    """src/types/binding/py_contour.rs:311-377."""
    Lumen = 0
    Eem = 1
    Calcification = 2
    Sidebranch = 3
    Catheter = 4
    Wall = 5

    @staticmethod
    def from_string(name: str) -> "PyContourType":
        key = name.lower()
        for t in PyContourType:
            if t.name.lower() == key:
                return t
        raise ValueError(f"Unknown contour type: '{name}'. Valid types are: lumen, eem, calcification, sidebranch, "
                         "catheter, wall")

    @staticmethod
    def all_types():
        return list(PyContourType)

    def __str__(self):
        return self.name

    def __repr__(self):  # py_contour.rs:357-359
        return f"PyContourType.{self.name}"


class PyContourPoint:
    """src/types/binding/py_contour_point.rs:33-97."""
    __slots__ = ("frame_index", "point_index", "x", "y", "z", "aortic")

    def __init__(self, frame_index, point_index, x, y, z, aortic):
        self.frame_index = int(frame_index)
        self.point_index = int(point_index)
        self.x, self.y, self.z = float(x), float(y), float(z)
        self.aortic = bool(aortic)

    def __repr__(self):
        return (f"Point(frame_id={self.frame_index}, pt_id={self.point_index}, x={self.x:.2f}, y={self.y:.2f}, "
                f"z={self.z:.2f}, aortic={'True' if self.aortic else 'False'})")

    __str__ = __repr__

    def distance(self, other: "PyContourPoint") -> float:
        dx, dy, dz = self.x - other.x, self.y - other.y, self.z - other.z
        return math.sqrt(dx * dx + dy * dy + dz * dz)

    def _row(self):
        return [float(self.frame_index), float(self.point_index), self.x, self.y, self.z, 1.0 if self.aortic else 0.0]

    def __eq__(self, o):
        return isinstance(o, PyContourPoint) and self._row() == o._row()


def _rows_from_points(points):
    if isinstance(points, np.ndarray):
        a = points if (points.dtype == np.float64 and points.flags.c_contiguous) else np.ascontiguousarray(points, dtype=np.float64)
        if a.ndim != 2 or a.shape[1] != 6:
            raise ValueError("point array must be (n, 6): frame_index, point_index, x, y, z, aortic")
        return a
    return np.array([p._row() for p in points], dtype=np.float64).reshape(-1, 6)


class PyContour:
    """src/types/binding/py_contour.rs:31-309."""

    def __init__(self, id, original_frame, points, centroid, aortic_thickness=None, pulmonary_thickness=None,
                 kind="Lumen"):
        self.id = int(id)
        self.original_frame = int(original_frame)
        self._rows = _rows_from_points(points)
        self._pts = None
        self.centroid = tuple(float(v) for v in centroid) if centroid is not None else (0.0, 0.0, 0.0)
        self.aortic_thickness = None if aortic_thickness is None else float(aortic_thickness)
        self.pulmonary_thickness = None if pulmonary_thickness is None else float(pulmonary_thickness)
        if isinstance(kind, PyContourType):
            kind = kind.name
        if kind not in KIND_ID:
            raise ValueError(f"Unknown contour type: {kind}")
        self.kind = kind
        self._has_centroid = centroid is not None

    @classmethod
    def _from_blob_rows(cls, h, rows):
        """Decoder fast path (PyGeometry.from_blob): `h` is the 12-value contour header as Python floats, `rows` a
        C-contiguous float64 (n, 6) view of the blob. Same state as __init__ leaves, without re-validating."""
        c = cls.__new__(cls)
        c.id = int(h[1])
        c.original_frame = int(h[2])
        c._rows = rows
        c._pts = None
        c._has_centroid = bool(h[3])
        c.centroid = (h[4], h[5], h[6]) if h[3] else (0.0, 0.0, 0.0)
        c.aortic_thickness = h[8] if h[7] else None
        c.pulmonary_thickness = h[10] if h[9] else None
        c.kind = KIND_NAMES[int(h[0])]
        return c

    # the reference's `points: Vec<PyContourPoint>` attribute
    @property
    def points(self):
        if self._pts is None:
            self._pts = [PyContourPoint(*r) for r in self._rows.tolist()]
        return self._pts

    @points.setter
    def points(self, value):
        self._rows = _rows_from_points(value)
        self._pts = None

    def _sync(self):
        if self._pts is not None:  # the user may have mutated point objects
            self._rows = _rows_from_points(self._pts)
        return self._rows

    def points_array(self) -> np.ndarray:
        """(n, 6) view [frame_index, point_index, x, y, z, aortic] (extension; no per-point objects)."""
        return self._sync()

    def __len__(self):
        return len(self._sync())

    def __repr__(self):
        return (f"Contour(id={self.id}, original_frame={self.original_frame}, points={len(self)}, "
                f"centroid=({self.centroid[0]:.2f}, {self.centroid[1]:.2f}, {self.centroid[2]:.2f}), kind={self.kind})")

    def compute_centroid(self):
        r = self._sync()
        if len(r):
            self.centroid = _centroid_rows(r)

    def points_as_tuples(self):
        return [tuple(r) for r in self._sync()[:, 2:5].tolist()]

    def _clone(self, rows=None, **kw):
        c = PyContour(self.id, self.original_frame, (self._sync() if rows is None else rows).copy(), self.centroid,
                      self.aortic_thickness, self.pulmonary_thickness, self.kind)
        c._has_centroid = getattr(self, "_has_centroid", True)
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    def _point(self, i):
        return PyContourPoint(*self._sync()[i])

    def _farthest_numpy(self):
        """contour.rs:227-243 — first pair with the largest 3-D distance."""
        r = self._sync()
        if len(r) == 0:
            raise ValueError("contour has no points")
        n = len(r)
        if n < 2:
            return (self._point(0), self._point(0)), 0.0
        x, y, z = r[:, 2], r[:, 3], r[:, 4]
        # Rows are taken in blocks of 64 so that the temporaries stay small (a full n x n matrix of a 500-point contour
        # is 2 MB per array: fresh pages every call). Block b holds the squared distances of rows [64 b, 64 b + 64) to
        # all points, with the pairs j <= i masked out.
        blocks, top = [], -1.0
        cols = np.arange(n)
        for i0 in range(0, n, 64):
            i1 = min(i0 + 64, n)
            dx, dy, dz = x[i0:i1, None] - x[None, :], y[i0:i1, None] - y[None, :], z[i0:i1, None] - z[None, :]
            d2 = dx * dx + dy * dy + dz * dz                # same order of operations as Point3D::distance_to
            d2[cols[None, :] <= np.arange(i0, i1)[:, None]] = -1.0
            blocks.append(d2)
            top = max(top, float(d2.max()))
        # The reference compares the SQUARE ROOTS with a strict `>`: two different squares can round to the same root,
        # so every pair whose square is within a few ulps of the largest is rooted, and the first largest root wins
        # (pairs in (i, j > i) row-major order, like the Rust loops).
        best, best_k = -1.0, 0
        for b, d2 in enumerate(blocks):
            flat = np.flatnonzero(d2.ravel() >= top * (1.0 - 1e-15))
            if len(flat):
                roots = np.sqrt(d2.ravel()[flat])
                m = int(np.argmax(roots))
                if float(roots[m]) > best:
                    best, best_k = float(roots[m]), b * 64 * n + int(flat[m])
        if not best > 0.0:                                   # nothing beats the initial (points[0], points[0]), 0.0
            return (self._point(0), self._point(0)), 0.0
        i, j = divmod(best_k, n)
        return (self._point(i), self._point(j)), best

    def _opposite_numpy(self):
        """contour.rs:247-309 — the pair closest to opposite (by angle about the centroid) with the shortest chord."""
        r = self._sync()
        n = len(r)
        if n <= 2:
            raise ValueError("Need at least 3 points")
        if getattr(self, "_has_centroid", True):
            cx, cy = self.centroid[0], self.centroid[1]
        else:                                               # no centroid stored: the sequential mean of the points
            cx, cy = float(np.cumsum(r[:, 2])[-1]) / n, float(np.cumsum(r[:, 3])[-1]) / n
        x, y = r[:, 2], r[:, 3]
        th = np.array(list(map(math.atan2, (y - cy).tolist(), (x - cx).tolist())))       # glibc atan2, like Rust's
        th = np.where(th < 0.0, th + 2.0 * math.pi, th)
        two_pi, idx = 2.0 * math.pi, np.arange(n)
        partner = np.empty(n, dtype=np.int64)
        for i0 in range(0, n, 64):                          # blocks of rows: small temporaries
            i1 = min(i0 + 64, n)
            delta = np.abs(th[None, :] - th[i0:i1, None])
            delta = np.where(delta > math.pi, two_pi - delta, delta)
            diff = np.abs(delta - math.pi)
            diff[idx[i0:i1] - i0, idx[i0:i1]] = np.inf      # j != i
            partner[i0:i1] = np.argmin(diff, axis=1)        # first minimum == the strict `<` scan over j
        dx, dy = x - x[partner], y - y[partner]
        chord = np.sqrt(dx * dx + dy * dy)                  # Point3D::distance_2d_to
        i = int(np.argmin(chord))                           # first minimum == the strict `<` scan over i
        return (self._point(i), self._point(int(partner[i]))), float(chord[i])

    def _minor_numpy(self):
        r = self._sync()
        n = len(r)
        j = (np.arange(n) + n // 2) % n
        d = r[:, 2:5] - r[j, 2:5]
        return float(np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]).min())

    # The four measurements below run the reference's own loops in the library (mmrs_contour_metrics, host only);
    # the numpy versions above give the same values bit for bit (tests/test_type_kats_cpu.py) and serve where the
    # library has not been built.
    def _metrics(self):
        from . import _native as nat
        r = self._sync()
        try:
            return nat.contour_metrics(r[:, 2:5], self.centroid if getattr(self, "_has_centroid", True) else None)
        except (OSError, AttributeError):
            return None

    def find_farthest_points(self):
        """contour.rs:227-243 — first pair with the largest 3-D distance."""
        if len(self._sync()) == 0:
            raise ValueError("contour has no points")
        m = self._metrics()
        if m is None:
            return self._farthest_numpy()
        _, (i, j, d), _, _ = m
        return (self._point(i), self._point(j)), d

    def find_closest_opposite(self):
        """contour.rs:247-296 — the pair closest to opposite (by angle about the centroid) with the shortest chord."""
        if len(self._sync()) <= 2:
            raise ValueError("Need at least 3 points")
        m = self._metrics()
        if m is None:
            return self._opposite_numpy()
        _, _, (i, j, d), _ = m
        return (self._point(i), self._point(j)), d

    def get_elliptic_ratio(self):
        """contour.rs:313-343: farthest pair over the shortest (i, i + n/2) chord, whichever way round is >= 1
        (0 / 0 is NaN, like the reference's f64 division)."""
        if len(self._sync()) <= 2:
            raise ValueError("Need at least 3 points")
        m = self._metrics()
        major, minor = (self._farthest_numpy()[1], self._minor_numpy()) if m is None else (m[1][2], m[3])
        num, den = (minor, major) if major < minor else (major, minor)
        if den == 0.0:
            return math.nan if num == 0.0 or num != num else math.copysign(math.inf, num)
        return num / den

    def get_area(self):
        """contour.rs:345-363 — half the norm of the summed cross products (O(n): the numpy form is already cheap, and
        equal to the library's value bit for bit)."""
        return self._area_numpy()

    def _area_numpy(self):
        """contour.rs:345-363 — half the norm of the summed cross products."""
        r = self._sync()
        n = len(r)
        if n < 3:
            return 0.0
        a, b = r[:, 2:5], np.roll(r[:, 2:5], -1, axis=0)
        # cumsum adds left to right, so its last element is the reference's sequential fold, bit for bit
        cx = float(np.cumsum(a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1])[-1])
        cy = float(np.cumsum(a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2])[-1])
        cz = float(np.cumsum(a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0])[-1])
        return 0.5 * math.sqrt(cx * cx + cy * cy + cz * cz)

    def rotate(self, angle_deg):
        """py_contour.rs:215-224 — about the contour's own (recomputed) centroid."""
        c = self._clone()
        c.compute_centroid()
        c._has_centroid = True
        _rotate_rows(c._rows, angle_deg * (math.pi / 180.0), c.centroid[0], c.centroid[1])
        return c

    def translate(self, dx, dy, dz):
        """py_contour.rs:245-249."""
        c = self._clone()
        c._rows[:, 2] = c._rows[:, 2] + dx
        c._rows[:, 3] = c._rows[:, 3] + dy
        c._rows[:, 4] = c._rows[:, 4] + dz
        return c

    def sort_contour_points(self):
        """contour.rs:368-405."""
        c = self._clone()
        c._rows = _sort_rows(c._rows)
        return c


def _rotate_rows(rows, angle, cx, cy):
    """ContourPoint::rotate on an (n, 6) row block, in place (contour_point.rs:38-52)."""
    if angle == 0.0 or len(rows) == 0:
        return
    ca, sa = math.cos(angle), math.sin(angle)
    x, y = rows[:, 2] - cx, rows[:, 3] - cy
    rows[:, 2], rows[:, 3] = x * ca - y * sa + cx, x * sa + y * ca + cy


def _sort_rows(rows):
    """Contour::sort_contour_points (contour.rs:368-405): stable ascending atan2 about the mean, the LAST
    highest-y point first, point_index = position."""
    n = len(rows)
    if n == 0:
        return rows
    # cumsum adds left to right: its last element is the reference's sequential sum, bit for bit
    cx, cy = float(np.cumsum(rows[:, 2])[-1]) / n, float(np.cumsum(rows[:, 3])[-1]) / n
    # math.atan2 is glibc's (what Rust's f64::atan2 calls); numpy's arctan2 may be a SIMD variant that differs in the
    # last bit, which could reorder nearly tied keys
    key = np.array(list(map(math.atan2, (rows[:, 3] - cy).tolist(), (rows[:, 2] - cx).tolist())))
    rows = rows[np.argsort(key, kind="stable")]
    y = rows[:, 3]
    if np.isnan(y).any():
        start = 0
        for i in range(1, n):
            if not (y[i] < y[start]):
                start = i
    else:
        start = n - 1 - int(np.argmax(y[::-1]))          # the LAST of the highest-y points
    rows = np.roll(rows, -start, axis=0)
    rows[:, 1] = np.arange(n)
    return rows


def _centroid_rows(rows):
    """Sequential left fold of the coordinates (contour.rs:213-224): cumsum adds left to right, so its last row is that
    fold bit for bit — unlike np.sum / np.mean, which add pairwise."""
    n = float(len(rows))
    sx, sy, sz = np.cumsum(rows[:, 2:5], axis=0)[-1].tolist()
    return (sx / n, sy / n, sz / n)


class PyFrame:
    """src/types/binding/py_frame.rs:32-130."""

    def __init__(self, id, centroid, lumen, extras=None, reference_point=None):
        self.id = int(id)
        self.centroid = tuple(float(v) for v in centroid)
        self.lumen = lumen
        self.extras = dict(extras or {})
        self.reference_point = reference_point

    def __repr__(self):
        return (f"Frame(id={self.id}, centroid=({self.centroid[0]:.2f}, {self.centroid[1]:.2f}, "
                f"{self.centroid[2]:.2f}), lumen_points={len(self.lumen)}, extras={sorted(self.extras)}, "
                f"has_reference_point={self.reference_point is not None})")

    def _clone(self):
        rp = self.reference_point
        return PyFrame(self.id, self.centroid, self.lumen._clone(), {k: c._clone() for k, c in self.extras.items()},
                       None if rp is None else PyContourPoint(*rp._row()))

    def rotate(self, angle_deg):
        """Frame::rotate about the frame centroid (py_frame.rs:89-95, frame.rs:40-63)."""
        f = self._clone()
        a = angle_deg * (math.pi / 180.0)
        if a == 0.0:
            return f
        cx, cy = f.centroid[0], f.centroid[1]
        for c in [f.lumen, *f.extras.values()]:
            _rotate_rows(c._rows, a, cx, cy)
        if f.reference_point is not None:
            row = np.array([f.reference_point._row()])
            _rotate_rows(row, a, cx, cy)
            f.reference_point = PyContourPoint(*row[0])
        x, y = f.centroid[0] - cx, f.centroid[1] - cy
        f.centroid = (x * math.cos(a) - y * math.sin(a) + cx, x * math.sin(a) + y * math.cos(a) + cy, f.centroid[2])
        return f

    def translate(self, dx, dy, dz):
        """Frame::translate (frame.rs:18-38): contour centroids are recomputed, the frame centroid shifted."""
        f = self._clone()
        for c in [f.lumen, *f.extras.values()]:
            c._rows[:, 2] = c._rows[:, 2] + dx
            c._rows[:, 3] = c._rows[:, 3] + dy
            c._rows[:, 4] = c._rows[:, 4] + dz
            if len(c._rows):
                c.centroid = _centroid_rows(c._rows)
                c._has_centroid = True
        if f.reference_point is not None:
            r = f.reference_point
            f.reference_point = PyContourPoint(r.frame_index, r.point_index, r.x + dx, r.y + dy, r.z + dz, r.aortic)
        f.centroid = (f.centroid[0] + dx, f.centroid[1] + dy, f.centroid[2] + dz)
        return f

    def sort_frame_points(self):
        f = self._clone()
        for c in [f.lumen, *f.extras.values()]:
            c._rows = _sort_rows(c._rows)
        return f


class PyGeometry:
    """src/types/binding/py_geometry.rs:24-110."""

    def __init__(self, frames, label):
        self.frames = list(frames)
        self.label = str(label)

    def __repr__(self):
        return f"Geometry({len(self.frames)} frames, label='{self.label}')"

    def __len__(self):
        return len(self.frames)

    def get_contours_by_type(self, contour_type: str):
        key = PyContourType.from_string(contour_type).name
        if key == "Lumen":
            return [f.lumen for f in self.frames]
        return [f.extras[key] for f in self.frames if key in f.extras]

    get_contours = get_contours_by_type

    def get_contours(self, contour_type: str):
        """py_geometry.rs:98-100."""
        return self.get_contours_by_type(contour_type)

    def sort_frame_points(self):
        """py_geometry.rs:152-156 -> Geometry::sort_frame_points_by_z (geometry.rs:257-276): every contour of every
        frame is rotated so that the (last) highest-z point of frame 0's lumen comes first; point_index re-assigned."""
        if not self.frames:
            return PyGeometry([], self.label)
        z = self.frames[0].lumen._sync()[:, 4]
        shift = 0
        for i in range(len(z)):          # Iterator::max_by keeps the LAST maximum
            if not (z[i] < z[shift]):
                shift = i
        out = []
        for f in self.frames:
            g = f._clone()
            for c in [g.lumen, *g.extras.values()]:
                n = len(c._rows)
                if n == 0 or shift == 0:
                    continue
                c._rows = np.roll(c._rows, -(shift % n), axis=0).copy()
                c._rows[:, 1] = np.arange(n, dtype=np.float64)
                c._pts = None
            out.append(g)
        return PyGeometry(out, self.label)

    def center_to_contour(self, contour_type):
        """py_geometry.rs:267-273 -> Geometry::center_to_contour (geometry.rs:383-442): every frame after the first is
        translated in x, y so that the centroid of its `contour_type` contour lands on frame 0's."""
        name = contour_type.name if isinstance(contour_type, PyContourType) else str(contour_type)
        if not self.frames:
            return PyGeometry([], self.label)

        def centroid_of(f):
            c = f.lumen if name == "Lumen" else f.extras.get(name)
            if c is None or len(c._sync()) == 0:
                return f.centroid
            return tuple(float(v) for v in _centroid_rows(c._sync()))

        frames = [self.frames[0]._clone()]
        c0 = frames[0].lumen if name == "Lumen" else frames[0].extras.get(name)
        if c0 is not None and len(c0._sync()):
            c0.centroid, c0._has_centroid = tuple(float(v) for v in _centroid_rows(c0._sync())), True
        ref = centroid_of(frames[0])
        for f in self.frames[1:]:
            cur = centroid_of(f)
            frames.append(f.translate(ref[0] - cur[0], ref[1] - cur[1], 0.0))
        return PyGeometry(frames, self.label)

    def get_lumen_contours(self):
        return [f.lumen for f in self.frames]

    def rotate(self, angle_deg):
        """Geometry::rotate_geometry (geometry.rs:241-250): every frame about its own centroid, then re-sorted."""
        if angle_deg * (math.pi / 180.0) == 0.0:
            return PyGeometry([f._clone() for f in self.frames], self.label)
        return PyGeometry([f.rotate(angle_deg).sort_frame_points() for f in self.frames], self.label)

    def translate(self, dx, dy, dz):
        return PyGeometry([f.translate(dx, dy, dz) for f in self.frames], self.label)

    def smooth_frames(self):
        """Geometry::smooth_frames (geometry.rs:165-239): 3-frame moving average of x, y for lumen, Eem, Wall."""
        out = []
        n = len(self.frames)
        for i, f in enumerate(self.frames):
            prev, nxt = self.frames[max(i - 1, 0)], self.frames[min(i + 1, n - 1)]
            g = f._clone()
            cnt = len(f.lumen)

            def avg(cur, p, q):
                a, b, c = cur._sync()[:cnt], p._sync()[:cnt], q._sync()[:cnt]
                if min(len(a), len(b), len(c)) < cnt:
                    raise IndexError("index out of bounds (smooth_frames)")
                rows = a.copy()
                rows[:, 2] = (b[:, 2] + a[:, 2] + c[:, 2]) / 3.0
                rows[:, 3] = (b[:, 3] + a[:, 3] + c[:, 3]) / 3.0
                o = cur._clone(rows)
                o.centroid = _centroid_rows(rows) if len(rows) else (0.0, 0.0, 0.0)
                o._has_centroid = len(rows) > 0
                return o

            g.lumen = avg(f.lumen, prev.lumen, nxt.lumen)
            for k in ("Eem", "Wall"):
                if k in f.extras and k in prev.extras and k in nxt.extras:
                    g.extras[k] = avg(f.extras[k], prev.extras[k], nxt.extras[k])
            out.append(g)
        return PyGeometry(out, self.label)

    def get_summary(self):
        """(mla, max_stenosis, stenosis_length_mm) — py_geometry.rs:190-254."""
        return self._summary([f.lumen.get_area() for f in self.frames], None)

    def _summary(self, areas, ellips):
        """`ellips` = the per-frame elliptic ratios when the caller already has them (PyGeometryPair.get_summary needs
        them for its table; each costs an all-pairs farthest-point search)."""
        if not self.frames:
            return (0.0, 0.0, 0.0)
        biggest, mla = max(areas), min(areas)
        max_sten = 1.0 - (mla / biggest) if biggest > 0.0 else 0.0
        if ellips is None:
            round_enough = all(f.lumen.get_elliptic_ratio() < 1.3 for f in self.frames)
        else:
            round_enough = all(e < 1.3 for e in ellips)
        thr = (0.70 if round_enough else 0.50) * biggest
        cen = [f.centroid for f in self.frames]
        longest, i = 0.0, 0
        while i < len(areas):
            if areas[i] < thr:
                end = i
                while end + 1 < len(areas) and areas[end + 1] < thr:
                    end += 1
                run = 0.0
                for k in range(i, end):
                    run += math.sqrt(sum((cen[k][d] - cen[k + 1][d]) ** 2 for d in range(3)))
                longest = max(longest, run)
                i = end + 1
            else:
                i += 1
        return (mla, max_sten, longest)

    def get_frame_at_z(self, z):
        if not self.frames:
            raise ValueError("geometry contains no frames")
        return min(self.frames, key=lambda f: abs(f.centroid[2] - z))

    def get_frame_at_index(self, index):
        if not 0 <= index < len(self.frames):
            raise IndexError(f"index {index} out of range for geometry with {len(self.frames)} frames")
        return self.frames[index]

    def replace_frame(self, index, frame):
        if not 0 <= index < len(self.frames):
            raise IndexError(f"index {index} is out of range for geometry with {len(self.frames)} frames")
        fr = list(self.frames)
        fr[index] = frame
        return PyGeometry(fr, self.label)

    def downsample(self, n_points):
        """py_geometry.rs:394-434: strided down-sampling (contour.rs:47-58) of every contour but the catheter."""
        def ds(c):
            rows = c._sync()
            if len(rows) <= n_points:
                return c._clone()
            step = len(rows) / n_points
            return c._clone(rows[[int(i * step) for i in range(n_points)]])

        out = []
        for f in self.frames:
            g = f._clone()
            g.lumen = ds(f.lumen)
            g.extras = {k: (c._clone() if k == "Catheter" else ds(c)) for k, c in f.extras.items()}
            out.append(g)
        return PyGeometry(out, self.label)

    # ---- blob codec ------------------------------------------------------------
    def to_blob(self) -> np.ndarray:
        out = [np.array([float(len(self.frames))])]
        for f in self.frames:
            rp = f.reference_point
            head = [float(f.id), *f.centroid, 1.0 if rp is not None else 0.0]
            head += rp._row() if rp is not None else [0.0] * 6
            cs = [f.lumen] + [f.extras[k] for k in sorted(f.extras, key=lambda k: KIND_ID[k])]
            head.append(float(len(cs)))
            out.append(np.array(head))
            for c in cs:
                rows = c._sync()
                hc = getattr(c, "_has_centroid", True)
                out.append(np.array([float(KIND_ID[c.kind]), float(c.id), float(c.original_frame),
                                     1.0 if hc else 0.0, *(c.centroid if hc else (0.0, 0.0, 0.0)),
                                     1.0 if c.aortic_thickness is not None else 0.0, c.aortic_thickness or 0.0,
                                     1.0 if c.pulmonary_thickness is not None else 0.0, c.pulmonary_thickness or 0.0,
                                     float(len(rows))]))
                out.append(rows.reshape(-1))
        return np.concatenate(out)

    @staticmethod
    def from_blob(blob, label="") -> "PyGeometry":
        b = np.asarray(blob, dtype=np.float64)
        nf = int(b[0])
        pos = 1
        frames = []
        new_contour = PyContour._from_blob_rows
        for _ in range(nf):
            # one .tolist() per header: Python floats instead of a dozen numpy scalar reads
            fid, cx, cy, cz, has_ref, r0, r1, r2, r3, r4, r5, nc = b[pos:pos + 12].tolist()
            pos += 12
            lumen, extras = None, {}
            for k in range(int(nc)):
                h = b[pos:pos + 12].tolist()
                n = int(h[11])
                rows = b[pos + 12:pos + 12 + 6 * n].reshape(n, 6)   # a view: the blob stays alive through it
                pos += 12 + 6 * n
                c = new_contour(h, rows)
                if k == 0:
                    lumen = c
                else:
                    extras[c.kind] = c
            frames.append(PyFrame(int(fid), (cx, cy, cz), lumen, extras,
                                  PyContourPoint(r0, r1, r2, r3, r4, r5) if has_ref else None))
        if pos != len(b):
            raise ValueError("geometry blob not fully consumed")
        return PyGeometry(frames, label)


class PyGeometryPair:
    """src/types/binding/py_geometry_pair.rs:26-68."""

    def __init__(self, geom_a, geom_b, label):
        self.geom_a, self.geom_b, self.label = geom_a, geom_b, str(label)

    def __repr__(self):  # py_geometry_pair.rs:48-55
        return (f"GeometryPair {self.label} (diastolic: {len(self.geom_a.frames)} frames, "
                f"systolic: {len(self.geom_b.frames)} frames)")

    def get_summary(self):
        """py_geometry_pair.rs:70-201 -> (((dia_mla, dia_max_stenosis, dia_len_mm), (sys ...)), table); table rows are
        [id, area_dia, ellip_dia, area_sys, ellip_sys, z]; the table is also printed like the reference does."""
        a, b = self.geom_a.get_lumen_contours(), self.geom_b.get_lumen_contours()
        area_a, area_b = [c.get_area() for c in a], [c.get_area() for c in b]
        ell_a, ell_b = [c.get_elliptic_ratio() for c in a], [c.get_elliptic_ratio() for c in b]   # once per contour
        dia, sys_ = self.geom_a._summary(area_a, ell_a), self.geom_b._summary(area_b, ell_b)
        n = len(a)
        if len(b) != n:
            print("ERROR: mismatched lengths between contour vectors")
        mat = [[float(a[i].id), area_a[i], ell_a[i], area_b[i], ell_b[i], a[i].centroid[2]] for i in range(n)]
        headers = ["id", "area_dia", "ellip_dia", "area_sys", "ellip_sys", "z"]
        rows = [[str(a[i].id)] + [f"{v:.2f}" for v in mat[i][1:]] for i in range(n)]
        widths = [max([len(h)] + [len(r[k]) for r in rows]) for k, h in enumerate(headers)]
        border = "+" + "".join("-" * (w + 2) + "+" for w in widths)
        print(border)
        print("|" + "".join(" " + " " * ((w - len(h)) // 2) + h + " " * ((w - len(h)) - (w - len(h)) // 2) + " |"
                            for h, w in zip(headers, widths)))
        print(border)
        for r in rows:
            print("|" + "".join(" " + c + " " * (w - len(c)) + " |" for c, w in zip(r, widths)))
        print(border)
        return (dia, sys_), mat


class PyCenterlinePoint:
    """src/types/binding/py_centerline_point.rs:8-48."""

    def __init__(self, contour_point, tangent, branch_id=0):
        self.contour_point = contour_point
        self.tangent = (float(tangent[0]), float(tangent[1]), float(tangent[2]))
        self.branch_id = int(branch_id)
        self.radius = 0.0

    def __repr__(self):
        t = self.tangent
        return (f"CenterlinePoint(point={self.contour_point!r}, tangent=({t[0]:.3f}, {t[1]:.3f}, {t[2]:.3f}), "
                f"branch={self.branch_id}, radius={self.radius:.3f})")

    __str__ = __repr__


class PyCenterline:
    """src/types/binding/py_centerline.rs:8-330: constructor, from_contour_points, __len__, points_as_tuples and the
    branch bookkeeping a raw centerline goes through before alignment (calculate_branches, find_sharp_angles, split /
    merge / get_branch, remove_branch_overlap, trim_start, resample, smooth, orient_*; `_centerline.py`)."""

    def __init__(self, points):
        self.points = list(points)
        self.branch_start_indices = [0] if self.points else []

    @staticmethod
    def from_contour_points(contour_points):
        # Centerline::from_contour_points, src/types/native/centerline.rs:14-43: tangent = normalised
        # forward difference, the last point repeats its predecessor's tangent
        pts = list(contour_points)
        out = []
        for i, cur in enumerate(pts):
            if i < len(pts) - 1:
                nx = pts[i + 1]
                dx, dy, dz = nx.x - cur.x, nx.y - cur.y, nx.z - cur.z
                n = math.sqrt(dx * dx + dy * dy + dz * dz)
                t = (dx / n, dy / n, dz / n) if n != 0.0 else (math.nan, math.nan, math.nan)
            elif out:
                t = out[i - 1].tangent
            else:
                raise IndexError("index out of bounds: the len is 0 but the index is 18446744073709551615")
            out.append(PyCenterlinePoint(cur, t, 0))
        return PyCenterline(out)

    def __len__(self):
        return len(self.points)

    def __repr__(self):  # py_centerline.rs:66-73
        from . import _centerline as c
        return (f"Centerline(len={len(self.points)}, spacing={c.mean_spacing(self):.2f} mm, "
                f"branches={len(self.branch_start_indices)})")

    __str__ = __repr__

    def points_as_tuples(self):
        return [(p.contour_point.x, p.contour_point.y, p.contour_point.z) for p in self.points]

    # Branch bookkeeping (py_centerline.rs:118-330): each returns a new PyCenterline, see _centerline.py.
    def calculate_branches(self, spacing_tolerance=1.0):
        from . import _centerline as c
        return c.calculate_branches(self, float(spacing_tolerance))

    def find_sharp_angles(self, branch_id, cos_threshold):
        from . import _centerline as c
        return c.find_sharp_angles(self, branch_id, float(cos_threshold))

    def split_branch(self, branch_id, point_index):
        from . import _centerline as c
        return c.split_branch(self, branch_id, point_index)

    def merge_branches(self, branch_id_a, branch_id_b):
        from . import _centerline as c
        return c.merge_branches(self, branch_id_a, branch_id_b)

    def get_branch(self, branch_id):
        from . import _centerline as c
        return c.get_branch(self, branch_id)

    def remove_branch_overlap(self):
        from . import _centerline as c
        return c.remove_branch_overlap(self)

    def trim_start(self, mm):
        from . import _centerline as c
        return c.trim_start(self, float(mm))

    def resample(self, spacing_mm):
        from . import _centerline as c
        return c.resample(self, float(spacing_mm))

    def smooth(self, sigma):
        from . import _centerline as c
        return c.smooth(self, float(sigma))

    def orient_by_max_z(self):
        from . import _centerline as c
        return c.orient_by_max_z(self)

    def orient_to_reference(self, reference):
        from . import _centerline as c
        return c.orient_to_reference(self, reference)

    def _rows(self):
        """(n, 8) [x, y, z, tx, ty, tz, branch_id, radius] — the row layout of mmrs_align_centerline."""
        return np.array([[p.contour_point.x, p.contour_point.y, p.contour_point.z, *p.tangent, float(p.branch_id),
                          float(p.radius)] for p in self.points], dtype=np.float64).reshape(-1, 8)


class PyRecord:
    """src/types/binding/py_record.rs:30-40."""

    def __init__(self, frame, phase, measurement_1=None, measurement_2=None):
        self.frame = int(frame)
        self.phase = str(phase)
        self.measurement_1 = measurement_1
        self.measurement_2 = measurement_2

    def __repr__(self):
        return (f"Record(frame={self.frame}, phase={self.phase}, m1={self.measurement_1}, "
                f"m2={self.measurement_2})")


class PyInputData:
    """src/types/binding/py_input_data.rs:42-101."""

    def __init__(self, lumen, eem=None, calcification=None, sidebranch=None, record=None, ref_point=None,
                 diastole=True, label=""):
        if ref_point is None:
            raise TypeError("ref_point is required")
        self.lumen = list(lumen)
        self.eem = None if eem is None else list(eem)
        self.calcification = None if calcification is None else list(calcification)
        self.sidebranch = None if sidebranch is None else list(sidebranch)
        self.record = None if record is None else list(record)
        self.ref_point = ref_point
        self.diastole = bool(diastole)
        self.label = str(label)

    def __repr__(self):
        return (f"InputData(lumen={len(self.lumen)}, eem={len(self.eem or [])}, "
                f"calcification={len(self.calcification or [])}, sidebranch={len(self.sidebranch or [])}, "
                f"record={len(self.record or [])}, ref_point={self.ref_point!r}, diastole={self.diastole}, "
                f"label='{self.label}')")

    # TryFrom<&PyInputData> for InputData (py_input_data.rs:103-172): flatten contours to point lists
    def _flat(self, contours):
        if contours is None:
            return None
        if not contours:
            return np.zeros((0, 4))
        parts = [c._sync() for c in contours]
        out = np.empty((sum(len(r) for r in parts), 4))
        o = 0
        for r in parts:                      # [frame, x, y, z] of every point, written in place (no (N, 6) temporary)
            n = len(r)
            out[o:o + n, 0] = r[:, 0]
            out[o:o + n, 1:4] = r[:, 2:5]
            o += n
        return out

    def _records(self):
        if self.record is None:
            return None
        out = np.full((len(self.record), 4), np.nan)
        for i, r in enumerate(self.record):
            out[i, 0] = r.frame
            out[i, 1] = 1.0 if r.phase == "D" else (0.0 if r.phase == "S" else -1.0)
            if r.measurement_1 is not None:
                out[i, 2] = r.measurement_1
            if r.measurement_2 is not None:
                out[i, 3] = r.measurement_2
        return out


def _records_from_array(arr):
    """_converters.py:301-356: rows [frame, phase, m1, m2] (plain, object or structured arrays; numeric phase 0 -> "D",
    anything else numeric -> "S"; NaN / unparsable measurements -> None); empty -> None."""
    if arr is None:
        return None
    arr = np.asarray(arr)
    if arr.ndim == 1 and arr.dtype.names:
        arr = np.array([tuple(r) for r in arr.tolist()], dtype=object).reshape(len(arr), -1)
    if arr.size == 0:
        return None
    if arr.ndim == 1:
        arr = arr.reshape(1, -1)

    def opt(v):
        try:
            f = float(v)
        except (TypeError, ValueError):
            return None
        return None if math.isnan(f) else f

    recs = []
    for row in arr:
        ph = row[1] if len(row) > 1 else ""
        if isinstance(ph, (bytes, bytearray)):
            ph = ph.decode("utf-8", errors="replace")
        elif isinstance(ph, (int, float, np.number)) and not isinstance(ph, bool):
            ph = "D" if int(ph) == 0 else "S"
        recs.append(PyRecord(int(row[0]), str(ph), opt(row[2]) if len(row) > 2 else None,
                             opt(row[3]) if len(row) > 3 else None))
    return recs or None


def numpy_to_inputdata(lumen_arr, ref_point, diastole, record=None, eem_arr=None, calcification=None,
                       sidebranch=None, label=""):
    """multimodars/_converters.py:204-437 — (N,4) [frame, x, y, z] arrays -> PyInputData, one
    PyContour per frame id, built from array slices (no per-point Python objects). Like the reference: the frames
    are the LUMEN's frame ids in ascending order, the other layers contribute only on those frames, layers without
    a contour become None, an unusable `ref_point` falls back to (0, 0, 0) on frame 0."""
    def num(a, name):
        if a is None:
            return np.zeros((0, 4))
        a = np.asarray(a)
        if a.ndim == 1 and a.dtype.names:
            try:
                a = np.vstack([a[n] for n in a.dtype.names]).T
            except Exception:
                raise ValueError(f"Could not convert structured array for {name}") from None
        a = np.asarray(a, dtype=float)
        return a.reshape(1, -1) if a.ndim == 1 else a

    def contours(a, kind, keep):
        if a.size == 0:
            return None
        frames = a[:, 0].astype(np.int64)
        n = len(frames)
        if n > 1 and not np.all(frames[1:] >= frames[:-1]):    # group rows by frame id, keeping row order
            order = np.argsort(frames, kind="stable")
            a, frames = a[order], frames[order]
        starts = np.concatenate(([0], np.flatnonzero(np.diff(frames)) + 1, [n]))
        # one (N, 6) [frame, point_index, x, y, z, aortic] array for the layer; every contour is a row-slice VIEW of it
        big = np.empty((n, 6))
        big[:, 0] = frames
        big[:, 1] = np.arange(n) - np.repeat(starts[:-1], np.diff(starts))
        big[:, 2:5] = a[:, 1:4]
        big[:, 5] = 0.0
        out = []
        for s, e in zip(starts[:-1].tolist(), starts[1:].tolist()):
            fid = int(frames[s])
            if keep is not None and fid not in keep:
                continue
            rows = big[s:e]
            out.append(PyContour(fid, fid, rows, (float(np.mean(rows[:, 2])), float(np.mean(rows[:, 3])),
                                                  float(np.mean(rows[:, 4]))), None, None, kind))
        return out or None

    lumen_arr = num(lumen_arr, "lumen_arr")
    eem_arr, calcification, sidebranch = num(eem_arr, "eem_arr"), num(calcification, "calcification"), num(sidebranch, "sidebranch")
    ref = None
    if ref_point is not None:
        try:
            rp = np.asarray(ref_point, dtype=float)
            rp = rp[:4] if rp.ndim == 1 else rp[0, :4]
            ref = PyContourPoint(int(rp[0]), 0, float(rp[1]), float(rp[2]), float(rp[3]), False)
        except Exception:
            ref = None
    if ref is None:
        ref = PyContourPoint(0, 0, 0.0, 0.0, 0.0, False)
    if lumen_arr.size == 0:
        raise ValueError("lumen_arr cannot be empty")
    lumen = contours(lumen_arr, "Lumen", None)
    keep = {c.id for c in lumen}
    return PyInputData(lumen, contours(eem_arr, "Eem", keep), contours(calcification, "Calcification", keep),
                       contours(sidebranch, "Sidebranch", keep), _records_from_array(record), ref, bool(diastole),
                       label or "")
