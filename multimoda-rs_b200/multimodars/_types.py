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
        key = name.strip().lower()
        for t in PyContourType:
            if t.name.lower() == key:
                return t
        raise ValueError(f"Unknown contour type: {name}")

    @staticmethod
    def all_types():
        return list(PyContourType)

    def __str__(self):
        return self.name


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

    # the reference's `points: Vec<PyContourPoint>` attribute
    @property
    def points(self):
        if self._pts is None:
            self._pts = [PyContourPoint(*r) for r in self._rows]
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
            sx = sy = sz = 0.0
            for row in r:  # sequential fold, contour.rs:213-224
                sx += row[2]
                sy += row[3]
                sz += row[4]
            n = float(len(r))
            self.centroid = (sx / n, sy / n, sz / n)

    def points_as_tuples(self):
        return [(float(r[2]), float(r[3]), float(r[4])) for r in self._sync()]


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

    def get_lumen_contours(self):
        return [f.lumen for f in self.frames]

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
        pos = 0
        nf = int(b[0])
        pos = 1
        frames = []
        for _ in range(nf):
            fid, cx, cy, cz, has_ref = b[pos:pos + 5]
            rp = b[pos + 5:pos + 11]
            nc = int(b[pos + 11])
            pos += 12
            lumen, extras = None, {}
            for k in range(nc):
                h = b[pos:pos + 12]
                n = int(h[11])
                rows = b[pos + 12:pos + 12 + 6 * n].reshape(n, 6)   # a view: the blob stays alive through it
                pos += 12 + 6 * n
                c = PyContour(int(h[1]), int(h[2]), rows, tuple(h[4:7]) if h[3] else None,
                              float(h[8]) if h[7] else None, float(h[10]) if h[9] else None, KIND_NAMES[int(h[0])])
                if k == 0:
                    lumen = c
                else:
                    extras[c.kind] = c
            frames.append(PyFrame(int(fid), (cx, cy, cz), lumen, extras, PyContourPoint(*rp) if has_ref else None))
        if pos != len(b):
            raise ValueError("geometry blob not fully consumed")
        return PyGeometry(frames, label)


class PyGeometryPair:
    """src/types/binding/py_geometry_pair.rs:26-68."""

    def __init__(self, geom_a, geom_b, label):
        self.geom_a, self.geom_b, self.label = geom_a, geom_b, str(label)

    def __repr__(self):
        return f"GeometryPair(label='{self.label}', geom_a={self.geom_a!r}, geom_b={self.geom_b!r})"


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
        rows = np.concatenate([c._sync() for c in contours], axis=0)
        return np.ascontiguousarray(rows[:, [0, 2, 3, 4]])

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


def numpy_to_inputdata(lumen_arr, ref_point, diastole, record=None, eem_arr=None, calcification=None,
                       sidebranch=None, label=""):
    """multimodars/_converters.py:204-437 — (N,4) [frame, x, y, z] arrays -> PyInputData, one
    PyContour per frame id, built from array slices (no per-point Python objects)."""
    def num(a):
        if a is None:
            return None
        a = np.asarray(a, dtype=float)
        return a.reshape(1, -1) if a.ndim == 1 else a

    def contours(a, kind):
        if a is None:
            return None
        out = []
        frames = a[:, 0].astype(np.int64)
        order = np.argsort(frames, kind="stable")          # one pass: group rows by frame id, keeping row order
        sf = frames[order]
        cuts = np.flatnonzero(np.diff(sf)) + 1
        for idx in np.split(order, cuts):
            sel = a[idx]
            fid = int(sel[0, 0])
            rows = np.zeros((len(sel), 6))
            rows[:, 0] = sel[:, 0]
            rows[:, 1] = np.arange(len(sel))
            rows[:, 2:5] = sel[:, 1:4]
            out.append(PyContour(fid, fid, rows, (float(np.mean(sel[:, 1])), float(np.mean(sel[:, 2])),
                                                  float(np.mean(sel[:, 3]))), None, None, kind))
        return out

    lumen_arr = num(lumen_arr)
    if lumen_arr is None or lumen_arr.size == 0:
        raise ValueError("lumen_arr is empty")
    rp = np.asarray(ref_point, dtype=float).reshape(-1)
    recs = None
    if record is not None:
        recs = []
        for row in np.asarray(record, dtype=object).reshape(-1, 4):
            m1 = None if row[2] is None or (isinstance(row[2], float) and math.isnan(row[2])) else float(row[2])
            m2 = None if row[3] is None or (isinstance(row[3], float) and math.isnan(row[3])) else float(row[3])
            recs.append(PyRecord(int(row[0]), str(row[1]), m1, m2))
    return PyInputData(contours(lumen_arr, "Lumen"), contours(num(eem_arr), "Eem"),
                       contours(num(calcification), "Calcification"), contours(num(sidebranch), "Sidebranch"), recs,
                       PyContourPoint(int(rp[0]), 0, rp[1], rp[2], rp[3], False), diastole, label)
