"""Array <-> object converters of the reference's Python layer (multimodars/_converters.py:19-687): `to_array`,
`numpy_to_geometry`, `numpy_to_centerline` (numpy_to_inputdata lives in _types.py). Same names, arguments, defaults,
return shapes and error messages; built on the (n, 6) row arrays the value types keep, so no per-point objects are
created on the way."""
from __future__ import annotations

import numpy as np

from ._types import (PyCenterline, PyContour, PyContourPoint, PyFrame, PyGeometry, PyGeometryPair, PyInputData,
                     PyRecord,
                     numpy_to_inputdata)  # noqa: F401  (the reference exports numpy_to_inputdata from this module)

_LAYERS = ("lumen", "eem", "calcification", "sidebranch", "catheter", "wall")


def _xyz_rows(contour) -> np.ndarray:
    r = contour.points_array()
    return r[:, [0, 2, 3, 4]].astype(float) if len(r) else np.zeros((0, 4), dtype=float)


def _frame_dict(frame) -> dict:
    out = {"lumen": _xyz_rows(frame.lumen)}
    for kind, c in frame.extras.items():
        out[str(kind).lower()] = _xyz_rows(c)
    rp = frame.reference_point
    out["reference"] = (np.array([[rp.frame_index, rp.x, rp.y, rp.z]], dtype=float) if rp is not None
                        else np.zeros((0, 4), dtype=float))
    return out


def _geometry_dict(geom) -> dict:
    parts = {k: [] for k in (*_LAYERS, "reference")}
    for f in geom.frames:
        for k, a in _frame_dict(f).items():
            if k in parts and len(a):
                parts[k].append(a)
    return {k: (np.vstack(v) if v else np.zeros((0, 4), dtype=float)) for k, v in parts.items()}


def _input_dict(inp) -> dict:
    out = {k: np.zeros((0, 4), dtype=float) for k in ("lumen", "eem", "calcification", "sidebranch", "reference")}
    out["diastole"], out["label"] = inp.diastole, inp.label
    for k in ("lumen", "eem", "calcification", "sidebranch"):
        cs = getattr(inp, k)
        if cs:
            rows = [_xyz_rows(c) for c in cs if len(c)]
            if rows:
                out[k] = np.vstack(rows)
    rp = inp.ref_point
    out["reference"] = np.array([[rp.frame_index, rp.x, rp.y, rp.z]], dtype=float)
    if inp.record:
        out["records"] = np.array([[r.frame, r.phase, np.nan if r.measurement_1 is None else r.measurement_1,
                                    np.nan if r.measurement_2 is None else r.measurement_2] for r in inp.record],
                                  dtype=object)
    return out


def to_array(generic):
    """_converters.py:19-92. PyContour / PyCenterline -> (N, 4) [frame_index, x, y, z]; PyFrame / PyGeometry -> dict of
    such arrays per layer plus "reference"; PyGeometryPair -> (dict, dict); PyInputData -> dict with metadata."""
    if isinstance(generic, PyContour):
        return _xyz_rows(generic)
    if isinstance(generic, PyCenterline):
        return np.array([(p.contour_point.frame_index, p.contour_point.x, p.contour_point.y, p.contour_point.z)
                         for p in generic.points], dtype=float)
    if isinstance(generic, PyFrame):
        return _frame_dict(generic)
    if isinstance(generic, PyGeometry):
        return _geometry_dict(generic)
    if isinstance(generic, PyGeometryPair):
        return _geometry_dict(generic.geom_a), _geometry_dict(generic.geom_b)
    if isinstance(generic, PyInputData):
        return _input_dict(generic)
    raise TypeError(f"Unsupported type for to_array: {type(generic)}")


def geometry_to_frames_array(geometry) -> dict:
    """_converters.py:967-1015: {str(frame.id): {layer: (N, 4) array, ..., "reference": (0|1, 4) array}}."""
    return {str(f.id): _frame_dict(f) for f in geometry.frames}


def array_to_pyinputdata(lumen=None, eem=None, calcification=None, sidebranch=None, records=None, reference=None,
                         diastole=True, label=""):
    """_converters.py:689-966: the inverse of `to_array(PyInputData)`. Layers are lists of PyContour (taken as they
    are) or (N, 4) [frame, x, y, z] arrays grouped by ascending frame id; records are PyRecord lists, row lists, plain
    or structured arrays; the reference point is the first non-zero row of `reference` ((0, 0, 0) on frame 0 if None)."""
    def layer(v, kind):
        if v is None:
            return []
        if isinstance(v, list) and v and hasattr(v[0], "points") and hasattr(v[0], "id"):
            return v
        a = np.asarray(v, dtype=object) if isinstance(v, (list, tuple)) else np.asarray(v)
        if a.dtype.names:
            try:
                a = np.vstack([a[n] for n in a.dtype.names]).T
            except Exception as e:
                raise ValueError(f"Could not convert structured array for layer: {e}") from None
        if a.size == 0:
            return []
        if a.ndim == 1:
            if a.shape[0] != 4:
                raise ValueError(f"layer 1D array must have length 4, got {a.shape}")
            a = a[np.newaxis, :]
        if a.ndim != 2 or a.shape[1] < 4:
            raise ValueError(f"layer must be (N,4)-like, got shape {a.shape}")
        a = a[:, :4].astype(float)
        frames = a[:, 0].astype(np.int64)
        out = []
        for fid in np.unique(frames).tolist():
            sel = a[frames == fid]
            rows = np.zeros((len(sel), 6))
            rows[:, 0] = fid
            rows[:, 1] = np.arange(len(sel))
            rows[:, 2:5] = sel[:, 1:4]
            out.append(PyContour(fid, fid, rows, (float(np.mean(sel[:, 1])), float(np.mean(sel[:, 2])),
                                                  float(np.mean(sel[:, 3]))), None, None, kind))
        return out

    def recs(r):
        if r is None:
            return None
        if isinstance(r, (list, tuple)):
            out = []
            for item in r:
                if hasattr(item, "frame") and hasattr(item, "phase"):
                    out.append(item)
                else:
                    out.append(PyRecord(int(item[0]), str(item[1]),
                                        None if len(item) < 3 or item[2] is None else float(item[2]),
                                        None if len(item) < 4 or item[3] is None else float(item[3])))
            return out
        if isinstance(r, np.ndarray):
            if r.dtype.names:
                low = {n.lower(): n for n in r.dtype.names}
                if "frame" not in low or "phase" not in low:
                    raise ValueError("Structured records must contain 'frame' and 'phase'")
                m1 = low.get("measurement_1", low.get("m1"))
                m2 = low.get("measurement_2", low.get("m2"))
                return [PyRecord(int(r[low["frame"]][i]), str(r[low["phase"]][i]),
                                 None if m1 is None else float(r[m1][i]), None if m2 is None else float(r[m2][i]))
                        for i in range(len(r))]
            a = r[np.newaxis, :] if r.ndim == 1 else r

            def opt(row, k):
                v = row[k] if len(row) > k else None
                return None if v is None or (isinstance(v, float) and np.isnan(v)) else float(v)

            return [PyRecord(int(row[0]), str(row[1]), opt(row, 2), opt(row, 3)) for row in a]
        raise ValueError("Unsupported records format")

    if reference is None:
        ref = PyContourPoint(0, 0, 0.0, 0.0, 0.0, False)
    else:
        a = np.asarray(reference)
        if a.ndim == 1:
            if a.shape[0] < 4:
                raise ValueError("reference must be length 4 or shape (1,4)")
            row = a[:4]
        else:
            if a.shape[1] < 4:
                raise ValueError("reference must be (N,4)-like")
            nz = np.any(a != 0, axis=1)
            row = (a[nz][0] if np.any(nz) else a[0])[:4]
        ref = PyContourPoint(int(row[0]), 0, float(row[1]), float(row[2]), float(row[3]), False)

    return PyInputData(layer(lumen, "Lumen"), layer(eem, "Eem") or None, layer(calcification, "Calcification") or None,
                       layer(sidebranch, "Sidebranch") or None, recs(records), ref, bool(diastole), str(label))


def _numeric(arr, name):
    if arr is None:
        return np.zeros((0, 4), dtype=float)
    try:
        a = np.asarray(arr, dtype=float)
    except (TypeError, ValueError) as e:
        raise ValueError(f"{name} must be convertible to a numeric array: {e}") from None
    return a


def numpy_to_geometry(lumen_arr, eem_arr=None, catheter_arr=None, wall_arr=None, reference_arr=None, label=""):
    """_converters.py:440-602: rows [frame_index, x, y, z] grouped by frame index into frames (ascending); every frame
    carries the same reference point (the first row of `reference_arr`); contour centroids are the coordinate means."""
    layers = {"Lumen": _numeric(lumen_arr, "lumen_arr"), "Eem": _numeric(eem_arr, "eem_arr"),
              "Catheter": _numeric(catheter_arr, "catheter_arr"), "Wall": _numeric(wall_arr, "wall_arr")}
    ref = _numeric(reference_arr, "reference_arr")
    if layers["Lumen"].size == 0:
        raise ValueError("lumen_arr cannot be empty")
    rp = None
    if ref.size > 0:
        fr, x, y, z = (ref if ref.ndim == 1 else ref[0])[:4]
        rp = (int(fr), 0, float(x), float(y), float(z), False)
    def grouped(kind):
        """{frame id: PyContour} of one layer in ONE pass (rows keep their order inside a frame; every contour is a
        row-slice view of one (N, 6) array) instead of one boolean mask per frame and layer."""
        a = layers[kind]
        if a.size == 0:
            return {}
        frames = a[:, 0].astype(int)
        n = len(frames)
        if n > 1 and not np.all(frames[1:] >= frames[:-1]):
            order = np.argsort(frames, kind="stable")
            a, frames = a[order], frames[order]
        starts = np.concatenate(([0], np.flatnonzero(np.diff(frames)) + 1, [n]))
        big = np.empty((n, 6))
        big[:, 0] = frames
        big[:, 1] = np.arange(n) - np.repeat(starts[:-1], np.diff(starts))
        big[:, 2:5] = a[:, 1:4]
        big[:, 5] = 0.0
        out = {}
        for s0, e0 in zip(starts[:-1].tolist(), starts[1:].tolist()):
            fid = int(frames[s0])
            rows = big[s0:e0]
            c = (float(np.mean(rows[:, 2])), float(np.mean(rows[:, 3])), float(np.mean(rows[:, 4])))
            out[fid] = PyContour(fid, fid, rows, c, None, None, kind)
        return out

    by_kind = {k: grouped(k) for k in ("Lumen", "Eem", "Catheter", "Wall")}
    frames = []
    for fid in sorted(by_kind["Lumen"]):          # a frame without a lumen contour is skipped, like the reference
        lum = by_kind["Lumen"][fid]
        extras = {k: by_kind[k][fid] for k in ("Eem", "Catheter", "Wall") if fid in by_kind[k]}
        frames.append(PyFrame(fid, lum.centroid, lum, extras, None if rp is None else PyContourPoint(*rp)))
    return PyGeometry(frames, label)


def numpy_to_centerline(arr, aortic=False):
    """_converters.py:605-686: (N, 3) [x, y, z] -> PyCenterline; NaNs are linearly interpolated along the index axis
    (edges take the nearest valid value)."""
    a = np.asarray(arr, dtype=float)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError("Input must be a (N,3) array")
    n = a.shape[0]
    if n == 0:
        raise ValueError("Input array must contain at least one point")
    if np.isnan(a).any():
        a = a.copy()
        idx = np.arange(n)
        for col in range(3):
            ok = ~np.isnan(a[:, col])
            if not ok.any():
                raise ValueError(f"All values are NaN for coordinate column {col}; cannot build centerline.")
            if not ok.all():
                a[:, col] = np.interp(idx, idx[ok], a[ok, col])
    if n < 2:
        raise ValueError("Centerline must contain at least two points after cleaning/interpolation.")
    pts = [PyContourPoint(i, i, float(x), float(y), float(z), aortic) for i, (x, y, z) in enumerate(a.tolist())]
    return PyCenterline.from_contour_points(pts)
