"""Array <-> object converters of the reference's Python layer (multimodars/_converters.py:19-687): `to_array`,
`numpy_to_geometry`, `numpy_to_centerline` (numpy_to_inputdata lives in _types.py). Same names, arguments, defaults,
return shapes and error messages; built on the (n, 6) row arrays the value types keep, so no per-point objects are
created on the way."""
from __future__ import annotations

import numpy as np

from ._types import (PyCenterline, PyContour, PyContourPoint, PyFrame, PyGeometry, PyGeometryPair, PyInputData,
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
    ids = sorted({int(v) for a in layers.values() if a.size for v in a[:, 0].astype(int)})

    def contour(kind, fid):
        a = layers[kind]
        if a.size == 0:
            return None
        pts = a[a[:, 0].astype(int) == fid]
        if len(pts) == 0:
            return None
        rows = np.column_stack([pts[:, 0].astype(int).astype(float), np.arange(len(pts), dtype=float), pts[:, 1:4],
                                np.zeros(len(pts))])
        c = (float(np.mean(pts[:, 1])), float(np.mean(pts[:, 2])), float(np.mean(pts[:, 3])))
        return PyContour(fid, fid, rows, c, None, None, kind)

    frames = []
    for fid in ids:
        lum = contour("Lumen", fid)
        if lum is None:
            continue
        extras = {k: c for k in ("Eem", "Catheter", "Wall") if (c := contour(k, fid)) is not None}
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
