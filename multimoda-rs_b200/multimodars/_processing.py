"""The reference's public processing entry points (multimodars/_processing.py:42-1009 over
src/intravascular/binding/functions.rs:143-1433), same names, argument order, defaults and
return shapes, running on the B200 through libmmrs_b200.so (mmrs_process_cases).

`write_obj=True` (the reference's default for the from_file_* functions) writes the OBJ/MTL/PNG
files of to_object::process_case through mmrs_export_pair / mmrs_export_single."""
from __future__ import annotations

import os

import numpy as np

from . import _native as nat
from ._types import PyCenterline, PyContourType, PyGeometry, PyGeometryPair, PyInputData
from ._vtp import read_centerline_vtp  # noqa: F401  (the reference defines it in _processing.py:1355)

_ctx = None


def get_context(device: int | None = None) -> nat.Context:
    """One lazily created context per process (device = LOCAL_RANK under torchrun, else 0)."""
    global _ctx
    if _ctx is None:
        if device is None:
            device = int(os.environ.get("MMRS_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        _ctx = nat.Context(device)
    return _ctx


def _default_contour_types():
    return [PyContourType.Lumen, PyContourType.Catheter, PyContourType.Wall]


def _kind_ids(contour_types):
    cts = _default_contour_types() if contour_types is None else list(contour_types)
    return [ct.value if isinstance(ct, PyContourType) else PyContourType.from_string(str(ct)).value for ct in cts]


class _Export:
    """What the *_processing_rs functions do after alignment when write_obj is set
    (binding/entry.rs:291-350, :539-568, :673-687, :741-775)."""

    def __init__(self, write_obj, watertight, contour_types, interpolation_steps, paths):
        self.on = bool(write_obj)
        self.watertight = watertight
        self.kinds = _kind_ids(contour_types)
        self.steps = int(interpolation_steps)
        self.paths = list(paths)

    def pairs(self, pairs, blobs):
        if self.on:
            # the pairs go to different directories: written concurrently (the C side releases the GIL); the first
            # failing pair, in the reference's order, is the error that surfaces
            jobs = list(enumerate(zip(pairs, self.paths)))
            _parallel(lambda j: nat.export_pair(blobs[2 * j[0]], blobs[2 * j[0] + 1], j[1][0].geom_a.label,
                                                j[1][0].label, j[1][1], self.steps, self.watertight, self.kinds), jobs)

    def single(self, geom, blob):
        if self.on:
            nat.export_single(blob, geom.label, self.paths[0], self.watertight, self.kinds, 0)


def _logs(arr):
    # logs_to_tuples, binding/functions.rs:26-40
    return [(int(r[0]), int(r[1]), float(r[2]), float(r[3]), float(r[4]), float(r[5]), float(r[6])) for r in arr]


def _basename(path):
    name = os.path.basename(os.path.normpath(str(path)))
    return name if name else "unknown"


def _blob_from_input(inp: PyInputData, image_center, radius, n_points):
    if not isinstance(inp, PyInputData):
        raise TypeError("expected PyInputData")
    return nat.geometry_from_arrays(inp._flat(inp.lumen), np.array(
        [inp.ref_point.frame_index, inp.ref_point.x, inp.ref_point.y, inp.ref_point.z]), inp._flat(inp.eem),
        inp._flat(inp.calcification), inp._flat(inp.sidebranch), inp._records(), inp.diastole, inp.label,
        image_center, radius, n_points)


def _pair(out, i, label_a, label_b):
    return PyGeometryPair(PyGeometry.from_blob(out[i], label_a), PyGeometry.from_blob(out[i + 1], label_b),
                          f"{label_a} - {label_b}")


def _run(mode, blobs, labels, step, rng, sample_size, smooth, bruteforce, postprocessing=False, export=None):
    ctx = get_context()
    out, logs, _ = nat.process_cases(ctx, mode, blobs, step, rng, sample_size, smooth, bruteforce, postprocessing)
    L = labels
    lg = tuple(_logs(l) for l in logs)
    if mode == 4:
        pairs = (_pair(out, 0, L[0], L[1]), _pair(out, 2, L[2], L[3]), _pair(out, 4, L[0], L[2]),
                 _pair(out, 6, L[1], L[3]))
    elif mode == 3:
        pairs = (_pair(out, 0, L[0], L[1]), _pair(out, 2, L[2], L[3]))
    elif mode == 2:
        pairs = (_pair(out, 0, L[0], L[1]),)
    else:
        geom = PyGeometry.from_blob(out[0], L[0])
        if export is not None:
            export.single(geom, out[0])
        return (geom, lg[0])
    if export is not None:
        export.pairs(pairs, out)
    return (*pairs, lg)


def _four_from_paths(path_ab, path_cd, labels, image_center, radius, n_points):
    # prepare_n_geometries(Full), preprocessing.rs:174-199: (a dia, a sys, b dia, b sys)
    use = labels is not None and len(labels) == 4
    jobs, names = [], []
    k = 0
    for p in (path_ab, path_cd):
        for dia in (True, False):
            name = labels[k] if use else _basename(p)
            jobs.append((p, name, dia))
            names.append(name)
            k += 1
    blobs = _parallel(lambda j: nat.geometry_from_dir(j[0], j[1], j[2], image_center, radius, n_points), jobs)
    return blobs, names


def from_file_full(input_path_ab, input_path_cd, labels=None, step_rotation_deg=0.5, range_rotation_deg=90.0,
                   sample_size=500, image_center=(4.5, 4.5), radius=0.5, n_points=20, write_obj=True,
                   watertight=True, contour_types=None, output_path_ab="output/rest",
                   output_path_cd="output/stress", output_path_ac="output/diastole",
                   output_path_bd="output/systole", interpolation_steps=0, bruteforce=False, smooth=True,
                   postprocessing=True):
    """functions.rs:143-245 -> (pair_ab, pair_cd, pair_ac, pair_bd, (logs_a, logs_b, logs_c, logs_d))."""
    export = _Export(write_obj, watertight, contour_types, interpolation_steps,
                     [output_path_ab, output_path_cd, output_path_ac, output_path_bd])
    blobs, names = _four_from_paths(input_path_ab, input_path_cd, labels, image_center, radius, n_points)
    return _run(4, blobs, names, step_rotation_deg, range_rotation_deg, sample_size, smooth, bruteforce, postprocessing,
                export)


def from_file_doublepair(input_path_ab, input_path_cd, labels=None, step_rotation_deg=0.5, range_rotation_deg=90.0,
                         sample_size=500, image_center=(4.5, 4.5), radius=0.5, n_points=20, write_obj=True,
                         watertight=True, contour_types=None, output_path_ab="output/rest",
                         output_path_cd="output/stress", interpolation_steps=0, bruteforce=False, smooth=True,
                         postprocessing=True):
    """functions.rs:332-413 -> (pair_ab, pair_cd, (logs x4))."""
    export = _Export(write_obj, watertight, contour_types, interpolation_steps, [output_path_ab, output_path_cd])
    blobs, names = _four_from_paths(input_path_ab, input_path_cd, labels, image_center, radius, n_points)
    return _run(3, blobs, names, step_rotation_deg, range_rotation_deg, sample_size, smooth, bruteforce, postprocessing,
                export)


def from_file_singlepair(input_path, labels=None, step_rotation_deg=0.5, range_rotation_deg=90.0, sample_size=500,
                         image_center=(4.5, 4.5), radius=0.5, n_points=20, write_obj=True, watertight=True,
                         contour_types=None, output_path="output/singlepair", interpolation_steps=0,
                         bruteforce=False, smooth=True, postprocessing=True):
    """functions.rs:498-564 -> (pair, (logs_a, logs_b))."""
    export = _Export(write_obj, watertight, contour_types, interpolation_steps, [output_path])
    use = labels is not None and len(labels) == 2
    names = [labels[i] if use else _basename(input_path) for i in range(2)]
    blobs = [nat.geometry_from_dir(input_path, names[i], dia, image_center, radius, n_points)
             for i, dia in enumerate((True, False))]
    return _run(2, blobs, names, step_rotation_deg, range_rotation_deg, sample_size, smooth, bruteforce, postprocessing,
                export)


def from_file_single(input_path, labels=None, diastole=True, step_rotation_deg=0.5, range_rotation_deg=90.0,
                     sample_size=500, image_center=(4.5, 4.5), radius=0.5, n_points=20, write_obj=True,
                     watertight=True, contour_types=None, output_path="output/single", bruteforce=False, smooth=True):
    """functions.rs:638-700 -> (geometry, logs)."""
    export = _Export(write_obj, watertight, contour_types, 0, [output_path])
    name = labels[0] if labels is not None and len(labels) == 1 else _basename(input_path)
    blob = nat.geometry_from_dir(input_path, name, diastole, image_center, radius, n_points)
    return _run(1, [blob], [name], step_rotation_deg, range_rotation_deg, sample_size, smooth, bruteforce,
                export=export)


_EXECUTOR = None


def _parallel(fn, args):
    """The pullbacks of a case are ingested concurrently (the C side releases the GIL), like the reference's
    crossbeam scope over its 4 geometries (binding/entry.rs:140-203). One executor for the life of the process: creating
    and joining four threads per call costs as much as ingesting a small pullback."""
    if len(args) <= 1:
        return [fn(a) for a in args]
    global _EXECUTOR
    if _EXECUTOR is None:
        from concurrent.futures import ThreadPoolExecutor
        _EXECUTOR = ThreadPoolExecutor(max_workers=8, thread_name_prefix="mmrs-ingest")
    return list(_EXECUTOR.map(fn, args))


def _from_inputs(mode, inputs, step, rng, sample_size, image_center, radius, n_points, smooth, bruteforce,
                 postprocessing=False, export=None):
    blobs = _parallel(lambda i: _blob_from_input(i, image_center, radius, n_points), list(inputs))
    return _run(mode, blobs, [i.label for i in inputs], step, rng, sample_size, smooth, bruteforce, postprocessing,
                export)


def from_array_full(input_data_a, input_data_b, input_data_c, input_data_d, step_rotation_deg=0.5,
                    range_rotation_deg=90.0, sample_size=500, image_center=(4.5, 4.5), radius=0.5, n_points=20,
                    write_obj=True, watertight=True, contour_types=None, output_path_ab="output/rest",
                    output_path_cd="output/stress", output_path_ac="output/diastole",
                    output_path_bd="output/systole", interpolation_steps=0, bruteforce=False, smooth=True,
                    postprocessing=True):
    """functions.rs:801-1010."""
    export = _Export(write_obj, watertight, contour_types, interpolation_steps,
                     [output_path_ab, output_path_cd, output_path_ac, output_path_bd])
    return _from_inputs(4, [input_data_a, input_data_b, input_data_c, input_data_d], step_rotation_deg,
                        range_rotation_deg, sample_size, image_center, radius, n_points, smooth, bruteforce,
                        postprocessing, export)


def from_array_doublepair(input_data_a, input_data_b, input_data_c, input_data_d, step_rotation_deg=0.5,
                          range_rotation_deg=90.0, sample_size=500, image_center=(4.5, 4.5), radius=0.5, n_points=20,
                          write_obj=True, watertight=True, contour_types=None, output_path_ab="output/rest",
                          output_path_cd="output/stress", interpolation_steps=0, bruteforce=False, smooth=True,
                          postprocessing=True):
    """functions.rs:1012-1187."""
    export = _Export(write_obj, watertight, contour_types, interpolation_steps, [output_path_ab, output_path_cd])
    return _from_inputs(3, [input_data_a, input_data_b, input_data_c, input_data_d], step_rotation_deg,
                        range_rotation_deg, sample_size, image_center, radius, n_points, smooth, bruteforce,
                        postprocessing, export)


def from_array_singlepair(input_data_a, input_data_b, step_rotation_deg=0.5, range_rotation_deg=90.0,
                          sample_size=500, image_center=(4.5, 4.5), radius=0.5, n_points=20, write_obj=True,
                          watertight=True, contour_types=None, output_path="output/singlepair",
                          interpolation_steps=0, bruteforce=False, smooth=True, postprocessing=True):
    """functions.rs:1189-1332."""
    export = _Export(write_obj, watertight, contour_types, interpolation_steps, [output_path])
    return _from_inputs(2, [input_data_a, input_data_b], step_rotation_deg, range_rotation_deg, sample_size,
                        image_center, radius, n_points, smooth, bruteforce, postprocessing, export)


def from_array_single(input_data, step_rotation_deg=0.5, range_rotation_deg=90.0, sample_size=500,
                      image_center=(4.5, 4.5), radius=0.5, n_points=20, write_obj=False, watertight=True,
                      contour_types=None, output_path="output/single", bruteforce=False, smooth=True):
    """functions.rs:1334-1433."""
    export = _Export(write_obj, watertight, contour_types, 0, [output_path])
    return _from_inputs(1, [input_data], step_rotation_deg, range_rotation_deg, sample_size, image_center, radius,
                        n_points, smooth, bruteforce, export=export)


def _align_centerline(method, centerline, geometry, write, watertight, interpolation_steps, output_dir,
                      contour_types, case_name, **kw):
    """Shared tail of align_three_point / align_manual / align_combined (binding/align.rs:65-463):
    PyGeometryPair or PyGeometry in, the same type out, (result, spacing_mm, total_rotation_deg)."""
    if not isinstance(centerline, PyCenterline):
        raise TypeError("centerline must be a PyCenterline")
    if isinstance(geometry, PyGeometryPair):
        geoms = [geometry.geom_a, geometry.geom_b]
    elif isinstance(geometry, PyGeometry):
        geoms = [geometry]
    else:
        raise TypeError("geometry must be a PyGeometry or PyGeometryPair")
    ctx = get_context() if method == 2 else None
    outs, spacing, rot, _ = nat.align_centerline(ctx, method, centerline._rows(), [g.to_blob() for g in geoms], **kw)
    res = [PyGeometry.from_blob(b, g.label) for b, g in zip(outs, geoms)]
    kinds = _kind_ids(contour_types)
    if len(res) == 2:
        out = PyGeometryPair(res[0], res[1], geometry.label)
        if write:  # Processable for GeometryPair, centerline_align/align.rs:26-43
            nat.export_pair(outs[0], outs[1], res[0].label, case_name, output_dir, interpolation_steps, watertight, kinds)
    else:
        out = res[0]
        if write:  # Processable for Geometry, align.rs:45-56 -> to_object::write_single_geometry
            nat.export_single(outs[0], case_name, output_dir, watertight, kinds, 1)
    return out, spacing, rot * (180.0 / 3.141592653589793)


def align_three_point(centerline, geometry, main_ref_pt, counterclockwise_ref_pt, clockwise_ref_pt, angle_step_deg=1.0,
                      write=False, watertight=True, interpolation_steps=0, output_dir="output/aligned",
                      contour_types=None, case_name="None", align_wall_anomalous=False):
    """binding/align.rs:65-155 -> (PyGeometry | PyGeometryPair, spacing_mm, total_rotation_deg). The three-point
    search (align_algorithms.rs:264-337) is a 3-point squared-error cost, host f64 in libmmrs_b200.so."""
    return _align_centerline(0, centerline, geometry, write, watertight, interpolation_steps, output_dir, contour_types,
                             case_name, main_ref_pt=main_ref_pt, ccw_ref_pt=counterclockwise_ref_pt,
                             cw_ref_pt=clockwise_ref_pt, angle_step_rad=angle_step_deg * (3.141592653589793 / 180.0),
                             align_wall_anomalous=align_wall_anomalous)


def align_manual(centerline, geometry, rotation_angle_deg, ref_point, write=False, watertight=True,
                 interpolation_steps=0, output_dir="output/aligned", contour_types=None, case_name="None",
                 align_wall_anomalous=False):
    """binding/align.rs:199-284."""
    return _align_centerline(1, centerline, geometry, write, watertight, interpolation_steps, output_dir, contour_types,
                             case_name, main_ref_pt=ref_point, manual_rotation_deg=rotation_angle_deg,
                             align_wall_anomalous=align_wall_anomalous)


def align_combined(centerline, geometry, main_ref_pt, counterclockwise_ref_pt, clockwise_ref_pt, points,
                   angle_step_deg=1.0, angle_range_deg=15.0, index_range=2, write=False, watertight=True,
                   interpolation_steps=0, output_dir="output/aligned", contour_types=None, case_name="None",
                   align_wall_anomalous=False):
    """binding/align.rs:334-463: three-point start + refine_alignment_hausdorff
    (align_algorithms.rs:339-451), every candidate Hausdorff-scored on the GPU in one batch."""
    rad = 3.141592653589793 / 180.0
    return _align_centerline(2, centerline, geometry, write, watertight, interpolation_steps, output_dir, contour_types,
                             case_name, main_ref_pt=main_ref_pt, ccw_ref_pt=counterclockwise_ref_pt,
                             cw_ref_pt=clockwise_ref_pt, angle_step_rad=angle_step_deg * rad, points=points,
                             angle_range_rad=angle_range_deg * rad, index_range=index_range,
                             align_wall_anomalous=align_wall_anomalous)


def to_obj(geometry, output_path, watertight=True, contour_types=None, filename_prefix=""):
    """binding/functions.rs:1435-1501: one OBJ (no UV coordinates) + MTL per requested contour type,
    "{prefix}_{type}.obj" or "{type}.obj"; contour types a geometry does not carry are skipped with a warning."""
    if not isinstance(geometry, PyGeometry):
        raise TypeError("geometry must be a PyGeometry")
    nat.export_single(geometry.to_blob(), str(filename_prefix), output_path, watertight, _kind_ids(contour_types), 2)
