"""The reference's public processing entry points (multimodars/_processing.py:42-1009 over
src/intravascular/binding/functions.rs:143-1433), same names, argument order, defaults and
return shapes, running on the B200 through libmmrs_b200.so (mmrs_process_cases).

Out of scope in this build: `write_obj=True` (OBJ/MTL/texture export) raises
NotImplementedError instead of silently doing something else; pass write_obj=False, as the
reference's own benchmarks do (benchmarks/benchmark_bruteforce_stepsize.py:30-57)."""
from __future__ import annotations

import os

import numpy as np

from . import _native as nat
from ._types import PyContourType, PyGeometry, PyGeometryPair, PyInputData

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


def _unsupported(write_obj, postprocessing=False):
    if write_obj:
        raise NotImplementedError("write_obj=True (OBJ/MTL export, to_object/process.rs:13) is outside this "
                                  "build's scope; pass write_obj=False")


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


def _run(mode, blobs, labels, step, rng, sample_size, smooth, bruteforce, postprocessing=False):
    ctx = get_context()
    out, logs, _ = nat.process_cases(ctx, mode, blobs, step, rng, sample_size, smooth, bruteforce, postprocessing)
    L = labels
    lg = tuple(_logs(l) for l in logs)
    if mode == 4:
        return (_pair(out, 0, L[0], L[1]), _pair(out, 2, L[2], L[3]), _pair(out, 4, L[0], L[2]),
                _pair(out, 6, L[1], L[3]), lg)
    if mode == 3:
        return (_pair(out, 0, L[0], L[1]), _pair(out, 2, L[2], L[3]), lg)
    if mode == 2:
        return (_pair(out, 0, L[0], L[1]), lg)
    return (PyGeometry.from_blob(out[0], L[0]), lg[0])


def _four_from_paths(path_ab, path_cd, labels, image_center, radius, n_points):
    # prepare_n_geometries(Full), preprocessing.rs:174-199: (a dia, a sys, b dia, b sys)
    use = labels is not None and len(labels) == 4
    blobs, names = [], []
    k = 0
    for p in (path_ab, path_cd):
        for dia in (True, False):
            name = labels[k] if use else _basename(p)
            blobs.append(nat.geometry_from_dir(p, name, dia, image_center, radius, n_points))
            names.append(name)
            k += 1
    return blobs, names


def from_file_full(input_path_ab, input_path_cd, labels=None, step_rotation_deg=0.5, range_rotation_deg=90.0,
                   sample_size=500, image_center=(4.5, 4.5), radius=0.5, n_points=20, write_obj=True,
                   watertight=True, contour_types=None, output_path_ab="output/rest",
                   output_path_cd="output/stress", output_path_ac="output/diastole",
                   output_path_bd="output/systole", interpolation_steps=0, bruteforce=False, smooth=True,
                   postprocessing=True):
    """functions.rs:143-245 -> (pair_ab, pair_cd, pair_ac, pair_bd, (logs_a, logs_b, logs_c, logs_d))."""
    _unsupported(write_obj, postprocessing)
    blobs, names = _four_from_paths(input_path_ab, input_path_cd, labels, image_center, radius, n_points)
    return _run(4, blobs, names, step_rotation_deg, range_rotation_deg, sample_size, smooth, bruteforce, postprocessing)


def from_file_doublepair(input_path_ab, input_path_cd, labels=None, step_rotation_deg=0.5, range_rotation_deg=90.0,
                         sample_size=500, image_center=(4.5, 4.5), radius=0.5, n_points=20, write_obj=True,
                         watertight=True, contour_types=None, output_path_ab="output/rest",
                         output_path_cd="output/stress", interpolation_steps=0, bruteforce=False, smooth=True,
                         postprocessing=True):
    """functions.rs:332-413 -> (pair_ab, pair_cd, (logs x4))."""
    _unsupported(write_obj, postprocessing)
    blobs, names = _four_from_paths(input_path_ab, input_path_cd, labels, image_center, radius, n_points)
    return _run(3, blobs, names, step_rotation_deg, range_rotation_deg, sample_size, smooth, bruteforce, postprocessing)


def from_file_singlepair(input_path, labels=None, step_rotation_deg=0.5, range_rotation_deg=90.0, sample_size=500,
                         image_center=(4.5, 4.5), radius=0.5, n_points=20, write_obj=True, watertight=True,
                         contour_types=None, output_path="output/singlepair", interpolation_steps=0,
                         bruteforce=False, smooth=True, postprocessing=True):
    """functions.rs:498-564 -> (pair, (logs_a, logs_b))."""
    _unsupported(write_obj, postprocessing)
    use = labels is not None and len(labels) == 2
    names = [labels[i] if use else _basename(input_path) for i in range(2)]
    blobs = [nat.geometry_from_dir(input_path, names[i], dia, image_center, radius, n_points)
             for i, dia in enumerate((True, False))]
    return _run(2, blobs, names, step_rotation_deg, range_rotation_deg, sample_size, smooth, bruteforce, postprocessing)


def from_file_single(input_path, labels=None, diastole=True, step_rotation_deg=0.5, range_rotation_deg=90.0,
                     sample_size=500, image_center=(4.5, 4.5), radius=0.5, n_points=20, write_obj=True,
                     watertight=True, contour_types=None, output_path="output/single", bruteforce=False, smooth=True):
    """functions.rs:638-700 -> (geometry, logs)."""
    _unsupported(write_obj, False)
    name = labels[0] if labels is not None and len(labels) == 1 else _basename(input_path)
    blob = nat.geometry_from_dir(input_path, name, diastole, image_center, radius, n_points)
    return _run(1, [blob], [name], step_rotation_deg, range_rotation_deg, sample_size, smooth, bruteforce)


def _from_inputs(mode, inputs, step, rng, sample_size, image_center, radius, n_points, smooth, bruteforce,
                 postprocessing=False):
    blobs = [_blob_from_input(i, image_center, radius, n_points) for i in inputs]
    return _run(mode, blobs, [i.label for i in inputs], step, rng, sample_size, smooth, bruteforce, postprocessing)


def from_array_full(input_data_a, input_data_b, input_data_c, input_data_d, step_rotation_deg=0.5,
                    range_rotation_deg=90.0, sample_size=500, image_center=(4.5, 4.5), radius=0.5, n_points=20,
                    write_obj=True, watertight=True, contour_types=None, output_path_ab="output/rest",
                    output_path_cd="output/stress", output_path_ac="output/diastole",
                    output_path_bd="output/systole", interpolation_steps=0, bruteforce=False, smooth=True,
                    postprocessing=True):
    """functions.rs:801-1010."""
    _unsupported(write_obj, postprocessing)
    return _from_inputs(4, [input_data_a, input_data_b, input_data_c, input_data_d], step_rotation_deg,
                        range_rotation_deg, sample_size, image_center, radius, n_points, smooth, bruteforce,
                        postprocessing)


def from_array_doublepair(input_data_a, input_data_b, input_data_c, input_data_d, step_rotation_deg=0.5,
                          range_rotation_deg=90.0, sample_size=500, image_center=(4.5, 4.5), radius=0.5, n_points=20,
                          write_obj=True, watertight=True, contour_types=None, output_path_ab="output/rest",
                          output_path_cd="output/stress", interpolation_steps=0, bruteforce=False, smooth=True,
                          postprocessing=True):
    """functions.rs:1012-1187."""
    _unsupported(write_obj, postprocessing)
    return _from_inputs(3, [input_data_a, input_data_b, input_data_c, input_data_d], step_rotation_deg,
                        range_rotation_deg, sample_size, image_center, radius, n_points, smooth, bruteforce,
                        postprocessing)


def from_array_singlepair(input_data_a, input_data_b, step_rotation_deg=0.5, range_rotation_deg=90.0,
                          sample_size=500, image_center=(4.5, 4.5), radius=0.5, n_points=20, write_obj=True,
                          watertight=True, contour_types=None, output_path="output/singlepair",
                          interpolation_steps=0, bruteforce=False, smooth=True, postprocessing=True):
    """functions.rs:1189-1332."""
    _unsupported(write_obj, postprocessing)
    return _from_inputs(2, [input_data_a, input_data_b], step_rotation_deg, range_rotation_deg, sample_size,
                        image_center, radius, n_points, smooth, bruteforce, postprocessing)


def from_array_single(input_data, step_rotation_deg=0.5, range_rotation_deg=90.0, sample_size=500,
                      image_center=(4.5, 4.5), radius=0.5, n_points=20, write_obj=False, watertight=True,
                      contour_types=None, output_path="output/single", bruteforce=False, smooth=True):
    """functions.rs:1334-1433."""
    _unsupported(write_obj, False)
    return _from_inputs(1, [input_data], step_rotation_deg, range_rotation_deg, sample_size, image_center, radius,
                        n_points, smooth, bruteforce)


def align_three_point(*args, **kwargs):
    """binding/align.rs:65-155. The 3-point centerline search is not Hausdorff-scored and is
    outside the hot path this build replaces."""
    raise NotImplementedError("align_three_point: centerline alignment is outside this build's scope (DESIGN.md §8)")
