"""Source compatibility of the Python surface with the reference's: every public function of multimodars/_processing.py
and _converters.py and every class / method / attribute of multimodars.pyi that belongs to this build must exist here
with the same parameter names, order and default values (tests/golden/reference_signatures.json, read from the
reference with `ast` by tests/golden/make_signature_golden.py). What is outside this build is listed explicitly."""
import ast
import inspect
import json
from pathlib import Path

import numpy as np
import pytest

import multimodars as mm
from multimodars import _converters, _processing

SIG = json.loads((Path(__file__).resolve().parent / "golden" / "reference_signatures.json").read_text())

# the CCTA / mesh side of the reference (src/ccta, multimodars/ccta, trimesh) — DESIGN.md §8 "out of scope"
OUTSIDE_FUNCTIONS = {"find_centerline_bounded_points_simple", "find_proximal_distal_scaling", "build_adjacency_map",
                     "discretize_vessel", "geometry_to_trimesh"}
OUTSIDE_CLASSES = {"PyDiscretizedVesselTree"}


def _ours(f):
    return [(n, None if p.default is inspect.Parameter.empty else p.default)
            for n, p in inspect.signature(f).parameters.items() if n not in ("self", "cls")]


def _same_default(stub, ours):
    if stub == "...":                      # the stub elides the value: any default will do
        return True
    try:
        want = ast.literal_eval(stub)
    except (ValueError, SyntaxError):
        return stub.replace('"', "'") == repr(ours).replace('"', "'")
    if isinstance(want, (list, tuple)) and not isinstance(ours, str) and ours is not None:
        return list(want) == [str(x) if not isinstance(x, (int, float)) else x for x in ours] or list(want) == list(ours)
    return want == ours


@pytest.mark.parametrize("name", sorted(SIG["functions"]))
def test_function_signature(name):
    ref = SIG["functions"][name]
    if name in OUTSIDE_FUNCTIONS:
        assert not hasattr(mm, name)       # no half-working stand-ins
        return
    mod = _processing if ref["module"] == "_processing" else _converters
    f = getattr(mod, name)
    assert getattr(mm, name, f) is f or name in ("array_to_pyinputdata", "geometry_to_frames_array")
    ours = _ours(f)
    assert [n for n, _ in ours] == [n for n, _ in ref["params"]]
    for (n, d), (_, s) in zip(ours, ref["params"]):
        if s is None:
            assert d is None, f"{name}({n}) is required in the reference"
        else:
            assert _same_default(s, d), f"{name}({n}): reference default {s}, here {d!r}"


@pytest.mark.parametrize("cls", sorted(SIG["classes"]))
def test_class_surface(cls):
    ref = SIG["classes"][cls]
    if cls in OUTSIDE_CLASSES:
        assert not hasattr(mm, cls)
        return
    c = getattr(mm, cls)
    for m, info in ref["methods"].items():
        assert hasattr(c, m), f"{cls}.{m} missing"
        if m.startswith("__") and m != "__init__":
            continue
        names = [n for n, _ in _ours(getattr(c, m))]
        want = [n for n, _ in info["params"]]
        # a constructor here may accept more trailing optional arguments than the stub shows, never fewer or renamed
        assert names[:len(want)] == want, f"{cls}.{m}: {names} vs {want}"
        for (n, d) in _ours(getattr(c, m))[len(want):]:
            assert d is not None or n in ("aortic_thickness", "pulmonary_thickness"), f"{cls}.{m}: extra required {n}"
        if info["static"]:
            assert isinstance(inspect.getattr_static(c, m), staticmethod), f"{cls}.{m} must be static"


def test_instance_attributes():
    p = mm.PyContourPoint(1, 2, 0.5, 1.5, 2.5, True)
    c = mm.PyContour(3, 3, [p, mm.PyContourPoint(1, 3, 1.0, 1.0, 2.5, False)], (0.75, 1.25, 2.5), None, None, "Lumen")
    f = mm.PyFrame(3, c.centroid, c, {}, p)
    g = mm.PyGeometry([f], "lbl")
    pair = mm.PyGeometryPair(g, g, "pair")
    clp = mm.PyCenterlinePoint(p, (0.0, 0.0, 1.0))
    cl = mm.PyCenterline([clp])
    rec = mm.PyRecord(1, "D", 1.0, None)
    inp = mm.PyInputData([c], None, None, None, [rec], p, True, "x")
    for obj in (p, c, f, g, pair, clp, cl, rec, inp, mm.PyContourType.Lumen):
        ref = SIG["classes"][type(obj).__name__]
        for a in ref["attributes"] + ref["properties"]:
            assert hasattr(obj, a), f"{type(obj).__name__}.{a}"
    assert repr(mm.PyContourType.Wall) == "PyContourType.Wall" and str(mm.PyContourType.Wall) == "Wall"
    assert mm.PyContourType.from_string("LUMEN") is mm.PyContourType.Lumen
    with pytest.raises(ValueError, match="Unknown contour type: 'x'. Valid types are: lumen, eem"):
        mm.PyContourType.from_string("x")
    assert np.isfinite(cl.points[0].radius)


def test_native_module_names_and_defaults():
    """`from multimodars.multimodars import ...` (the PyO3 module of the reference, src/lib.rs:25-102) resolves here, and
    its entry points carry the Rust-level defaults: sample_size = 200 for from_array_* and from_file_single
    (functions.rs:645, :810, :1021, :1196, :1339), write_obj = False only for from_array_single (:1343)."""
    from multimodars import multimodars as native
    for name in ("PyInputData", "PyContourPoint", "PyContour", "PyContourType", "PyFrame", "PyGeometry",
                 "PyGeometryPair", "PyCenterlinePoint", "PyCenterline", "PyRecord"):          # lib.rs:90-99
        assert getattr(native, name) is getattr(mm, name)
    want = {"from_file_full": (500, True), "from_file_doublepair": (500, True), "from_file_singlepair": (500, True),
            "from_file_single": (200, True), "from_array_full": (200, True), "from_array_doublepair": (200, True),
            "from_array_singlepair": (200, True), "from_array_single": (200, False)}
    for name, (sample, write) in want.items():
        p = inspect.signature(getattr(native, name)).parameters
        assert (p["sample_size"].default, p["write_obj"].default) == (sample, write), name
        assert list(p) == list(inspect.signature(getattr(mm, name)).parameters), name
    for name in ("align_three_point", "align_manual", "align_combined", "to_obj"):
        assert getattr(native, name) is getattr(mm, name)
    assert list(inspect.signature(native.read_centerline_vtp).parameters) == ["path"]     # functions.rs:1541
    # the replaced default really reaches the call: the wrapper raises its own TypeError for a non-PyInputData
    # argument only after binding, so a bad keyword is still refused with the native parameter list
    with pytest.raises(TypeError):
        native.from_array_single("not input data", sample=3)
    import multimodars._processing as proc
    assert native.from_array_single.__wrapped__ is proc.from_array_single
