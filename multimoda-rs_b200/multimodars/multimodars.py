"""`multimodars.multimodars` — in the reference this is the PyO3 extension module itself (src/lib.rs:25-102) and code
such as `from multimodars.multimodars import PyContour, from_file_full` imports from it directly. Here it is a thin
module with the same names: the value classes and the native-level entry points, whose defaults are the Rust
`#[pyo3(signature = ...)]` ones where those differ from the Python wrappers' (`sample_size = 200` for `from_array_*`
and `from_file_single`, functions.rs:645, :810, :1021, :1196, :1339). Everything runs through libmmrs_b200.so."""
from __future__ import annotations

import functools
import inspect

from . import _processing as _p
from ._types import (PyCenterline, PyCenterlinePoint, PyContour, PyContourPoint, PyContourType, PyFrame,  # noqa: F401
                     PyGeometry, PyGeometryPair, PyInputData, PyRecord)
from ._vtp import read_centerline_vtp as _read_vtp


def _native(fn, **defaults):
    """The wrapper `fn` with some keyword defaults replaced (signature included, so `inspect` shows the native one)."""
    sig = inspect.signature(fn)
    params = [p.replace(default=defaults[n]) if n in defaults else p for n, p in sig.parameters.items()]

    @functools.wraps(fn)
    def call(*a, **k):
        bound = sig.bind(*a, **k)
        for n, v in defaults.items():
            bound.arguments.setdefault(n, v)
        return fn(*bound.args, **bound.kwargs)

    call.__signature__ = sig.replace(parameters=params)
    return call


from_file_full = _p.from_file_full
from_file_doublepair = _p.from_file_doublepair
from_file_singlepair = _p.from_file_singlepair
from_file_single = _native(_p.from_file_single, sample_size=200)
from_array_full = _native(_p.from_array_full, sample_size=200)
from_array_doublepair = _native(_p.from_array_doublepair, sample_size=200)
from_array_singlepair = _native(_p.from_array_singlepair, sample_size=200)
from_array_single = _native(_p.from_array_single, sample_size=200)
align_three_point = _p.align_three_point
align_manual = _p.align_manual
align_combined = _p.align_combined
to_obj = _p.to_obj


def read_centerline_vtp(path):
    """functions.rs:1541-1546."""
    return _read_vtp(path)
