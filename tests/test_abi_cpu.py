"""CPU-side checks of the C-ABI library: it loads, exports every symbol the
header declares, fails loudly without a GPU, and its host-side grid / stage
arithmetic matches the oracle (process_utils.rs:43-67, align_within.rs:208-246)."""
import ctypes as C
import math
import re
from pathlib import Path

import numpy as np
import pytest

from multimodars import _native as nat
from oracle import oracle_py as ora

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    hdr = (ROOT / "include" / "mmrs_b200.h").read_text()
    declared = set(re.findall(r"\b(mmrs_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(nat.EXPORTS), declared ^ set(nat.EXPORTS)
    L = nat.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.mmrs_version()


def test_struct_layouts_match_header():
    assert C.sizeof(nat.Grid) == 40
    assert C.sizeof(nat.UnitResult) == 40
    assert nat.RESULT_DTYPE.itemsize == 40


def test_compute_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nat.MmrsError, match="no CPU fallback|no CUDA device"):
        nat.Context(0)


GRID_CASES = [
    (1.0, 90.0, None, 90.0), (0.5, 90.0, None, 90.0), (0.01, 180.0, None, 180.0), (0.005, 180.0, None, 180.0),
    (0.05, 90.0, None, 90.0), (0.01, 6.0, None, 6.0), (0.1, 5.0, 0.3, 90.0), (0.1, 5.0, -1.55, 90.0),
    (0.01, 0.1, 1.2345, 180.0), (0.01, 0.1, math.pi - 1e-3, 180.0), (0.001, 0.01, -0.3, 20.0),
    (1.0, 180.0, None, 90.0), (0.5, 45.0, 0.8, 180.0), (0.1, 0.2, 0.0, 180.0), (3.0, 10.0, None, 10.0),
    (7.0, 20.0, 0.1, 20.0), (1.0, 0.5, None, 0.5), (0.3, 1.0, 3.1, 180.0),
]


@pytest.mark.parametrize("step,rng,center,limes", GRID_CASES)
def test_grid_matches_oracle(step, rng, center, limes):
    want, fb = ora.search_grid(step, rng, center, limes)
    g = nat.make_grid(step, rng, center, limes)
    assert not g.degenerate
    assert g.n_cand == len(want)
    got = np.array([nat.grid_angle(g, i) for i in range(0, g.n_cand, max(1, g.n_cand // 997))])
    assert (got == want[::max(1, g.n_cand // 997)]).all()      # bit-exact
    assert nat.grid_angle(g, g.n_cand - 1) == want[-1]


@pytest.mark.parametrize("step,rng,center,limes", [(0.0, 90.0, 1.0, 180.0), (-1.0, 90.0, 0.5, 180.0),
                                                   (0.0, 90.0, None, 180.0), (1.0, 0.0, 0.25, 90.0),
                                                   (1.0, 10.0, 3.0, 5.0), (float("nan"), 10.0, 0.2, 10.0)])
def test_degenerate_grids(step, rng, center, limes):
    want, fb = ora.search_grid(step, rng, center, limes)
    assert want is None or len(want) == 0   # early return, or an empty take_while -> unwrap_or(center)
    g = nat.make_grid(step, rng, center, limes)
    assert g.degenerate
    assert g.fallback == fb


@pytest.mark.parametrize("step,rng", [(1.0, 90.0), (2.5, 30.0), (0.5, 90.0), (0.1, 90.0), (0.1, 3.0), (0.05, 90.0),
                                      (0.01, 180.0), (0.01, 0.05), (0.005, 180.0), (0.001, 20.0), (0.0, 20.0),
                                      (-1.0, 20.0), (0.999999, 30.0), (0.0999999, 30.0)])
def test_stage_plan_matches_reference_arms(step, rng):
    plan = nat.stage_plan(step, rng)
    r5 = 5.0 if rng > 5.0 else rng
    r10 = 10.0 * step if rng > 10.0 * step else rng
    if step >= 1.0:
        want = [(step, rng)]
    elif 0.1 <= step < 1.0:
        want = [(1.0, rng), (step, r5)]
    elif 0.01 <= step < 0.1:
        want = [(1.0, rng), (0.1, r5), (step, r10)]
    else:
        want = [(1.0, rng), (0.1, r5), (0.01, 0.1 if rng > 0.1 else rng), (step, r10)]
    assert plan == want


def test_integration_md_binds_every_declared_symbol():
    """INTEGRATION.md's Rust `extern "C"` block (what a maintainer pastes into the -sys crate) names every function
    include/mmrs_b200.h declares, and its mmrs_sweep_opts carries every field of the C struct in order."""
    import re

    root = Path(__file__).resolve().parent.parent
    header = (root / "include" / "mmrs_b200.h").read_text()
    doc = (root / "INTEGRATION.md").read_text()
    declared = set(re.findall(r"^(?:int|void|double|const char\*)\s+(mmrs_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert len(declared) >= 30
    bound = set(re.findall(r"pub fn (mmrs_[a-z0-9_]+)\s*\(", doc))
    assert declared <= bound, sorted(declared - bound)
    end = header.index("} mmrs_sweep_opts;")
    opts_c = header[header.rindex("typedef struct", 0, end):end]
    opts_c = re.sub(r"/\*.*?\*/", "", opts_c, flags=re.S)
    fields_c = re.findall(r"(?:double|int32_t|int64_t|float)\s+([a-z0-9_]+);", opts_c)
    opts_rs = re.search(r"pub struct mmrs_sweep_opts \{(.*?)\}", doc, flags=re.S).group(1)
    fields_rs = re.findall(r"pub ([a-z0-9_]+):", opts_rs)
    assert fields_c == fields_rs, (fields_c, fields_rs)
