"""Drop-in check at the Python boundary: the REFERENCE's own pytest files for this part of its surface
(tests/test_core.py, test_converters.py, test_intravascular.py, test_wrappers.py) are run, unmodified and from where
they lie, with `multimodars` resolving to THIS package. Only possible where /root/reference exists (this container);
elsewhere the test is skipped — nothing here runs on the GPU box. The CCTA test file (test_ccta.py: meshes, labeling)
is outside this build."""
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest

REF = Path("/root/reference")
PKG = Path(__file__).resolve().parent.parent / "multimoda-rs_b200"
FILES = ["tests/test_core.py", "tests/test_converters.py", "tests/test_intravascular.py", "tests/test_wrappers.py"]


@pytest.mark.skipif(not (REF / "tests" / "test_core.py").exists(), reason="reference checkout not present")
def test_reference_python_tests_pass_against_this_package(tmp_path):
    # the tests open "data/..." and "examples/..." relative to the cwd; running inside /root/reference itself would
    # put the reference's (unbuilt) `multimodars` directory first on sys.path, so use a scratch dir of symlinks
    for name in ("data", "examples", "tests"):
        (tmp_path / name).symlink_to(REF / name)
    env = dict(os.environ, PYTHONPATH=str(PKG), PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", *FILES, "-q", "--no-header", "-p", "no:cacheprovider", "-rs"],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    tail = r.stdout[-3000:] + r.stderr[-1000:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 30, tail
    # the skips are the reference's own (unconditional `pytest.skip` placeholders and test_wrappers.py passing keyword
    # arguments its own wrappers do not have — SURVEY.md §4), not missing features here
    skipped = re.findall(r"SKIPPED \[\d+\] (\S+?):\d+: (.*)", r.stdout)
    for where, why in skipped:
        assert ("needs more complex setup" in why or "unexpected keyword argument 'label'" in why), (where, why)
