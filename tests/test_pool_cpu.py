"""csrc/mmrs_pool.hpp (the library's persistent host pool) under ThreadSanitizer: six concurrent submitters, nested jobs,
and the exception of the LOWEST failing index winning like a serial loop (tests/native/pool_stress.cpp)."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("sanitize", [True, False])
def test_host_pool_stress(tmp_path, sanitize):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = tmp_path / "pool_stress"
    cmd = [gxx, "-std=c++17", "-O1", "-g", f"-I{ROOT / 'multimoda-rs_b200' / 'csrc'}", str(ROOT / "tests" / "native" / "pool_stress.cpp"),
           "-o", str(exe), "-lpthread"]
    if sanitize:
        cmd.insert(1, "-fsanitize=thread")
    b = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    if b.returncode != 0 and sanitize:
        pytest.skip("ThreadSanitizer runtime not available: " + b.stderr[-200:])
    assert b.returncode == 0, b.stderr[-2000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-3000:]
    assert "WARNING: ThreadSanitizer" not in r.stderr
    assert "lowest-index exceptions 200 / 200" in r.stdout
