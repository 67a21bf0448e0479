"""sincos_mid (b2g_math.h), the branch-free sine/cosine of the straight-line position kernel, against
sincos_ref (the restated glibc sincosf the whole engine uses): bitwise equal on every float of its domain
|y| < 120.  The comparison is a C++ program (tools/sincos_mid_check.cpp) over all 2.2e9 values."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_sincos_mid_matches_sincos_ref_exhaustively(tmp_path):
    exe = str(tmp_path / "sincos_mid_check")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-mfma", "-std=c++17", "-pthread",
                    os.path.join(ROOT, "tools", "sincos_mid_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0, out.stdout[-500:]
    assert " 0 mismatches" in out.stdout, out.stdout[-500:]
    checked = int(out.stdout.split("checked")[1].split()[0])
    assert checked > 2_000_000_000
