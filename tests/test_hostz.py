"""Host DEFLATE decoder and CRC-32 of the BGZF readers (csrc/hostz.cpp) against zlib, under the address and
undefined-behaviour sanitizers (tests/hostz_check.cpp).  Pure host code: runs without a GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hostz_against_zlib(tmp_path):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not found")
    exe = str(tmp_path / "hostz_check")
    csrc = os.path.join(ROOT, "cnvetti_b200", "csrc")
    cmd = [gxx, "-O2", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-I", csrc,
           os.path.join(csrc, "hostz.cpp"), os.path.join(ROOT, "tests", "hostz_check.cpp"), "-lz", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:   # (no sanitizer runtime on this box: plain build)
        cmd = [c for c in cmd if "sanitize" not in c]
        r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "300"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "crc ok" in r.stdout and "inflate ok" in r.stdout, r.stdout
