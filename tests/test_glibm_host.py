"""csrc/glibm.cuh (the device sin/cos/acos/atan/pow(x,2) of the danger-zone path) compiled for the host with g++ and
compared bit for bit with the system libm, i.e. with the arithmetic the reference's numpy/python calls perform
(satellite_function.py:161-255, :317-373, :462-565). CPU test: no GPU, no oracle involved."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ppo-rl-satellite_b200", "csrc")


def _glibc_version():
    import ctypes
    f = ctypes.CDLL(None).gnu_get_libc_version
    f.restype = ctypes.c_char_p
    return f().decode()


@pytest.fixture(scope="module")
def host_binary(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("glibm") / "glibm_host")
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-I", CSRC, os.path.join(ROOT, "tests", "glibm_host.cpp"), "-o", exe, "-lm"]
    subprocess.run(cmd, check=True)
    return exe


def test_glibm_bit_identical_to_system_libm(host_binary):
    if not _glibc_version().startswith("2.39"):
        pytest.skip("glibm.cuh restates glibc 2.39; this host runs %s" % _glibc_version())
    r = subprocess.run([host_binary, "400000", "11"], capture_output=True, text=True)
    lines = r.stdout.strip().splitlines()
    assert lines[-1] == "total_mismatches 0", r.stdout[-2000:]
    assert r.returncode == 0
    per_fn = {}
    for ln in lines[:-1]:
        name, dist, n, bad = ln.split()
        per_fn[name] = per_fn.get(name, 0) + int(n)
        assert int(bad) == 0
    assert set(per_fn) == {"sin", "cos", "sincos.s", "sincos.c", "acos", "atan", "pow2"}
    assert min(per_fn.values()) > 1_000_000


def test_tables_match_this_hosts_libm():
    if not os.path.exists("/lib/x86_64-linux-gnu/libm.so.6") or not _glibc_version().startswith("2.39"):
        pytest.skip("needs the x86-64 glibc 2.39 image")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "extract_libm_tables.py"), "--check"], capture_output=True, text=True)
    if r.returncode != 0 and "AssertionError" in r.stderr:
        pytest.skip("a different libm build (tables at other addresses)")
    assert r.returncode == 0, r.stdout + r.stderr
