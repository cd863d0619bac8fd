"""The C++ host mirror (lp_b200/host/ripped.hpp) against the reference's own unit tests, restated in
tests/cpp/test_ripped.cpp.  CPU run: builder / slack-form / error cases + "fails loudly without a GPU";
GPU run (-m gpu): the known-answer problems through the host-driven phase calls of the C ABI."""
import os
import re
import subprocess

import pytest

from lp_b200 import _ffi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "cpp", "test_ripped.cpp")
OUT = os.path.join(HERE, "cpp", "_build", "test_ripped")


def build_exe():
    if not os.path.exists(_ffi.LIB_PATH):
        from lp_b200 import build
        build.build(verbose=False)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "lp_b200", "host", "ripped.hpp"), os.path.join(ROOT, "lp_b200", "csrc", "ipm_driver.hpp"),
            os.path.join(ROOT, "include", "lpb200.h"), _ffi.LIB_PATH]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        libdir = os.path.dirname(_ffi.LIB_PATH)
        subprocess.check_call(["g++", "-O2", "-std=c++17", SRC, "-o", OUT, "-L" + libdir, "-llpb200",
                               "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"])
    return OUT


def test_cpp_host_mirror_cpu_cases():
    exe = build_exe()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_host_mirror_reference_known_answers_on_gpu():
    exe = build_exe()
    r = subprocess.run([exe, "--gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"symmetric: fun=(\S+) iterations=(\d+)", r.stdout)
    assert m and abs(float(m.group(1)) + 1000.0) < 1e-6 and int(m.group(2)) == 4


def test_distributed_cholesky_schedule_and_tile_lists():
    """lp_b200/csrc/dist_schedule.hpp on the CPU for world sizes 2, 3, 4, 8 (tests/cpp/test_dist_schedule.cpp):
    the owned-column tile decode and the per-rank operation order of the panel-broadcast factorisation -- and, for the
    peer-memory hand-off of the first panel rows (one-sided writes into the peers' slot rings, eager large broadcast),
    1280 random interleavings of ranks and streams per world size: no stale read with a ring of 2 x world slots, and a
    two-slot control that must (and does) produce them."""
    src = os.path.join(HERE, "cpp", "test_dist_schedule.cpp")
    out = os.path.join(HERE, "cpp", "_build", "test_dist_schedule")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", src, "-o", out])
    r = subprocess.run([out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "dist_schedule ok" in r.stdout, r.stdout + r.stderr
