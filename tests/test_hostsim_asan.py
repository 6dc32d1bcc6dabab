"""Memory safety of the kernel bodies: the host simulation is rebuilt with -fsanitize=address and a driver runs every
kernel body on small / ragged problems (compute-sanitizer is closed on the GPU pool)."""
import os
import subprocess
import sys

import pytest

from tests.conftest import ROOT


def test_kernel_bodies_under_address_sanitizer(tmp_path):
    libasan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(libasan) or not os.path.exists(libasan):
        pytest.skip("libasan not available")
    so = str(tmp_path / "libqspush_hostsim_asan.so")
    csrc = os.path.join(ROOT, "uclv_qs_pushing_matlab_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-Wno-unknown-pragmas", "-fno-fast-math", "-ffp-contract=off",
                           "-fsanitize=address", "-fno-omit-frame-pointer", "-shared", "-x", "c++", "-o", so,
                           os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp"), os.path.join(csrc, "qs_model.cpp")])
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "asan_driver.py"), so], capture_output=True, text=True, env=env, timeout=900)
    assert "ASAN-DRIVER-OK" in r.stdout and "AddressSanitizer" not in r.stderr, r.stderr[-3000:]
