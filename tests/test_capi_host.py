"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/qspush.h declares, the host-only
model ingest works, and every compute entry point fails loudly without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import uclv_qs_pushing_matlab_b200 as q
from uclv_qs_pushing_matlab_b200 import _lib as L
from tests.conftest import REFERENCE_CAD, ROOT
from tests.workloads import OBJECT_ORDER, OBJECT_TABLE, gpu_model, oracle_model, packaged_tables

HAVE_GPU = L.lib().qspush_device_count() > 0


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "qspush.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(qspush_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(L.SIGNATURES), (declared ^ set(L.SIGNATURES))
    lib = C.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert L.lib().qspush_version().decode().startswith("qspush-b200")


def test_no_torch_types_in_the_abi():
    hdr = open(os.path.join(ROOT, "include", "qspush.h")).read()
    assert "torch" not in hdr and "at::" not in hdr and "#include <cuda" not in hdr


@pytest.mark.parametrize("name", OBJECT_ORDER)
def test_model_tables_match_oracle(name):
    gm, om = gpu_model(name), oracle_model(name)
    assert (gm.n, gm.nknots, gm.b, gm.c_ellipse, gm.mu_sp) == (om.n, om.nknots, om.b, om.c_ellipse, om.mu_sp)
    assert np.array_equal(gm.S, om.S) and np.array_equal(gm.P, om.P)
    assert np.array_equal(gm.cj_1_vect, om.c1) and np.array_equal(gm.cj_2_vect, om.c2)


@pytest.mark.skipif(not os.path.isdir(REFERENCE_CAD), reason="reference .ply files only exist in the build container")
@pytest.mark.parametrize("name", OBJECT_ORDER)
def test_ply_ingest_product_side(name):
    o = OBJECT_TABLE[name]
    m = q.Model.from_ply(os.path.join(REFERENCE_CAD, o["pcl_path"]), name in ("montana", "pulirapid"), 3, o["mu_sg"], o["mu_sp"], o["m"], o["tau_max"])
    t = packaged_tables()[name]
    assert np.array_equal(m.S, np.array(t["knots"])) and np.array_equal(m.P, np.array(t["ctrl_xy"])) and m.c_ellipse == t["c_ellipse"]


def test_ply_ingest_synthetic_file(tmp_path):
    """A synthetic binary_little_endian outline (decagon in mm) goes through sortCadPoints + getSpline."""
    ang = np.linspace(0, 2 * np.pi, 11)[:-1]
    pts = np.stack([40 * np.cos(ang), 30 * np.sin(ang)], 1).astype(np.float32)
    perm = np.random.default_rng(0).permutation(10)
    rec = np.zeros((10, 6), dtype="<f4"); rec[:, :2] = pts[perm]
    path = tmp_path / "decagon.ply"
    hdr = ("ply\nformat binary_little_endian 1.0\ncomment test\nelement vertex 10\nproperty float x\nproperty float y\nproperty float z\n"
           "property float nx\nproperty float ny\nproperty float nz\nelement face 0\nproperty list uchar int vertex_indices\nend_header\n")
    path.write_bytes(hdr.encode() + rec.tobytes())
    m = q.Model.from_ply(path, False, 3, 0.3, 0.2, 0.25, 0.02)
    assert m.n == 11 and m.nknots == 15 and np.allclose(m.P[0], m.P[-1])
    assert abs(m.P[0, 0] + 0.040) < 1e-7                          # starts at the min-x vertex, mm -> m
    seg = np.linalg.norm(np.diff(m.P, axis=0), axis=1)
    assert abs(seg.sum() - m.b) < 1e-6 and seg.max() < 0.03        # nearest-neighbour chain follows the outline
    with pytest.raises(q.QspushError):
        q.Model.from_ply(tmp_path / "missing.ply", False, 3, 0.3, 0.2, 0.25, 0.02)
    bad = tmp_path / "ascii.ply"; bad.write_text("ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nend_header\n0 0\n1 0\n0 1\n")
    with pytest.raises(q.QspushError):
        q.Model.from_ply(bad, False, 3, 0.3, 0.2, 0.25, 0.02)


def test_argument_errors():
    t = packaged_tables()["santal"]
    with pytest.raises(q.QspushError):
        q.Model.from_tables(t["knots"][:-1], t["ctrl_xy"], 3, 0.19, 0.02)      # nknots != n + p + 1
    with pytest.raises(q.QspushError):
        q.Model.from_tables(t["knots"], t["ctrl_xy"], 2, 0.19, 0.02)           # only cubic outlines


@pytest.mark.skipif(HAVE_GPU, reason="checks the behaviour on a box without a GPU")
def test_compute_fails_loudly_without_gpu():
    gm = gpu_model("santal")
    with pytest.raises(q.QspushError, match="no CUDA device"):
        gm.eval_spline([0.0])
    with pytest.raises(q.QspushError, match="no CUDA device"):
        gm.eval_erk4_sens(np.zeros((1, 4)), np.array([[0.01, 0.0]]), 0.05)
    with pytest.raises(q.QspushError, match="no CUDA device"):
        q.Solver([gm], 10, 0.05, 4)
    s = q.object_selection("santal")
    with pytest.raises(q.QspushError, match="no CUDA device"):
        q.PusherSliderModel("real_plant", s, 0, s.cad_model_path, 3, s.pcl_path, "santal")


def test_product_never_touches_the_oracle():
    """The package must not import, link or execute anything under oracle/ or tests/hostsim."""
    pkg = os.path.join(ROOT, "uclv_qs_pushing_matlab_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("the oracle", "").replace("CPU oracle", "") or f in ("qs_solver.cuh",), (f, "mentions oracle")
                assert "hostsim" not in src or f.endswith((".cuh",)), f
    import subprocess
    out = subprocess.run(["ldd", L.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "hostsim" not in out


def test_mex_gateway_compiles_against_the_header():
    """matlab/qspush_mex.c (the binding INTEGRATION.md hands to a maintainer of the reference) is type-checked against
    include/qspush.h with a stub of the MEX API: MATLAB itself does not exist in this container."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types",
                        "-fsyntax-only", "-I" + os.path.join(root, "tests", "stubs"), "-I" + os.path.join(root, "include"),
                        os.path.join(root, "matlab", "qspush_mex.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
