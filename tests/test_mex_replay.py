"""The MEX gateway executed outside MATLAB (SURVEY 8b / VERDICT r01 row N1): matlab/qspush_mex.c compiled against a working
miniature of the MEX array API (tests/stubs/mex_runtime.c) and driven by tests/mex_replay.c, which issues the mexFunction calls of
matlab/qspush_ocp.m under the unmodified NMPC_controller.solve / helper.closed_loop_matlab sequence (config 1: 201 periods).

  * CPU: the three copies of the field table (qspush_ocp.m `F`, mex_replay.c, include/qspush.h) agree; the gateway, the runtime
    and the driver compile and link; without a GPU the gateway turns the C-ABI's QSPUSH_ERR_NO_DEVICE into a MEX error (the
    process ends like mexErrMsgIdAndTxt does, nothing is computed on the CPU);
  * GPU: both solver modes, every period bit-for-bit equal to the same sequence through include/qspush.h directly.
"""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "mex_replay")
SRCS = [os.path.join(ROOT, "tests", "mex_replay.c"), os.path.join(ROOT, "tests", "stubs", "mex_runtime.c"), os.path.join(ROOT, "matlab", "qspush_mex.c")]


def _exe():
    deps = SRCS + [os.path.join(ROOT, "tests", "stubs", "mex.h"), os.path.join(ROOT, "include", "qspush.h")]
    if not os.path.exists(EXE) or max(os.path.getmtime(p) for p in deps) > os.path.getmtime(EXE):
        lib = os.path.join(ROOT, "uclv_qs_pushing_matlab_b200")
        subprocess.check_call(["gcc", "-O1", "-std=gnu99", "-Wall", "-I" + os.path.join(ROOT, "tests", "stubs"), "-I" + os.path.join(ROOT, "include"), *SRCS,
                               "-L" + lib, "-lqspush", "-Wl,-rpath," + lib, "-Wl,-rpath,$ORIGIN/../uclv_qs_pushing_matlab_b200", "-lm", "-o", EXE])
    return EXE


def _santal_ply(path):
    """A binary_little_endian outline in the layout of the reference's files (nv x 6 float32: x y z nx ny nz, millimetres,
    SURVEY A1.2) made from the packaged santal control points, so that the replay runs on the santal geometry on a box without
    /root/reference.  Both runs of mex_replay ingest this same file."""
    from uclv_qs_pushing_matlab_b200.workloads import packaged_tables
    P = np.asarray(packaged_tables()["santal"]["ctrl_xy"], dtype=np.float64)[:-1]           # closed polygon: drop the repeated first point
    v = np.zeros((len(P), 6), dtype="<f4"); v[:, :2] = (P * 1000.0).astype(np.float32); v[:, 5] = 1.0
    hdr = ("ply\nformat binary_little_endian 1.0\ncomment synthetic santal outline\nelement vertex %d\n" % len(P)
           + "".join("property float %s\n" % n for n in ("x", "y", "z", "nx", "ny", "nz")) + "end_header\n")
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii")); f.write(v.tobytes())
    return path


def test_field_tables_of_matlab_class_driver_and_header_agree():
    m = open(os.path.join(ROOT, "matlab", "qspush_ocp.m")).read()
    body = re.search(r"F = struct\((.*?)\);", m, re.S).group(1).replace("...", " ")
    F_m = {k: int(v) for k, v in re.findall(r"'(\w+)'\s*,\s*(\d+)", body)}
    c = open(os.path.join(ROOT, "tests", "mex_replay.c")).read()
    F_c = {k: int(v) for k, v in re.findall(r'\{"(\w+)", (\d+)\}', re.search(r"static const fent F\[\] = \{(.*?)\};", c, re.S).group(1))}
    assert F_m == F_c and len(F_m) == 14
    h = open(os.path.join(ROOT, "include", "qspush.h")).read()
    enum = {k: int(v) for k, v in re.findall(r"QSPUSH_(\w+) = (\d+)", h)}
    want = dict(constr_x0="X0", cost_y_ref="YREF", cost_y_ref_e="YREF_E", init_x="X", init_u="U", init_pi="PI", x="X", u="U", pi="PI",
                cost_W="W", constr_lh="LH", constr_uh="UH", status="STATUS", sqp_iter="SQP_ITER")
    for f, e in want.items():
        assert F_m[f] == enum[e], (f, e)
    assert enum["COST"] == 7 and enum["RES"] == 8               # get_cost / print of qspush_ocp.m use the numbers directly
    # the stage defaulting of single-stage fields (set('cost_y_ref_e', y, Hp) of NMPC_controller.m:348) is in the class
    assert "any(strcmp(field, {'constr_x0','cost_y_ref_e'})), stage = -1" in m


def test_gateway_links_and_reports_a_missing_gpu_as_a_mex_error(tmp_path):
    import torch
    exe = _exe()
    ply = _santal_ply(str(tmp_path / "santal.ply"))
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run itself is the gpu test below")
    r = subprocess.run([exe, ply, "0", "3", str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 4 and "MEX error qspush:error" in r.stderr, (r.returncode, r.stderr)   # model_from_ply worked on the host, solver_create refused
    assert not os.path.exists(tmp_path / "o.bin")


def test_synthetic_outline_reproduces_the_santal_model(tmp_path):
    import uclv_qs_pushing_matlab_b200 as q
    from uclv_qs_pushing_matlab_b200.workloads import packaged_tables
    t = packaged_tables()["santal"]
    o = q.OBJECT_TABLE["santal"]
    m = q.Model.from_ply(_santal_ply(str(tmp_path / "santal.ply")), False, 3, o["mu_sg"], o["mu_sp"], o["m"], o["tau_max"])
    assert abs(m.b - float(np.asarray(t["knots"])[-1])) < 1e-6 and m.n == len(t["ctrl_xy"])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1], ids=["sqp_rti", "sqp"])
def test_mex_gateway_replay_of_config1_is_the_direct_c_abi_run_bit_for_bit(tmp_path, mode):
    exe = _exe()
    ply = _santal_ply(str(tmp_path / "santal.ply"))
    out = str(tmp_path / "o.bin")
    r = subprocess.run([exe, ply, str(mode), "201", out], capture_output=True, text=True)
    assert r.returncode == 0 and "bit-identical" in r.stdout, (r.returncode, r.stdout, r.stderr)
    raw = np.fromfile(out, dtype=np.float64).reshape(2, 201, 10)
    assert np.array_equal(raw[0], raw[1])
    rows = raw[0]
    if mode == 0:
        assert (rows[:, 6] == 0).all() and (rows[:, 7] == 1).all()
    else:
        assert set(np.unique(rows[:, 6])) <= {0.0, 2.0, 3.0, 4.0} and rows[:, 7].max() <= 30
    assert abs(rows[-1, 0] - 0.10) < 3e-3 and np.isfinite(rows).all()     # the slider arrives at the end of the 0.10 m reference
    assert struct.calcsize("d") == 8
