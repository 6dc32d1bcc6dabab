"""ctypes binding of oracle/libqs_arbiter.so — TEST INFRASTRUCTURE ONLY (see qs_arbiter.cpp).

`solve_exact(qp_data, guess)` returns the __float128 active-set solution of each QP of a batch (rounded to FP64) together
with its KKT certificate, so that the oracle's IPM and the CUDA kernels can each be compared with the exact solution.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libqs_arbiter.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "qs_arbiter.cpp")
        if not os.path.exists(_LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(_LIB_PATH):
            subprocess.check_call(["make", "-C", _HERE, "-s", "libqs_arbiter.so"])
        L = C.CDLL(_LIB_PATH)
        L.arb_qp_solve.restype = C.c_int
        L.arb_qp_solve.argtypes = [C.c_int] + [_dp] * 10 + [_ip, _ip, _dp, _ip] + [_dp] * 5 + [_ip]
        _lib = L
    return _lib


def working_set_from_ipm(lam, t, on, kappa=1e-6):
    """Initial working set (-1 lower, 0 free, +1 upper) from an IPM point: a side is active when lam > kappa * t."""
    lam = np.asarray(lam); t = np.asarray(t)
    lo = lam[..., :3] > kappa * t[..., :3]
    up = lam[..., 3:] > kappa * t[..., 3:]
    act = np.where(lo, -1, np.where(up, 1, 0)).astype(np.int32)
    return np.ascontiguousarray(act * (np.asarray(on) != 0))


def solve_exact(d, act=None, nthreads=8):
    """d = Ocp.qp_data(...) of the oracle.  Returns dict(du, dx, pi, lam, act, kkt [nb][5], iters, status)."""
    L = lib()
    nb, N = d["g"].shape[:2]
    act = np.zeros((nb, N, 3), dtype=np.int32) if act is None else np.ascontiguousarray(act, dtype=np.int32).copy()
    out = dict(du=np.zeros((nb, N, 2)), dx=np.zeros((nb, N + 1, 4)), pi=np.zeros((nb, N, 4)), lam=np.zeros((nb, N, 6)),
               act=act, kkt=np.zeros((nb, 5)), iters=np.zeros(nb, dtype=np.int32), status=np.zeros(nb, dtype=np.int32))
    arrs = {k: np.ascontiguousarray(d[k], dtype=np.float64) for k in ("H", "g", "A", "B", "b", "QN", "qN", "dx0", "dl", "du", "beta")}
    on = np.ascontiguousarray(d["on"], dtype=np.int32); ci = np.ascontiguousarray(d["ci"], dtype=np.int32)

    def one(i):
        p = lambda a: a[i].ctypes.data_as(_dp)                     # noqa: E731
        it = C.c_int(0)
        out["status"][i] = L.arb_qp_solve(
            N, p(arrs["H"]), p(arrs["g"]), p(arrs["A"]), p(arrs["B"]), p(arrs["b"]), arrs["QN"].ctypes.data_as(_dp),
            p(arrs["qN"]), p(arrs["dx0"]), p(arrs["dl"]), p(arrs["du"]), on[i].ctypes.data_as(_ip), ci[i].ctypes.data_as(_ip),
            p(arrs["beta"]), act[i].ctypes.data_as(_ip), p(out["du"]), p(out["dx"]), p(out["pi"]), p(out["lam"]),
            p(out["kkt"]), C.byref(it))
        out["iters"][i] = it.value

    with ThreadPoolExecutor(max_workers=max(1, nthreads)) as ex:
        list(ex.map(one, range(nb)))
    return out
