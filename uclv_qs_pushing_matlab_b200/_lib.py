"""Loader of the in-tree C-ABI library (uclv_qs_pushing_matlab_b200/libqspush.so, include/qspush.h).

The library is the product: there is no Python or CPU fallback.  Importing the package without the
built library raises, and every compute entry point fails loudly without a CUDA device
(QSPUSH_ERR_NO_DEVICE).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqspush.so")
CSRC = os.path.join(_HERE, "csrc")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
vp = C.c_void_p


class QspushError(RuntimeError):
    pass


class Opts(C.Structure):
    """qspush_opts (include/qspush.h)."""
    _fields_ = [
        ("mode", C.c_int), ("max_sqp_iter", C.c_int),
        ("tol_stat", C.c_double), ("tol_eq", C.c_double), ("tol_ineq", C.c_double), ("tol_comp", C.c_double),
        ("qp_max_iter", C.c_int), ("qp_tol", C.c_double), ("qp_mu0", C.c_double), ("qp_thr", C.c_double),
        ("qp_tau", C.c_double), ("globalization", C.c_int),
        ("alpha_min", C.c_double), ("alpha_reduction", C.c_double), ("eps_sufficient_descent", C.c_double),
        ("matlab_single_quirk", C.c_int), ("problems_per_warp", C.c_int), ("qp_kernel", C.c_int),
        ("h_variant", C.c_int),
        ("qp_tol_comp", C.c_double), ("qp_t_min", C.c_double), ("qp_gamma_f", C.c_double), ("qp_stall", C.c_int),
    ]


class Ctrl(C.Structure):
    """qspush_ctrl (include/qspush.h)."""
    _fields_ = [("v_alpha", C.c_double), ("d_v_bound", C.c_double), ("t_angle0", C.c_double),
                ("u_t_ub", C.c_double), ("u_n_lb", C.c_double)]


class LoopOpts(C.Structure):
    """qspush_loop_opts (include/qspush.h)."""
    _fields_ = [("idx0", C.c_int), ("noise_sigma", C.c_double * 4), ("seed", C.c_ulonglong), ("t_dist", C.c_int),
                ("amplitude_dist", C.c_double), ("xwidth", C.c_double), ("delay_plant", C.c_int), ("delay_comp", C.c_int)]


# qspush_field / qspush_mem / qspush_mode / qspush_stat
X0, YREF, YREF_E, X, U, PI, LAM, COST, RES = 0, 1, 2, 3, 4, 5, 6, 7, 8
W, LH, UH = 16, 17, 18
STATUS, SQP_ITER, QP_ITER, OBJECT_ID, COLD = 32, 33, 34, 35, 36
MEM_HOST, MEM_DEVICE = 0, 1
MODE_RTI, MODE_SQP = 0, 1
TIME_TOT, TIME_LIN, TIME_QP, TIME_PREP = 0, 1, 2, 3
STEP_SHIFT, STEP_RESTORE_GUESS = 1, 2

# every symbol include/qspush.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "qspush_last_error": (C.c_char_p, []),
    "qspush_version": (C.c_char_p, []),
    "qspush_device_count": (C.c_int, []),
    "qspush_model_create": (C.c_int, [dp, C.c_int, dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(vp)]),
    "qspush_model_create_from_ply": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.POINTER(vp)]),
    "qspush_model_free": (None, [vp]),
    "qspush_model_info": (C.c_int, [vp, ip, ip, dp, dp, dp]),
    "qspush_model_tables": (C.c_int, [vp, dp, dp, dp, dp]),
    "qspush_eval_spline": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, vp]),
    "qspush_eval_dynamics": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp]),
    "qspush_eval_erk4_sens": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_double, vp, vp, vp]),
    "qspush_eval_v_bound": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, C.POINTER(Ctrl), C.c_int, vp, vp]),
    "qspush_opts_default": (None, [C.POINTER(Opts)]),
    "qspush_ctrl_default": (None, [C.POINTER(Ctrl)]),
    "qspush_solver_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.POINTER(Opts), C.POINTER(vp)]),
    "qspush_solver_free": (None, [vp]),
    "qspush_solver_set_opts": (C.c_int, [vp, C.POINTER(Opts)]),
    "qspush_solver_set_ctrl": (C.c_int, [vp, C.POINTER(Ctrl)]),
    "qspush_set": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]),
    "qspush_get": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]),
    "qspush_set_int": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]),
    "qspush_get_int": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]),
    "qspush_prepare": (C.c_int, [vp]),
    "qspush_solve": (C.c_int, [vp]),
    "qspush_shift": (C.c_int, [vp]),
    "qspush_plant_step": (C.c_int, [vp, vp, vp, C.c_int]),
    "qspush_set_reference_trajectory": (C.c_int, [vp, vp, C.c_int, vp, C.c_int]),
    "qspush_set_reference_window": (C.c_int, [vp, C.c_int]),
    "qspush_closed_loop": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int, C.POINTER(LoopOpts), vp, vp, vp, C.c_int]),
    "qspush_step": (C.c_int, [vp, vp, C.c_int, C.c_uint, vp, vp, C.c_int]),
    "qspush_snapshot_guess": (C.c_int, [vp]),
    "qspush_measure_fp64_peak": (C.c_int, [C.c_int, dp]),
    "qspush_sync": (C.c_int, [vp]),
    "qspush_stream": (vp, [vp]),
    "qspush_get_stat": (C.c_int, [vp, C.c_int, dp]),
    "qspush_launch_count": (C.c_longlong, [vp]),
}


def build(verbose: bool = False) -> str:
    """Compile libqspush.so in-tree with nvcc for sm_100a (make -C csrc)."""
    subprocess.check_call(["make", "-C", CSRC] + ([] if verbose else ["-s"]))
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QspushError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C uclv_qs_pushing_matlab_b200/csrc`.  There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)      # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise QspushError(f"qspush error {rc}: {lib().qspush_last_error().decode()}")
