"""ctypes binding of tests/hostsim/libqspush_hostsim.so — TEST-ONLY host simulation of the CUDA kernels.

The kernel bodies in uclv_qs_pushing_matlab_b200/csrc/*.cuh are __host__ __device__; this module runs them
thread-by-thread on the CPU so `pytest -m "not gpu"` can check the kernel logic against the oracle in a
container without a GPU.  The product never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libqspush_hostsim.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = C.CDLL(_LIB)
        L.hs_last_error.restype = C.c_char_p
        L.hs_model_create.restype = C.c_void_p
        L.hs_model_create.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
        L.hs_model_from_ply.restype = C.c_void_p
        L.hs_model_from_ply.argtypes = [C.c_char_p, C.c_int, C.c_int] + [C.c_double] * 4
        L.hs_model_free.argtypes = [C.c_void_p]
        L.hs_model_info.argtypes = [C.c_void_p, _ip, _ip, _dp, _dp, _dp]
        L.hs_model_tables.argtypes = [C.c_void_p] + [_dp] * 5
        L.hs_eval_spline.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int] + [_dp] * 6
        L.hs_eval_dynamics.argtypes = [C.c_void_p, C.c_int] + [_dp] * 5
        L.hs_eval_erk4.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_double, _dp, _dp, _dp]
        L.hs_eval_vbound.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int, _dp, _dp]
        L.hs_solve.argtypes = ([C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_double, C.c_int, _ip] + [_dp] * 4 +
                               [_dp, _ip, _dp] + [_dp] * 7 + [_ip] + [_ip, _dp] + [_dp] * 7)
        _lib = L
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _c(a, dt=np.float64):
    return np.ascontiguousarray(a, dtype=dt)


class Model:
    def __init__(self, h):
        if not h:
            raise RuntimeError("hostsim model: " + lib().hs_last_error().decode())
        self.h = h
        n, nk = C.c_int(), C.c_int()
        b, ce, mu = C.c_double(), C.c_double(), C.c_double()
        lib().hs_model_info(h, C.byref(n), C.byref(nk), C.byref(b), C.byref(ce), C.byref(mu))
        self.n, self.nknots, self.b, self.c_ellipse, self.mu_sp = n.value, nk.value, b.value, ce.value, mu.value
        self.S = np.zeros(self.nknots); self.P = np.zeros((self.n, 2)); self.c1 = np.zeros((self.n, 2)); self.c2 = np.zeros((self.n, 2))
        self.blob = np.zeros(lib().hs_model_doubles())
        lib().hs_model_tables(h, _d(self.S), _d(self.P), _d(self.c1), _d(self.c2), _d(self.blob))

    @classmethod
    def create(cls, S, P, p, mu_sp, c_ellipse, single=True):
        S, P = _c(S), _c(P)
        return cls(lib().hs_model_create(_d(S), len(S), _d(P), P.shape[0], p, mu_sp, c_ellipse, int(single)))

    @classmethod
    def from_ply(cls, path, flip, p, mu_sg, mu_sp, mass, tau_max):
        return cls(lib().hs_model_from_ply(str(path).encode(), int(flip), p, mu_sg, mu_sp, mass, tau_max))

    def __del__(self):
        try:
            lib().hs_model_free(self.h)
        except Exception:
            pass

    def eval_spline(self, s, wrap=0):
        s = _c(np.atleast_1d(s)); k = len(s)
        o = dict(C=np.zeros((k, 2)), Cd=np.zeros((k, 2)), Cdd=np.zeros((k, 2)), t=np.zeros((k, 2)), n=np.zeros((k, 2)), kappa=np.zeros(k))
        lib().hs_eval_spline(self.h, k, _d(s), wrap, _d(o["C"]), _d(o["Cd"]), _d(o["Cdd"]), _d(o["t"]), _d(o["n"]), _d(o["kappa"]))
        return o

    def dynamics(self, x, u):
        x, u = _c(np.atleast_2d(x)), _c(np.atleast_2d(u)); k = x.shape[0]
        f, Jx, Ju = np.zeros((k, 4)), np.zeros((k, 4, 4)), np.zeros((k, 4, 2))
        lib().hs_eval_dynamics(self.h, k, _d(x), _d(u), _d(f), _d(Jx), _d(Ju))
        return f, Jx, Ju

    def erk4_sens(self, x, u, dt):
        x, u = _c(np.atleast_2d(x)), _c(np.atleast_2d(u)); k = x.shape[0]
        Phi, A, B = np.zeros((k, 4)), np.zeros((k, 4, 4)), np.zeros((k, 4, 2))
        lib().hs_eval_erk4(self.h, k, _d(x), _d(u), float(dt), _d(Phi), _d(A), _d(B))
        return Phi, A, B

    def v_bound(self, s, ctrl=(1.0, 0.0, 3.0, 0.05, 0.0), single=True):
        s = _c(np.atleast_1d(s)); k = len(s)
        vb, ta = np.zeros(k), np.zeros(k)
        c5 = _c(ctrl)
        lib().hs_eval_vbound(self.h, k, _d(s), _d(c5), int(single), _d(vb), _d(ta))
        return vb, ta


def solve(models, N, dt, x0, yref, yref_e, x, u, pi=None, lam=None, cold=None, objid=None, W=None, We=None,
          lh=(-0.06, 0.0, -0.05), uh=(0.011, 0.03, 0.05), mode="rti", prepare=False, shift=False,
          qp_tol=1e-11, qp_mu0=0.1, qp_thr=1e-3, qp_tau=0.9995, qp_tol_comp=1e-18, qp_t_min=1e-12, qp_gamma_f=0.01, qp_stall=10, qp_max_iter=50, max_sqp_iter=30,
          tol=(1e-6, 1e-6, 1e-6, 1e-6), globalization=1, alpha_min=0.05, alpha_red=0.7, eps_sd=1e-4,
          single_quirk=True, ctrl=(1.0, 0.0, 3.0, 0.05, 0.0), qp_kernel=1, h_variant=0):
    nb = np.asarray(x0).shape[0]
    if W is None:
        W = np.tile(np.diag([1.0, 1.0, 1e-3, 0.0, 1e-3, 1e-3]).reshape(1, 36), (N, 1))
    if We is None:
        We = np.diag([2e5, 2e5, 20.0, 0.0])
    W = _c(np.asarray(W).reshape(N, 36)); We = np.asfortranarray(We, dtype=np.float64)
    x0 = _c(x0).copy(); yref = _c(yref); yref_e = _c(yref_e); x = _c(x).copy(); u = _c(u).copy()
    pi = np.zeros((nb, N, 4)) if pi is None else _c(pi).copy()
    lam = np.zeros((nb, N, 6)) if lam is None else _c(lam).copy()
    cold = np.zeros(nb, dtype=np.int32) if cold is None else _c(cold, np.int32).copy()
    objid = np.zeros(nb, dtype=np.int32) if objid is None else _c(objid, np.int32)
    od = _c([qp_tol, qp_mu0, qp_thr, qp_tau, *tol, alpha_min, alpha_red, eps_sd, qp_tol_comp, qp_t_min, qp_gamma_f, qp_stall])
    oi = _c([{"rti": 0, "sqp": 1, "qp": 2}[mode], qp_max_iter, max_sqp_iter, globalization, int(single_quirk), int(prepare), int(shift), int(qp_kernel), int(h_variant)], np.int32)
    c5 = _c(ctrl); lh = _c(lh); uh = _c(uh)
    si = np.zeros((nb, 3), dtype=np.int32); sd = np.zeros((nb, 6))
    z = np.zeros((nb, N + 1, 6)); qpi = np.zeros((nb, N, 4)); qlam = np.zeros((nb, N, 6))
    A = np.zeros((nb, N, 8)); B = np.zeros((nb, N, 8)); bb = np.zeros((nb, N, 4)); g = np.zeros((nb, N, 6))
    arr = (C.c_void_p * len(models))(*[m.h for m in models])
    rc = lib().hs_solve(arr, len(models), N, float(dt), nb, _i(objid), _d(W), We.ctypes.data_as(_dp), _d(lh), _d(uh),
                   _d(od), _i(oi), _d(c5), _d(x0), _d(yref), _d(yref_e), _d(x), _d(u), _d(pi), _d(lam), _i(cold),
                   _i(si), _d(sd), _d(z), _d(qpi), _d(qlam), _d(A), _d(B), _d(bb), _d(g))
    if rc != 0:
        raise RuntimeError("hs_solve: " + lib().hs_last_error().decode())
    return dict(x0=x0, x=x, u=u, pi=pi, lam=lam, cold=cold, status=si[:, 0], sqp_iter=si[:, 1], qp_iter=si[:, 2],
                cost=sd[:, 0], res=sd[:, 1:5], alpha=sd[:, 5], du=z[:, :N, :2], dx=z[:, :, 2:], qp_pi=qpi, qp_lam=qlam,
                A=A, B=B, b=bb, g=g)


def closed_loop(models, N, dt, traj, x, steps, offset=None, objid=None, W=None, We=None,
                lh=(-0.06, 0.0, -0.05), uh=(0.011, 0.03, 0.05), mode="rti",
                qp_tol=1e-11, qp_mu0=0.1, qp_thr=1e-3, qp_tau=0.9995, qp_tol_comp=1e-18, qp_t_min=1e-12, qp_gamma_f=0.01, qp_stall=10, qp_max_iter=50, max_sqp_iter=30,
                tol=(1e-6, 1e-6, 1e-6, 1e-6), globalization=1, alpha_min=0.05, alpha_red=0.7, eps_sd=1e-4,
                single_quirk=True, ctrl=(1.0, 0.0, 3.0, 0.05, 0.0), qp_kernel=1, h_variant=0,
                idx0=1, noise_sigma=(0.0, 0.0, 0.0, 0.0), seed=0, t_dist=0, amplitude_dist=0.0, xwidth=0.0, delay_plant=0, delay_comp=0):
    """Host mirror of qspush_closed_loop (same kernel bodies, same order)."""
    x = _c(x).copy(); nb = x.shape[0]
    if W is None:
        W = np.tile(np.diag([1.0, 1.0, 1e-3, 0.0, 1e-3, 1e-3]).reshape(1, 36), (N, 1))
    if We is None:
        We = np.diag([2e5, 2e5, 20.0, 0.0])
    W = _c(np.asarray(W).reshape(N, 36)); We = np.asfortranarray(We, dtype=np.float64)
    traj = _c(traj); T = traj.shape[0]
    objid = np.zeros(nb, dtype=np.int32) if objid is None else _c(objid, np.int32)
    od = _c([qp_tol, qp_mu0, qp_thr, qp_tau, *tol, alpha_min, alpha_red, eps_sd, qp_tol_comp, qp_t_min, qp_gamma_f, qp_stall])
    oi = _c([{"rti": 0, "sqp": 1}[mode], qp_max_iter, max_sqp_iter, globalization, int(single_quirk), 1, 1, int(qp_kernel), int(h_variant)], np.int32)
    c5 = _c(ctrl); lh = _c(lh); uh = _c(uh)
    ld = _c([*noise_sigma, amplitude_dist, xwidth]); li = _c([idx0, t_dist, delay_plant, delay_comp], np.int32)
    lx = np.zeros((steps, nb, 4)); lu = np.zeros((steps, nb, 2)); ls = np.zeros((steps, nb), dtype=np.int32)
    off = None if offset is None else _c(offset)
    arr = (C.c_void_p * len(models))(*[m.h for m in models])
    f = lib().hs_closed_loop
    f.restype = C.c_int
    rc = f(arr, len(models), N, C.c_double(dt), nb, _i(objid), _d(W), We.ctypes.data_as(_dp), _d(lh), _d(uh), _d(od), _i(oi), _d(c5),
           _d(traj), T, (_d(off) if off is not None else None), _d(x), steps, _d(ld), _i(li), C.c_ulonglong(seed), _d(lx), _d(lu), _i(ls))
    if rc != 0:
        raise RuntimeError("hs_closed_loop: " + lib().hs_last_error().decode())
    return dict(x=x, x_log=lx, u_log=lu, status_log=ls)
