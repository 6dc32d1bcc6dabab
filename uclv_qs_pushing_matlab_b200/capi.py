"""Thin object layer over the C-ABI (include/qspush.h): `Model` and `Solver`.

Buffers may be numpy float64 arrays (host), torch CPU tensors (host, pinned or not) or torch CUDA
tensors on the solver's device (device pointers, no copies).  Layouts are the C-ABI's:
[batch][stage][dim] for stage = -1, [batch][dim] otherwise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

try:  # torch is plumbing only (device buffers, streams); the package works with numpy alone
    import torch
except Exception:  # pragma: no cover
    torch = None


def _buf(a, dtype=np.float64, writable=False):
    """-> (address, mem, keepalive)"""
    if torch is not None and isinstance(a, torch.Tensor):
        want = torch.float64 if dtype == np.float64 else torch.int32
        if a.dtype != want or not a.is_contiguous():
            raise L.QspushError("torch buffers must be contiguous %s" % want)
        return a.data_ptr(), (L.MEM_DEVICE if a.is_cuda else L.MEM_HOST), a
    arr = np.ascontiguousarray(a, dtype=dtype)
    if writable and arr is not a:
        raise L.QspushError("output numpy buffers must be contiguous arrays of the right dtype")
    return arr.ctypes.data, L.MEM_HOST, arr


class Model:
    """PusherSliderModel + bspline_shape tables (qspush_model)."""

    def __init__(self, handle):
        self._h = handle
        n, nk = C.c_int(), C.c_int()
        b, ce, mu = C.c_double(), C.c_double(), C.c_double()
        L.check(L.lib().qspush_model_info(self._h, C.byref(n), C.byref(nk), C.byref(b), C.byref(ce), C.byref(mu)))
        self.n, self.nknots, self.b, self.c_ellipse, self.mu_sp = n.value, nk.value, b.value, ce.value, mu.value
        self.p = self.nknots - self.n - 1
        self.S = np.zeros(self.nknots)
        self.P = np.zeros((self.n, 2))
        self.cj_1_vect = np.zeros((self.n, 2))
        self.cj_2_vect = np.zeros((self.n, 2))
        L.check(L.lib().qspush_model_tables(self._h, self.S.ctypes.data_as(L.dp), self.P.ctypes.data_as(L.dp),
                                            self.cj_1_vect.ctypes.data_as(L.dp), self.cj_2_vect.ctypes.data_as(L.dp)))

    @classmethod
    def from_tables(cls, knots, ctrl_xy, degree, mu_sp, c_ellipse, single_coeffs=True):
        S = np.ascontiguousarray(knots, dtype=np.float64)
        P = np.ascontiguousarray(ctrl_xy, dtype=np.float64)
        h = L.vp()
        L.check(L.lib().qspush_model_create(S.ctypes.data_as(L.dp), len(S), P.ctypes.data_as(L.dp), P.shape[0],
                                            int(degree), float(mu_sp), float(c_ellipse), int(single_coeffs), C.byref(h)))
        return cls(h)

    @classmethod
    def from_ply(cls, ply_path, flip_order, degree, mu_sg, mu_sp, mass, tau_max):
        h = L.vp()
        L.check(L.lib().qspush_model_create_from_ply(str(ply_path).encode(), int(flip_order), int(degree),
                                                     float(mu_sg), float(mu_sp), float(mass), float(tau_max), C.byref(h)))
        return cls(h)

    def __del__(self):
        try:
            if self._h:
                L.lib().qspush_model_free(self._h)
                self._h = None
        except Exception:
            pass

    # ---- stateless batched evaluation (numpy in/out; device = CUDA ordinal)
    def eval_spline(self, s, wrap=0, device=0, want=("C", "Cd", "Cdd", "t", "n", "kappa")):
        s = np.ascontiguousarray(np.atleast_1d(s), dtype=np.float64)
        k = len(s)
        out = {"C": np.zeros((k, 2)), "Cd": np.zeros((k, 2)), "Cdd": np.zeros((k, 2)),
               "t": np.zeros((k, 2)), "n": np.zeros((k, 2)), "kappa": np.zeros(k)}
        ptr = [out[nm].ctypes.data if nm in want else None for nm in ("C", "Cd", "Cdd", "t", "n", "kappa")]
        L.check(L.lib().qspush_eval_spline(self._h, device, L.MEM_HOST, k, s.ctypes.data, int(wrap), *ptr))
        return {nm: out[nm] for nm in want}

    def eval_dynamics(self, x, u, jac=False, device=0):
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        u = np.ascontiguousarray(np.atleast_2d(u), dtype=np.float64)
        k = x.shape[0]
        f = np.zeros((k, 4))
        if not jac:
            L.check(L.lib().qspush_eval_dynamics(self._h, device, L.MEM_HOST, k, x.ctypes.data, u.ctypes.data, f.ctypes.data, None, None))
            return f
        Jx, Ju = np.zeros((k, 4, 4)), np.zeros((k, 4, 2))
        L.check(L.lib().qspush_eval_dynamics(self._h, device, L.MEM_HOST, k, x.ctypes.data, u.ctypes.data, f.ctypes.data,
                                             Jx.ctypes.data, Ju.ctypes.data))
        return f, Jx, Ju

    def eval_erk4_sens(self, x, u, dt, device=0):
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        u = np.ascontiguousarray(np.atleast_2d(u), dtype=np.float64)
        k = x.shape[0]
        Phi, A, B = np.zeros((k, 4)), np.zeros((k, 4, 4)), np.zeros((k, 4, 2))
        L.check(L.lib().qspush_eval_erk4_sens(self._h, device, L.MEM_HOST, k, x.ctypes.data, u.ctypes.data, float(dt),
                                              Phi.ctypes.data, A.ctypes.data, B.ctypes.data))
        return Phi, A, B

    def eval_erk4_sens_device(self, x, u, dt, Phi, A, B):
        """torch CUDA tensors in and out (no copies): x [n,4], u [n,2], Phi [n,4], A [n,4,4], B [n,4,2]."""
        n = x.shape[0]
        L.check(L.lib().qspush_eval_erk4_sens(self._h, x.device.index or 0, L.MEM_DEVICE, n, x.data_ptr(), u.data_ptr(), float(dt),
                                              Phi.data_ptr(), A.data_ptr() if A is not None else None,
                                              B.data_ptr() if B is not None else None))

    def eval_v_bound(self, s, ctrl=None, single_quirk=True, device=0):
        s = np.ascontiguousarray(np.atleast_1d(s), dtype=np.float64)
        k = len(s)
        vb, ta = np.zeros(k), np.zeros(k)
        c = ctrl if ctrl is not None else default_ctrl()
        L.check(L.lib().qspush_eval_v_bound(self._h, device, L.MEM_HOST, k, s.ctypes.data, C.byref(c), int(single_quirk),
                                            vb.ctypes.data, ta.ctypes.data))
        return vb, ta


def measure_fp64_peak(device=0) -> float:
    """FP64 DFMA issue rate of the device in TFLOP/s (qspush_measure_fp64_peak)."""
    v = C.c_double()
    L.check(L.lib().qspush_measure_fp64_peak(int(device), C.byref(v)))
    return v.value


def default_opts(**kw) -> L.Opts:
    o = L.Opts()
    L.lib().qspush_opts_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise L.QspushError(f"unknown option {k}")
        setattr(o, k, v)
    return o


def default_ctrl(**kw) -> L.Ctrl:
    c = L.Ctrl()
    L.lib().qspush_ctrl_default(C.byref(c))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


_FIELDS = {"x0": L.X0, "yref": L.YREF, "yref_e": L.YREF_E, "x": L.X, "u": L.U, "pi": L.PI, "lam": L.LAM,
           "cost": L.COST, "res": L.RES, "W": L.W, "lh": L.LH, "uh": L.UH}
_IFIELDS = {"status": L.STATUS, "sqp_iter": L.SQP_ITER, "qp_iter": L.QP_ITER, "object_id": L.OBJECT_ID, "cold": L.COLD}


class Solver:
    """Batched OCP solver = the acados_ocp replacement (qspush_solver).

    Stream contract.  The solver owns a non-blocking CUDA stream (`Solver.stream`); every call enqueues its work there.
    Calls with HOST buffers return after their copies have completed.  Calls with torch CUDA tensors are asynchronous;
    with `order_with_torch = True` (default) every such call makes the solver's stream wait for the work already queued
    on torch's current stream (the producer of the tensors) and makes torch's current stream wait for the call's work
    (the consumer), so tensors can be used right away like the results of any torch op.  Set it to False when the caller
    orders the streams itself (bench.py does, to keep its CUDA-event windows free of extra event traffic)."""

    order_with_torch = True

    class _Ordered:
        def __init__(self, solver, mem):
            self.on = solver.order_with_torch and mem == L.MEM_DEVICE and torch is not None
            self.s = solver

        def __enter__(self):
            if self.on:
                dev = torch.device("cuda", self.s.device)
                self.ext = torch.cuda.ExternalStream(self.s.stream, device=dev)
                self.ext.wait_stream(torch.cuda.current_stream(dev))
            return self

        def __exit__(self, *exc):
            if self.on:
                torch.cuda.current_stream(torch.device("cuda", self.s.device)).wait_stream(self.ext)
            return False

    def __init__(self, models, N, dt, batch, device=0, **opts):
        self.models = list(models) if isinstance(models, (list, tuple)) else [models]
        self.N, self.dt, self.batch, self.device = int(N), float(dt), int(batch), int(device)
        self.opts = default_opts(**opts)
        arr = (L.vp * len(self.models))(*[m._h for m in self.models])
        h = L.vp()
        L.check(L.lib().qspush_solver_create(arr, len(self.models), self.N, self.dt, self.batch, self.device,
                                             C.byref(self.opts), C.byref(h)))
        self._h = h
        self.ctrl = default_ctrl()

    def __del__(self):
        try:
            if self._h:
                L.lib().qspush_solver_free(self._h)
                self._h = None
        except Exception:
            pass

    def set_opts(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.opts, k):
                raise L.QspushError(f"unknown option {k}")
            setattr(self.opts, k, v)
        L.check(L.lib().qspush_solver_set_opts(self._h, C.byref(self.opts)))

    def set_ctrl(self, **kw):
        for k, v in kw.items():
            setattr(self.ctrl, k, v)
        L.check(L.lib().qspush_solver_set_ctrl(self._h, C.byref(self.ctrl)))

    def _shape(self, field, stage, nb):
        N = self.N
        dims = {"x0": (4, 1), "yref": (6, N), "yref_e": (4, 1), "x": (4, N + 1), "u": (2, N), "pi": (4, N),
                "lam": (6, N), "cost": (1, 1), "res": (4, 1)}[field]
        if stage < 0 and dims[1] > 1:
            return (nb, dims[1], dims[0])
        return (nb, dims[0])

    def set(self, field, data, stage=-1, lo=0, hi=None):
        hi = self.batch if hi is None else hi
        if field == "W":
            a = np.asfortranarray(data, dtype=np.float64)
            L.check(L.lib().qspush_set(self._h, L.W, int(stage), 0, 0, a.ctypes.data, L.MEM_HOST))
            return
        if field in ("lh", "uh"):
            a = np.ascontiguousarray(data, dtype=np.float64)
            L.check(L.lib().qspush_set(self._h, _FIELDS[field], -1, 0, 0, a.ctypes.data, L.MEM_HOST))
            return
        ptr, mem, keep = _buf(data)
        want = int(np.prod(self._shape(field, stage, hi - lo)))
        have = keep.numel() if (torch is not None and isinstance(keep, torch.Tensor)) else keep.size
        if have != want:
            raise L.QspushError(f"{field}: expected {want} values, got {have}")
        with self._Ordered(self, mem):
            L.check(L.lib().qspush_set(self._h, _FIELDS[field], int(stage), lo, hi, ptr, mem))
        self._keep = keep

    def get(self, field, stage=-1, lo=0, hi=None, out=None):
        hi = self.batch if hi is None else hi
        if field in ("lh", "uh"):
            a = np.zeros(3)
            L.check(L.lib().qspush_get(self._h, _FIELDS[field], -1, 0, 0, a.ctypes.data, L.MEM_HOST))
            return a
        shape = self._shape(field, stage, hi - lo)
        if out is None:
            out = np.zeros(shape)
        ptr, mem, keep = _buf(out, writable=True)
        with self._Ordered(self, mem):
            L.check(L.lib().qspush_get(self._h, _FIELDS[field], int(stage), lo, hi, ptr, mem))
        if field == "cost" and isinstance(out, np.ndarray):
            return out.reshape(-1)
        return out

    def set_int(self, field, data, lo=0, hi=None):
        hi = self.batch if hi is None else hi
        ptr, mem, keep = _buf(data, dtype=np.int32)
        with self._Ordered(self, mem):
            L.check(L.lib().qspush_set_int(self._h, _IFIELDS[field], lo, hi, ptr, mem))

    def get_int(self, field, lo=0, hi=None, out=None):
        hi = self.batch if hi is None else hi
        if out is None:
            out = np.zeros(hi - lo, dtype=np.int32)
        ptr, mem, keep = _buf(out, dtype=np.int32, writable=True)
        with self._Ordered(self, mem):
            L.check(L.lib().qspush_get_int(self._h, _IFIELDS[field], lo, hi, ptr, mem))
        return out

    def prepare(self):
        L.check(L.lib().qspush_prepare(self._h))

    def solve(self):
        L.check(L.lib().qspush_solve(self._h))

    def shift(self):
        L.check(L.lib().qspush_shift(self._h))

    def plant_step(self, x, u):
        """x <- x + dt*f(x,u) in place; numpy (host) or torch CUDA (device) [batch,4] / [batch,2]."""
        px, mem, kx = _buf(x, writable=True)
        pu, mem_u, ku = _buf(u)
        if mem != mem_u:
            raise L.QspushError("x and u must live in the same memory space")
        with self._Ordered(self, mem):
            L.check(L.lib().qspush_plant_step(self._h, px, pu, mem))
        return x

    def closed_loop(self, traj, x, steps, offset=None, idx0=1, noise_sigma=(0.0, 0.0, 0.0, 0.0), seed=0, t_dist=0,
                    amplitude_dist=0.0, xwidth=0.0, log=True, delay_plant=0, delay_comp=0):
        """Device-resident closed loop (qspush_closed_loop; helper.closed_loop_matlab for the whole batch).

        traj (T,6) reference columns [x_ref; u_ref]; x (batch,4) initial plant state (updated in place); offset
        (batch,6) optional per-problem shift of the reference.  numpy arrays (host) or torch CUDA tensors (device, no
        host round trip at all); returns dict(x=final state, x_log (steps,batch,4), u_log (steps,batch,2),
        status_log (steps,batch)) in the same memory space.  delay_plant / delay_comp: input delays in control periods
        (helper.m:211, NMPC_controller.m:108); with delay_comp pass the padded reference and idx0 = 1 + delay_comp."""
        on_dev = torch is not None and isinstance(x, torch.Tensor) and x.is_cuda
        B = self.batch
        pt, mem_t, kt = _buf(traj)
        px, mem, kx = _buf(x, writable=True)
        T = (kt.shape[0] if hasattr(kt, "shape") else len(kt))
        po = None
        if offset is not None:
            po, mem_o, ko = _buf(offset)
            if mem_o != mem:
                raise L.QspushError("offset and x must live in the same memory space")
        if mem_t != mem:
            raise L.QspushError("traj and x must live in the same memory space")
        lx = lu = ls = None
        plx = plu = pls = None
        if log:
            if on_dev:
                lx = torch.empty(steps, B, 4, dtype=torch.float64, device=x.device)
                lu = torch.empty(steps, B, 2, dtype=torch.float64, device=x.device)
                ls = torch.empty(steps, B, dtype=torch.int32, device=x.device)
                plx, plu, pls = lx.data_ptr(), lu.data_ptr(), ls.data_ptr()
            else:
                lx = np.zeros((steps, B, 4)); lu = np.zeros((steps, B, 2)); ls = np.zeros((steps, B), dtype=np.int32)
                plx, plu, pls = lx.ctypes.data, lu.ctypes.data, ls.ctypes.data
        lo = L.LoopOpts()
        lo.idx0 = int(idx0); lo.seed = int(seed); lo.t_dist = int(t_dist)
        lo.amplitude_dist = float(amplitude_dist); lo.xwidth = float(xwidth)
        lo.delay_plant = int(delay_plant); lo.delay_comp = int(delay_comp)
        for i in range(4):
            lo.noise_sigma[i] = float(noise_sigma[i])
        with self._Ordered(self, mem):
            L.check(L.lib().qspush_closed_loop(self._h, pt, int(T), po, px, int(steps), C.byref(lo), plx, plu, pls, mem))
        return dict(x=x, x_log=lx, u_log=lu, status_log=ls)

    def set_reference_trajectory(self, traj, offset=None):
        """NMPC_controller.set_reference_trajectory (NMPC_controller.m:425-431) for the batch: traj (T,6) columns
        [x_ref; u_ref], offset (batch,6) optional per-problem shift; kept on the device (qspush_set_reference_trajectory)."""
        pt, mem, kt = _buf(traj)
        po = None
        if offset is not None:
            po, mem_o, ko = _buf(offset)
            if mem_o != mem:
                raise L.QspushError("traj and offset must live in the same memory space")
            if tuple(ko.shape) != (self.batch, 6):
                raise L.QspushError("offset must be (batch, 6)")
        if len(kt.shape) != 2 or kt.shape[1] != 6:
            raise L.QspushError("traj must be (T, 6)")
        with self._Ordered(self, mem):
            L.check(L.lib().qspush_set_reference_trajectory(self._h, pt, int(kt.shape[0]), po, mem))

    def set_reference_window(self, idx):
        """Reference window of control period idx (1-based) into cost_y_ref / cost_y_ref_e of every stage
        (NMPC_controller.m:307-313, 343-348), on the device (qspush_set_reference_window)."""
        L.check(L.lib().qspush_set_reference_window(self._h, int(idx)))

    def step(self, x0, idx=1, u0=None, status=None, shift=False, restore_guess=False):
        """One control period at the controller boundary = NMPC_controller.solve(x0, index_time) for the batch as one CUDA
        graph launch (qspush_step): x0 (batch,4) in, u0 (batch,2) and status (batch) out, all host (numpy / pinned torch)
        or all torch CUDA tensors.  Returns (u0, status)."""
        px, mem, kx = _buf(x0)
        on_dev = mem == L.MEM_DEVICE
        if u0 is None:
            u0 = torch.empty(self.batch, 2, dtype=torch.float64, device=x0.device) if on_dev else np.zeros((self.batch, 2))
        if status is None:
            status = torch.empty(self.batch, dtype=torch.int32, device=x0.device) if on_dev else np.zeros(self.batch, dtype=np.int32)
        pu, mem_u, ku = _buf(u0, writable=True)
        ps, mem_s, ks = _buf(status, dtype=np.int32, writable=True)
        if mem_u != mem or mem_s != mem:
            raise L.QspushError("x0, u0 and status must live in the same memory space")
        flags = (L.STEP_SHIFT if shift else 0) | (L.STEP_RESTORE_GUESS if restore_guess else 0)
        with self._Ordered(self, mem):
            L.check(L.lib().qspush_step(self._h, px, int(idx), flags, pu, ps, mem))
        return u0, status

    def snapshot_guess(self):
        """Save the current input trajectory on the device for step(restore_guess=True)."""
        L.check(L.lib().qspush_snapshot_guess(self._h))

    def sync(self):
        L.check(L.lib().qspush_sync(self._h))

    @property
    def stream(self):
        return L.lib().qspush_stream(self._h)

    def stat(self, which):
        v = C.c_double()
        L.check(L.lib().qspush_get_stat(self._h, {"time_tot": 0, "time_lin": 1, "time_qp_sol": 2, "time_prep": 3}[which], C.byref(v)))
        return v.value

    @property
    def launches(self):
        return int(L.lib().qspush_launch_count(self._h))
