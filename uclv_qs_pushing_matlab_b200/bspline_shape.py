"""acados_nmpc/bspline_shape.m — clamped B-spline outline of the slider, evaluated on the GPU.

The reference builds CasADi Functions of the curvilinear abscissa s (FC, FC_dot, FC_dot_dot, t_fun,
n_fun, R_NT_fun, FC_angle_dot); here each of them is a small callable that evaluates batches of s
through qspush_eval_spline (kernel k_eval_spline).  Method names and semantics follow the reference.
"""
from __future__ import annotations

import numpy as np

from .capi import Model


class _SplineFn:
    """Stand-in for a casadi.Function of s: F(s) -> array with one row per s."""

    def __init__(self, owner, kind):
        self._o, self.kind = owner, kind

    def __call__(self, s):
        s = np.atleast_1d(np.asarray(s, dtype=np.float64))
        return self._o._eval(self.kind, s, wrap=0)


class bspline_shape:
    # bspline_shape.m:25-38
    def __init__(self, S, P, p, _model: Model | None = None, device: int = 0):
        self.p = int(p)
        self.S = np.asarray(S, dtype=np.float64).reshape(-1)
        self.P = np.asarray(P, dtype=np.float64).reshape(-1, 2)
        self.n = self.P.shape[0]            # length(P)
        self.m = self.S.shape[0]            # length(S)
        self.a = 0.0
        self.device = device
        self._model = _model if _model is not None else Model.from_tables(self.S, self.P, self.p, 0.0, 1.0, True)
        self.b = self._model.b              # sum(vecnorm(diff(P)')) == last knot
        self.cj_1_vect = self._model.cj_1_vect
        self.max_curvature = None
        # "Functions" (bspline_shape.m:82,103,113-115,134,143)
        self.FC = _SplineFn(self, "C")
        self.FC_dot = _SplineFn(self, "Cd")
        self.FC_dot_dot = _SplineFn(self, "Cdd")
        self.t_fun = _SplineFn(self, "t")
        self.n_fun = _SplineFn(self, "n")
        self.R_NT_fun = _SplineFn(self, "R_NT")
        self.FC_angle_dot = _SplineFn(self, "kappa")

    # the get* builders of the reference only construct the CasADi graphs; the tables already exist here
    def getSymbolicSpline(self, ord=None): return None          # noqa: E704  bspline_shape.m:74-83
    def getSymboliSplineDot(self, ord=None): return None        # noqa: E704  :85-104
    def getSymbolicSplineDotDot(self, ord=None): return None    # noqa: E704  :118-135
    def getNormalTangentialVersors(self): return None           # noqa: E704  :106-116
    def getSymbolicAngleCurvatures(self): return None           # noqa: E704  :137-144

    def _eval(self, kind, s, wrap):
        if kind == "R_NT":
            r = self._model.eval_spline(s, wrap=wrap, device=self.device, want=("t", "n"))
            R = np.zeros((len(s), 2, 2))
            R[:, :, 0] = r["n"]             # R_NT = [nvers' tvers']   (:111)
            R[:, :, 1] = r["t"]
            return R
        return self._model.eval_spline(s, wrap=wrap, device=self.device, want=(kind,))[kind]

    # bspline_shape.m:192-199
    def evalSpline(self, F, s_values):
        s = np.atleast_1d(np.asarray(s_values, dtype=np.float64))
        return self._eval(F.kind, s, wrap=1)

    # bspline_shape.m:146-152
    def getAngleCurvatures(self, s_values):
        s = np.atleast_1d(np.asarray(s_values, dtype=np.float64))
        return self._eval("kappa", s, wrap=1)

    # bspline_shape.m:154-179 (seam-blended |C''|)
    def getCurvatures(self, s_values):
        s = np.atleast_1d(np.asarray(s_values, dtype=np.float64))
        delta_01 = delta_0n = 0.011
        s1, s0 = self.a + delta_01, self.a - delta_0n
        sn, sn_1 = self.b + delta_01, self.b - delta_0n
        # mod first (as :155); the kernel applies the same (single-precision) MATLAB mod again, idempotent
        sm = self._mod(s)
        pts = np.concatenate([sm, [s1, s0, sn, sn_1]])
        nrm = np.linalg.norm(self._eval("Cdd", pts, wrap=1), axis=1)
        y1, y0, yn, yn_1 = nrm[-4:]
        cur = nrm[:-4].copy()
        head = (sm <= s1) & (sm >= s0)
        tail = (sm <= sn) & (sm >= sn_1) & ~head
        cur[head] = (y1 - y0) * (sm[head] - s0) / (s1 - s0) + y0
        cur[tail] = (yn - yn_1) * (sm[tail] - sn_1) / (sn - sn_1) + yn_1
        return cur

    def _mod(self, s):
        # MATLAB mod(double, single) -> single (b is a `single` because pcread returns single); the builtin's algorithm:
        # r = fmod(x, y), 0 when x / y is an integer within eps |x / y|, else r += y when the signs differ (may round to y)
        x, y = s.astype(np.float32), np.float32(self.b)
        with np.errstate(invalid="ignore", divide="ignore"):
            r = np.fmod(x, y)
            q = np.abs(x / y)
            req0 = (r == 0) | ~(np.abs(q - np.floor(q + np.float32(0.5))) > np.float32(1.1920929e-7) * q)
            r = np.where(req0, np.float32(0.0), np.where((x < 0) != (y < 0), r + y, r)).astype(np.float32)
        return r.astype(np.float64)

    # bspline_shape.m:181-185
    def getMaxCurvature(self):
        s_values = np.arange(self.a, self.b + 1e-12, 0.001)
        self.max_curvature = float(np.max(self.getCurvatures(s_values)))
        return self.max_curvature

    # bspline_shape.m:187-190
    def getNormalizedCurvature(self, s):
        if self.max_curvature is None:
            self.getMaxCurvature()
        return self.getCurvatures(s) / self.max_curvature
