"""Synthetic workloads of BASELINE.json's configs (made concrete in SURVEY.md section 8d).

Pure numpy; shared by bench.py, the tests and smoke().  RNG = numpy PCG64 with the seed each config names.
Ranges come from the reference's own literals (main.m:53-56, NMPC_controller.m:23-26, 251-252).
"""
from __future__ import annotations

import json
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "objects.json")
OBJECT_ORDER = ("santal", "balea", "montana", "pulirapid")


def packaged_tables():
    with open(_DATA) as f:
        return json.load(f)


def packaged_model(name):
    """Product-side slider model (C-ABI handle) of one objects_database entry, built from the packaged outline tables."""
    from .capi import Model
    from .object_selection import OBJECT_TABLE
    t = packaged_tables()[name]
    return Model.from_tables(t["knots"], t["ctrl_xy"], 3, OBJECT_TABLE[name]["mu_sp"], t["c_ellipse"], True)


def reference_line(x0, N, dt, speed=0.01):
    """Straight-line reference x_ref = x0.x + speed*t, y_ref = x0.y, theta = s = 0, u_ref = 0 (config 1 / 3)."""
    x0 = np.atleast_2d(x0)
    B = x0.shape[0]
    t = np.arange(N) * dt
    yref = np.zeros((B, N, 6))
    yref[:, :, 0] = x0[:, 0:1] + speed * t[None, :]
    yref[:, :, 1] = x0[:, 1:2]
    return yref, np.ascontiguousarray(yref[:, N - 1, :4])


def make_rti_workload(batch, N, dt=0.05, seed=2, n_objects=1, mixed_modes=False):
    """Configs 3 / 4 / 5: random initial poses inside the feasible s-range, straight-line reference,
    initial guess u = [0.01; 0] (or mixed sticking/sliding inputs for config 5)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x0 = np.stack([rng.uniform(-0.03, 0.03, batch), rng.uniform(-0.03, 0.03, batch),
                   np.deg2rad(rng.uniform(-10.0, 10.0, batch)), rng.uniform(-0.04, 0.005, batch)], axis=1)
    yref, yref_e = reference_line(x0, N, dt)
    u = np.zeros((batch, N, 2))
    if mixed_modes:
        u[:, :, 0] = rng.uniform(0.002, 0.02, (batch, N))
        u[:, :, 1] = u[:, :, 0] * rng.uniform(-1.5, 1.5, (batch, N))
    else:
        u[:, :, 0] = 0.01
    obj = (np.arange(batch) % n_objects).astype(np.int32)
    return dict(x0=x0, yref=yref, yref_e=yref_e, u_init=u, object_id=obj)


def make_feasible_start_workload(batch, N, dt=0.05, seed=4, n_objects=1):
    """Config 5, feasible-start variant: the same horizon and solver settings, but a start from which the NLP is well posed —
    lateral / heading errors of a tracking controller (+-3 mm, +-1 deg instead of +-30 mm, +-10 deg), the contact point near the
    middle of the pushed edge (s in [-10, 5] mm) and an initial guess inside the friction cone (|u_t| <= 0.1 u_n, sticking).
    Meant for the symmetric outline (balea), where the no-rotation contact point is s = 0 and the minimiser lies inside the sticking
    mode; on the other three outlines the minimiser slides along the edge with u_n at its bound 0, i.e. ON the kink of the mode
    indicators, and full SQP chatters whatever the start."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x0 = np.stack([rng.uniform(-0.03, 0.03, batch), 0.1 * rng.uniform(-0.03, 0.03, batch),
                   0.1 * np.deg2rad(rng.uniform(-10.0, 10.0, batch)), rng.uniform(-0.01, 0.005, batch)], axis=1)
    yref, yref_e = reference_line(x0, N, dt)
    yref[:, :, 1] = 0.0; yref_e[:, 1] = 0.0                       # the reference line is y = 0: the lateral error is x0's
    u = np.zeros((batch, N, 2))
    u[:, :, 0] = rng.uniform(0.002, 0.02, (batch, N))
    u[:, :, 1] = u[:, :, 0] * rng.uniform(-0.1, 0.1, (batch, N))
    obj = (np.arange(batch) % n_objects).astype(np.int32)
    return dict(x0=x0, yref=yref, yref_e=yref_e, u_init=u, object_id=obj)


def make_samples_config2(b, n, seed=1, n_adversarial=4096, knots=None):
    """Config 2: n (x, u) samples populating all three contact modes plus adversarial corner cases."""
    rng = np.random.Generator(np.random.PCG64(seed))
    m = n - n_adversarial if n > 2 * n_adversarial else n
    x = np.stack([rng.uniform(-0.1, 0.1, m), rng.uniform(-0.1, 0.1, m), rng.uniform(-np.pi, np.pi, m), rng.uniform(-b, b, m)], axis=1)
    un = rng.uniform(1e-3, 0.03, m)
    ut = np.where(rng.random(m) < 0.5, rng.uniform(-0.05, 0.05, m), un * rng.uniform(-2.0, 2.0, m))
    u = np.stack([un, ut], axis=1)
    if m < n:
        k = n - m
        xa = np.stack([rng.uniform(-0.1, 0.1, k), rng.uniform(-0.1, 0.1, k), rng.uniform(-np.pi, np.pi, k), rng.uniform(-b, b, k)], axis=1)
        ua = np.stack([rng.uniform(1e-3, 0.03, k), rng.uniform(-0.05, 0.05, k)], axis=1)
        q = k // 4
        ua[:q, 0] = 0.0; ua[:q // 2, 1] = 0.0                       # u_n = 0 with u_t = 0 and u_t != 0
        special = [0.0, np.nextafter(b, 0.0), -np.nextafter(b, 0.0)]
        if knots is not None:
            special += [float(v) for v in np.unique(knots)[1:-1]]
        xa[q:2 * q, 3] = rng.choice(np.array(special), q)           # s on knots / at the seam
        x = np.concatenate([x, xa]); u = np.concatenate([u, ua])
    return np.ascontiguousarray(x), np.ascontiguousarray(u)
