"""GPU parity (through the C-ABI) of the stateless kernels: spline (K1), dynamics and ERK4 + forward
sensitivities (K2) vs the CPU oracle, the committed golden fixtures, and size-independent properties at
BASELINE.json's config-2 size (1M samples).  Tolerance: 1e-10 relative (north_star)."""
import os

import numpy as np
import pytest

from tests.test_hostsim_parity import REL, rel_err
from tests.workloads import OBJECT_ORDER, make_samples_config2, packaged_model_pair

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", OBJECT_ORDER)
def test_eval_kernels_vs_oracle(name):
    gm, om = packaged_model_pair(name)
    x, u = make_samples_config2(om.b, 20480, seed=1, n_adversarial=4096, knots=om.S)
    for wrap in (0, 1, 2):
        a, b = om.eval_spline(x[:, 3], wrap=wrap), gm.eval_spline(x[:, 3], wrap=wrap)
        for k in ("C", "Cd", "Cdd", "t", "n", "kappa"):
            assert rel_err(b[k], a[k]) < REL, (wrap, k)
    fo, Jxo, Juo = om.dynamics(x, u, jac=True)
    fg, Jxg, Jug = gm.eval_dynamics(x, u, jac=True)
    assert rel_err(fg, fo) < REL and rel_err(Jxg, Jxo) < REL and rel_err(Jug, Juo) < REL
    assert rel_err(gm.eval_dynamics(x, u), fo) < REL
    Po, Ao, Bo = om.erk4_sens(x, u, 0.05, nthreads=8)
    Pg, Ag, Bg = gm.eval_erk4_sens(x, u, 0.05)
    assert rel_err(Pg, Po) < REL and rel_err(Ag, Ao) < REL and rel_err(Bg, Bo) < REL
    vo = np.array([om.v_bound(s) for s in x[:1024, 3]])
    vg, tg = gm.eval_v_bound(x[:1024, 3])
    assert rel_err(vg, vo[:, 0]) < REL and rel_err(tg, vo[:, 1]) < 1e-9


@pytest.mark.parametrize("name", OBJECT_ORDER)
def test_eval_kernels_vs_golden(name):
    gm, _ = packaged_model_pair(name)
    g = np.load(os.path.join(GOLD, f"eval_{name}.npz"))
    sp = gm.eval_spline(g["x"][:, 3], wrap=2)
    for k in ("C", "Cd", "Cdd", "t", "n", "kappa"):
        assert rel_err(sp[k], g[k]) < REL, k
    f, Jx, Ju = gm.eval_dynamics(g["x"], g["u"], jac=True)
    Phi, A, B = gm.eval_erk4_sens(g["x"], g["u"], 0.05)
    for a, b, nm in ((f, g["f"], "f"), (Jx, g["Jx"], "Jx"), (Ju, g["Ju"], "Ju"), (Phi, g["Phi"], "Phi"), (A, g["A"], "A"), (B, g["B"], "B")):
        assert rel_err(a, b) < REL, nm
    vb, ta = gm.eval_v_bound(g["x"][:256, 3])
    assert rel_err(vb, g["v_bound"]) < REL and rel_err(ta, g["t_angle"]) < 1e-9


def test_ieee_corner_cases_on_device():
    gm, om = packaged_model_pair("santal")
    b = om.b
    x = np.array([[0, 0, 0.3, -0.01], [0, 0, 0.3, -0.01], [0, 0, 0.3, -b], [0, 0, 0.3, -1e-20], [0.1, 0.2, 0.3, 0.0]])
    u = np.array([[0.0, 0.0], [0.0, 0.02], [0.01, 0.0], [0.01, 0.0], [0.01, 0.001]])
    f, Jx, Ju = gm.eval_dynamics(x, u, jac=True)
    assert np.all(f[0] == 0) and np.all(Jx[0] == 0) and np.all(Ju[0] == 0)
    assert np.array_equal(f[1], [0, 0, 0, 0.02])
    assert np.all(np.isnan(f[2, :3])) and np.all(np.isnan(f[3, :3]))
    Phi, A, B = gm.eval_erk4_sens(x[:1], u[:1], 0.05)
    assert np.array_equal(A[0], np.eye(4)) and np.all(B == 0) and np.array_equal(Phi[0], x[0])
    assert gm.eval_spline(np.zeros(0))["C"].shape == (0, 2)       # empty input


def test_config2_full_size_properties():
    """1M (x,u) samples, santal (BASELINE config 2): size-independent properties + oracle on a strided subset."""
    import torch
    gm, om = packaged_model_pair("santal")
    n = 1 << 20
    x, u = make_samples_config2(om.b, n, seed=1, knots=om.S)
    dev = torch.device("cuda:0")
    xd, ud = torch.from_numpy(x).to(dev), torch.from_numpy(u).to(dev)
    Phi = torch.empty(n, 4, dtype=torch.float64, device=dev); A = torch.empty(n, 4, 4, dtype=torch.float64, device=dev)
    B = torch.empty(n, 4, 2, dtype=torch.float64, device=dev)
    gm.eval_erk4_sens_device(xd, ud, 0.05, Phi, A, B)
    torch.cuda.synchronize()
    Phi, A, B = Phi.cpu().numpy(), A.cpu().numpy(), B.cpu().numpy()
    fin = np.isfinite(Phi).all(1)
    assert fin.mean() > 0.999                                     # only the sigma == b corner cases are NaN
    assert np.array_equal(A[fin][:, :, :2], np.broadcast_to(np.eye(4)[:, :2], (fin.sum(), 4, 2)))   # df/dx = df/dy = 0
    stick = (B[:, 3, 0] == 0) & (B[:, 3, 1] == 0) & fin
    assert 0.1 < stick.mean() < 0.9 and np.array_equal(Phi[stick][:, 3], x[stick][:, 3])           # s_dot = 0 when sticking
    # translation invariance: f does not depend on (x, y)
    x2 = x.copy(); x2[:, :2] += 0.5
    Phi2 = torch.empty_like(torch.from_numpy(Phi)).to(dev)
    gm.eval_erk4_sens_device(torch.from_numpy(x2).to(dev), ud, 0.05, Phi2, None, None)
    torch.cuda.synchronize()
    d = Phi2.cpu().numpy()[fin] - Phi[fin]
    assert np.abs(d[:, :2] - 0.5).max() < 1e-15 + 1e-16 and np.array_equal(d[:, 2:], np.zeros_like(d[:, 2:]))
    idx = np.arange(0, n, 53)
    Po, Ao, Bo = om.erk4_sens(x[idx], u[idx], 0.05, nthreads=8)
    assert rel_err(Phi[idx], Po) < REL and rel_err(A[idx], Ao) < REL and rel_err(B[idx], Bo) < REL
