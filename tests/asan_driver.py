"""Driver executed under AddressSanitizer by tests/test_hostsim_asan.py: every kernel body (both QP kernels, all
chunk sizes of the warp kernel, full SQP, multi-object batches, ragged sizes) on small problems."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import tests.hostsim.hostsim as hs  # noqa: E402

hs._LIB = sys.argv[1]
hs.subprocess.check_call = lambda *a, **k: 0          # use the ASan build as is
from tests.workloads import OBJECT_ORDER, hostsim_model, make_rti_workload  # noqa: E402

mhs = [hostsim_model(n) for n in OBJECT_ORDER]
for N, B, kern, mode in ((40, 7, 1, "rti"), (10, 5, 1, "rti"), (100, 3, 1, "rti"), (40, 33, 0, "rti"), (10, 4, 1, "sqp"),
                         (64, 3, 1, "rti"), (31, 3, 1, "rti"), (95, 2, 1, "rti"), (1, 2, 1, "rti"), (130, 2, 1, "rti"),
                         (47, 3, 1, "rti"), (55, 3, 1, "rti"), (63, 2, 1, "rti"), (20, 5, 1, "sqp")):
    wl = make_rti_workload(None, batch=B, N=N, seed=1, n_objects=4)
    r = hs.solve(mhs, N, 0.05, wl["x0"], wl["yref"], wl["yref_e"], np.zeros((B, N + 1, 4)), wl["u_init"], objid=wl["object_id"],
                 mode=mode, prepare=True, shift=True, qp_kernel=kern, max_sqp_iter=3)
    assert np.isfinite(r["u"]).all()
for N, B, mode in ((40, 5, "rti"), (70, 3, "rti"), (100, 2, "rti"), (10, 4, "sqp")):   # parked constraint set h = [u_n; u_t -+ v_bound(s)]
    wl = make_rti_workload(None, batch=B, N=N, seed=3, n_objects=4)
    r = hs.solve(mhs, N, 0.05, wl["x0"], wl["yref"], wl["yref_e"], np.zeros((B, N + 1, 4)), wl["u_init"], objid=wl["object_id"],
                 mode=mode, prepare=True, shift=True, qp_kernel=1, max_sqp_iter=3, h_variant=1, lh=(0.0, -0.1, 0.0), uh=(0.03, 0.0, 0.1))
    assert np.isfinite(r["u"]).all()
T = 60; traj = np.zeros((T, 6)); traj[:, 0] = 0.0005 * np.arange(T)                      # device-resident closed loop bodies
r = hs.closed_loop(mhs, 10, 0.05, traj, np.zeros((5, 4)), 8, offset=np.zeros((5, 6)), objid=np.arange(5) % 4,
                   t_dist=3, amplitude_dist=0.003, xwidth=0.068, noise_sigma=(1e-5, 1e-5, 1e-3, 1e-4), seed=3)
assert np.isfinite(r["x"]).all()
x = np.random.default_rng(0).uniform(-0.7, 0.7, (500, 4)); u = np.random.default_rng(1).uniform(-0.05, 0.05, (500, 2))
mhs[3].eval_spline(x[:, 3], wrap=2); mhs[3].dynamics(x, u); mhs[3].erk4_sens(x, u, 0.05); mhs[3].v_bound(x[:, 3])
print("ASAN-DRIVER-OK")
