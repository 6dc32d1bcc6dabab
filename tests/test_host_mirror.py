"""Host-side mirror of the reference's MATLAB interface: pieces that need no GPU."""
import numpy as np
import pytest

import uclv_qs_pushing_matlab_b200 as q
from uclv_qs_pushing_matlab_b200 import sharding


def test_object_selection_table():
    s = q.object_selection("santal")                                   # object_selection.m:3-12
    assert (s.mu_sg, s.mu_sp, s.m, s.tau_max) == (0.32, 0.19, 0.2875, 0.0251)
    assert abs(s.area - 0.068 * 0.082) < 1e-15 and s.pcl_path == "planar_surface_santal_36_uniformed.ply"
    assert q.object_selection("pulirapid").mu_sp == 0.1
    with pytest.raises(ValueError):
        q.object_selection("nope")


def test_trajectory_generator():
    tg = q.TrajectoryGenerator(0.05, 0.01)
    tg.set_target([0, 0, 0, 0, 0], [0.3, 0.03, 0.1, 0, 0.07], 0, 10)
    t, tr = tg.straight_line(False)                                     # TrajectoryGenerator.m:44-52
    assert tr.shape == (5, 201) and np.allclose(tr[:, 0], 0) and np.allclose(tr[:, -1], [0.3, 0.03, 0.1, 0, 0.07])
    assert np.all(np.diff(tr[0]) >= -1e-15)
    tg.waypoints_ = np.array([[0, 0, 0], [0.10, 0, 0]]); tg.waypoints_velocities = [0.01]
    t, tr = tg.waypoint_gen_fixed_angle()                               # :81-95
    assert tr.shape[0] == 5 and abs(tr[0, -1] - 0.10) < 1e-12 and abs(t[-1] - 10.0) < 1e-9
    t, tr = tg.waypoints_gen()                                          # documented stand-in for :96-143
    assert tr.shape == (4, 201) and abs(tr[0, 100] - 0.05) < 1e-12 and np.allclose(tr[3], 0.01)


def test_shard_ranges_cover_the_batch():
    for total in (1, 7, 4096, 65536):
        for w in (1, 2, 3, 8):
            r = [sharding.shard_range(total, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == total and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    perm, ranges = sharding.bucket_by_object(np.arange(20) % 4, 2)
    assert list((np.arange(20) % 4)[perm]) == sorted(np.arange(20) % 4) and ranges == [(0, 10), (10, 20)]


def test_acados_bags_and_controller_construction_need_no_solver():
    from uclv_qs_pushing_matlab_b200.acados_shim import acados_ocp_model, acados_ocp_opts
    m = acados_ocp_model(); m.set("T", 0.5); assert m.get("T") == 0.5
    o = acados_ocp_opts(); o.set("nlp_solver", "sqp"); assert o.get("nlp_solver") == "sqp"

    class _Plant:            # minimal stand-in carrying what NMPC_controller reads at construction time
        name = "real_plant"; sym_model = {"nx": 4, "nu": 2}
    c = q.NMPC_controller("NMPC", _Plant(), 0.05, 10)
    assert c.T == 0.5 and c.v_alpha == 1.0 and c.t_angle0 == 3 and c.d_v_bound == 0.0
    assert np.allclose(np.diag(c.W_x), [1, 1, 1e-3, 0]) and np.allclose(np.diag(c.W_x_e), [2e5, 2e5, 20, 0])
    mdl = c.create_ocp_model()
    assert mdl.get("constr_lh") == [-0.06, 0.0, -0.05] and mdl.get("constr_uh") == [0.011, 0.03, 0.05]
    opts = c.create_ocp_opts()
    assert opts.get("nlp_solver") == "sqp" and opts.get("nlp_solver_max_iter") == 30 and opts.get("qp_solver_cond_N") == 5
    c.set_delay_comp(0.12)
    assert c.delay_buff_comp == 3 and c.u_buff_contr.shape == (2, 3)
    c.set_reference_trajectory(np.arange(12.0).reshape(6, 2))
    assert c.y_ref.shape == (6, 5) and c.y_ref[5, 0] == c.y_ref[5, 3]
    assert np.array_equal(c.get_y_ref(99), c.y_ref[:, -1]) and np.array_equal(c.get_y_ref(1), c.y_ref[:, 0])
