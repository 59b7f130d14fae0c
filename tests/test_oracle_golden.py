"""CPU: pin the oracle against every known-answer test the reference holds for the hot path."""
import numpy as np
import pytest

from conftest import get_model, mocap_of
from mujoco_mpc_b200.blob import to_blob
from mujoco_mpc_b200 import task as T


def test_backward_pass_golden(oracle_lib):
    # mjpc/test/ilqg_planner/backward_pass_test.cc:29-140 (LQR n=2, m=1, T=3; lqr.cc:24-105)
    n, m, H = 2, 1, 3
    A = np.tile(np.array([[1.0, 1], [0, 1]]), (H, 1, 1)); B = np.tile(np.array([[0.0], [1.0]]), (H, 1, 1))
    u = np.full((H, 1), 0.5); x = np.zeros((H, 2))
    for t in range(H - 1):
        x[t + 1] = [x[t, 0] + x[t, 1], x[t, 1] + u[t, 0]]
    cx, cu = x.copy(), u.copy(); cu[H - 1] = 0
    cxx = np.tile(np.eye(2), (H, 1, 1)); cuu = np.ones((H, 1, 1)); cxu = np.zeros((H, 2, 1))
    o = oracle_lib.backward_pass(A, B, cx, cu, cxx, cxu, cuu, u, np.array([[-1.0, 1.0]]))
    assert o["status"] == 1
    np.testing.assert_allclose(o["Vx"].ravel(), [0.0, 0.0, 0.5, 1.25, 0.5, 1.0], atol=1e-5)
    np.testing.assert_allclose(o["Vxx"].ravel(), [2.71428571, 2.0, 2.0, 4.0, 2.0, 1.0, 1.0, 2.5, 1.0, 0.0, 0.0, 1.0], atol=1e-5)
    np.testing.assert_allclose(o["K"][:2].ravel(), [-0.285714285, -1.0, 0.0, -0.5], atol=1e-5)
    np.testing.assert_allclose(o["du"][:2].ravel(), [-0.5, -0.75], atol=1e-5)


def test_spline_golden(oracle_lib):
    # mjpc/test/spline/spline_test.cc:115-158
    s = oracle_lib.spline_sample
    np.testing.assert_allclose(s([1, 2], [[1.0, 2], [3, 4]], 0, 1.5), [1, 2])
    np.testing.assert_allclose(s([1, 2], [[1.0, 2], [3, 4]], 1, 1.5), [2, 3])
    np.testing.assert_allclose(s([1, 2], [[1.0, 2], [3, 4]], 2, 1.5), [2, 3])
    np.testing.assert_allclose(s([0, 1, 2, 3], [[1.0, 2], [1, 2], [3, 4], [3, 4]], 2, 1.5), [2, 3])
    for x in np.arange(0.0, 1.0001, 0.125):
        np.testing.assert_allclose(s([-1, 0, 1], [[1.0], [0.0], [1.0]], 2, x), [-x ** 3 + 2 * x ** 2], atol=1e-12)
    # outside the knot range the end knots are held; empty spline samples zero (spline.cc:108-123)
    np.testing.assert_allclose(s([1, 2], [[1.0, 2], [3, 4]], 2, 5.0), [3, 4])
    np.testing.assert_allclose(s([1, 2], [[1.0, 2], [3, 4]], 2, -5.0), [1, 2])


def test_host_spline_matches_oracle(oracle_lib):
    from mujoco_mpc_b200.planner import sample_spline
    rng = np.random.default_rng(1)
    times = np.cumsum(rng.uniform(0.1, 0.5, 6)); vals = rng.normal(size=(6, 3))
    for interp in (0, 1, 2):
        for t in np.linspace(times[0] - 0.3, times[-1] + 0.3, 41):
            np.testing.assert_allclose(sample_spline(times, vals, interp, t), oracle_lib.spline_sample(times, vals, interp, t), atol=1e-12)


@pytest.mark.parametrize("ntype,params", [(0, []), (1, [0.1, 2.0]), (2, [0.1]), (3, [0.2]), (5, [2.5]), (6, [0.1]),
                                          (7, [0.1, 2.0]), (8, [0.3])])
def test_norm_gradients(oracle_lib, ntype, params):
    # mjpc/test/agent/norm_test.cc:42-109: analytic gradient / Hessian vs finite differences
    rng = np.random.default_rng(ntype)
    for _ in range(5):
        x = rng.uniform(0.2, 1.0, 3) * rng.choice([-1, 1], 3)
        y, g, H = oracle_lib.norm(x, np.array(params, float), ntype, grad=True, hess=True)
        eps = 1e-6
        gfd = np.zeros(3); Hfd = np.zeros((3, 3))
        for i in range(3):
            e = np.zeros(3); e[i] = eps
            yp = oracle_lib.norm(x + e, np.array(params, float), ntype)[0]
            ym = oracle_lib.norm(x - e, np.array(params, float), ntype)[0]
            gfd[i] = (yp - ym) / (2 * eps)
            gp = oracle_lib.norm(x + e, np.array(params, float), ntype, grad=True)[1]
            gm = oracle_lib.norm(x - e, np.array(params, float), ntype, grad=True)[1]
            Hfd[:, i] = (gp - gm) / (2 * eps)
        np.testing.assert_allclose(g, gfd, rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(H, Hfd, rtol=1e-2, atol=1e-5)


def test_task_cost_golden(oracle_lib):
    # mjpc/test/tasks/task_test.cc:49-99
    m = get_model("particle")
    assert abs(m.task_risk - 1.0) < 1e-5 and len(m.task_parameters) == 2
    np.testing.assert_allclose(m.task_parameters, [0.05, -0.1])
    assert m.task_num_residual == 4 and m.task_num_term == 2
    assert list(m.task_dim_norm_residual) == [2, 2] and list(m.task_num_norm_parameter) == [0, 0]
    assert list(m.task_norm) == [T.NORM_QUADRATIC] * 2
    np.testing.assert_allclose(m.task_weight, [5.0, 0.1])
    o = oracle_lib.Oracle(to_blob(m), m, 64)
    r = np.array([1e-3, 2e-3, 3e-3, 4e-3])
    c = 5.0 * 0.5 * r[:2] @ r[:2] + 0.1 * 0.5 * r[2:] @ r[2:]
    o.set_task(risk=0.0)
    v, terms = o.cost_value(r, terms=True)
    assert abs(terms.sum() - c) < 1e-5 and abs(v - c) < 1e-12
    o.set_task(risk=0.2)
    assert abs(o.cost_value(r) - (np.exp(0.2 * c) - 1) / 0.2) < 1e-5


def test_particle_rollout_golden(oracle_lib):
    # mjpc/test/agent/rollout_test.cc:67-153: PD policy reaches (0.1, 0.1); residual == states (L1 < 1e-5)
    m = get_model("particle_copy", agent_timestep=False)   # the test steps the model at its own dt = 0.01
    assert abs(m.opt_timestep - 0.01) < 1e-15
    o = oracle_lib.Oracle(to_blob(m), m, 64)
    H = 100
    K = np.zeros((H, 2, 4)); K[:, 0, 0] = K[:, 1, 1] = -10.0; K[:, 0, 2] = K[:, 1, 3] = -2.5
    xn = np.zeros((H, 4)); xn[:, :2] = 0.1
    r = o.rollout_feedback(np.zeros(4), 0.0, mocap_of(m), np.zeros((H, 2)), xn, np.arange(H) * 0.01, K, np.zeros((H, 2)), [1.0], 3)
    assert r["failure"][0] == 0
    assert np.abs(r["states"][0, -1, :2] - 0.1).sum() < 0.1 and np.abs(r["states"][0, -1, 2:]).sum() < 0.1
    assert np.abs(r["states"][0] - r["residual"][0]).sum() < 1e-5
    # closed form of the linear particle (mass .3, damping 1, implicit-damped Euler): one step from rest, u=(1,0)
    r1 = o.forward_debug(np.zeros(2), np.zeros(2), np.array([1.0, 0.0]), mocap_of(m))
    np.testing.assert_allclose(r1["next_qvel"], [0.01 * 1.0 / (0.3 + 0.01 * 1.0), 0.0], rtol=1e-12, atol=1e-15)


def test_mass_matrix_and_gravity_cross_check(oracle_lib, quadruped):
    # independent formulation (world-frame Jacobians) vs the oracle's composite-rigid-body / RNE recursions
    from mujoco_mpc_b200 import refmath
    m = quadruped
    o = oracle_lib.Oracle(to_blob(m), m, 64)
    rng = np.random.default_rng(0)
    q = m.key_qpos[0].copy()
    q[3:7] = rng.normal(size=4); q[3:7] /= np.linalg.norm(q[3:7]); q[7:] += rng.normal(size=12) * 0.3; q[2] = 1.0
    r = o.forward_debug(q, np.zeros(m.nv), np.zeros(m.nu), mocap_of(m))
    M, Jb = refmath.mass_matrix_and_jacobians(m, q)
    assert np.abs(M - r["qM"]).max() < 1e-12
    G = np.zeros(m.nv)
    for b in range(1, m.nbody):
        G -= m.body_mass[b] * Jb[b][0:3].T @ np.asarray(m.opt_gravity)
    assert np.abs(G - r["qfrc_bias"]).max() < 1e-12


def test_free_fall_and_energy(oracle_lib, quadruped):
    m = quadruped
    o = oracle_lib.Oracle(to_blob(m), m, 64)
    q = m.key_qpos[0].copy(); q[2] = 2.0
    r = o.forward_debug(q, np.zeros(m.nv), np.zeros(m.nu), mocap_of(m))
    assert r["ncon"] == 0
    np.testing.assert_allclose(r["qacc"][:3], [0, 0, -9.81], atol=1e-9)


def test_contact_solution_is_a_minimum(oracle_lib, quadruped):
    """Solver invariant: normal forces non-negative, friction inside the cone, robot weight carried."""
    m = quadruped
    o = oracle_lib.Oracle(to_blob(m), m, 64)
    q = m.key_qpos[0].copy(); q[2] = 0.245
    r = o.forward_debug(q, np.zeros(m.nv), np.zeros(m.nu), mocap_of(m))
    assert r["ncon"] >= 4
    dims = r["contact"][:, 7].astype(int)
    assert r["nefc"] == 12 + dims.sum()
    a = 12
    for d in dims:
        f = r["efc_force"][a:a + d]
        assert f[0] >= 0
        if d >= 3:
            assert np.hypot(f[1], f[2]) <= 1.0 * f[0] + 1e-9   # mu <= 1 for every pair of this model
        a += d
    # constraint force is J^T f and pushes the trunk up
    assert r["qfrc_constraint"][2] > 0


def test_fp32_oracle_tracks_fp64(oracle_lib, quadruped):
    from conftest import quadruped_inputs
    m = quadruped
    b = to_blob(m)
    o64, o32 = oracle_lib.Oracle(b, m, 64), oracle_lib.Oracle(b, m, 32)
    state, mocap, knots, kt = quadruped_inputs(m, N=4, H=16)
    r64 = o64.rollout_spline(state, 0.0, mocap, knots, kt, 2, 16, full=False)
    r32 = o32.rollout_spline(state, 0.0, mocap, knots, kt, 2, 16, full=False)
    np.testing.assert_allclose(r32["returns"], r64["returns"], rtol=2e-4)


def test_interpolation_helpers_golden(oracle_lib):
    """mjpc/test/agent/agent_utilities_test.cc:237-283 (FindInterval, LinearInterpolation) + the zero-order and cubic
    variants used by iLQGPolicy::Action (utilities.cc:303-422): endpoints clamp, interior is the usual formula."""
    seq = [-1.0, 0.0, 1.0, 2.0]
    assert oracle_lib.find_interval(seq, 0.5) == (1, 2)
    assert oracle_lib.find_interval(seq, -2.0) == (0, 0)
    assert oracle_lib.find_interval(seq, 2.1) == (3, 3)
    x, y = [1.0, 2.0], np.array([[1.0], [2.0]])
    assert abs(oracle_lib.interpolate(1.5, x, y, 1)[0] - 1.5) < 1e-12
    assert abs(oracle_lib.interpolate(0.5, x, y, 1)[0] - 1.0) < 1e-12      # lower extrapolation clamps
    assert abs(oracle_lib.interpolate(2.5, x, y, 1)[0] - 2.0) < 1e-12      # upper extrapolation clamps
    assert abs(oracle_lib.interpolate(1.5, x, y, 0)[0] - 1.0) < 1e-12      # zero-order hold
    # cubic Hermite with finite-difference slopes reproduces a straight line exactly, and the nodes
    xs = np.array([0.0, 0.3, 0.7, 1.0, 1.6]); ys = (2.0 * xs - 1.0)[:, None]
    for q in (0.1, 0.5, 0.9, 1.3):
        assert abs(oracle_lib.interpolate(q, xs, ys, 2)[0] - (2 * q - 1)) < 1e-12
    ys2 = np.sin(xs)[:, None]
    for k, q in enumerate(xs[:-1]):
        assert abs(oracle_lib.interpolate(q, xs, ys2, 2)[0] - ys2[k, 0]) < 1e-12
