"""GPU: the Gradient planner and the iLQS planner - C++ host classes through their C wrappers against the
Python mirrors driving the same ABI, plus the reference's behavioural criterion (particle reaches the goal,
mjpc/test/planners: sampling_planner_test.cc / ilqg_test.cc style)."""
import numpy as np
import pytest

from conftest import get_model, mocap_of

pytestmark = pytest.mark.gpu


def test_cpp_gradient_planner_matches_python_mirror():
    from mujoco_mpc_b200.engine import CppGradientPlanner, Engine
    from mujoco_mpc_b200.gradient import GradientPlanner
    for name, H in (("particle", 26), ("quadruped", 32)):
        m = get_model(name)
        state = np.concatenate([m.key_qpos[0] if m.nkey else m.qpos0, np.zeros(m.nv)])
        cpp = CppGradientPlanner(m, H, num_trajectory=8, num_spline_points=5, representation=1, fd_tolerance=1e-3, fd_mode=0)
        e = Engine(m, 8, H)
        py = GradientPlanner(m, e, horizon=H, num_trajectory=8, num_spline_points=5, representation=1, fd_tolerance=1e-3, fd_mode=0)
        cpp.reset(); cpp.set_state(state, 0.0, mocap_of(m)); py.set_state(state, 0.0, mocap_of(m))
        for it in range(4):
            ok_c = cpp.optimize_policy(); ok_p = py.optimize_policy()
            r = cpp.result()
            assert bool(ok_c) == bool(ok_p), (name, it)
            np.testing.assert_allclose(r["total_return"], py.total_return, rtol=1e-6)
            np.testing.assert_allclose(r["parameters"], py.parameters, atol=2e-6)
            np.testing.assert_allclose(r["times"], py.times, atol=1e-12)
            assert r["winner"] == py.winner
        a = cpp.action_from_policy(0.03)
        np.testing.assert_allclose(a, py.action_from_policy(0.03), atol=2e-6)
        cpp.close(); e.close()


def test_gradient_planner_descends_on_particle():
    from mujoco_mpc_b200.engine import CppGradientPlanner
    m = get_model("particle")
    pl = CppGradientPlanner(m, 26, num_trajectory=16, num_spline_points=6, representation=1, fd_tolerance=1e-3)
    pl.reset(); pl.set_state(np.zeros(4), 0.0, mocap_of(m))
    rets = []
    for _ in range(40):
        pl.optimize_policy(); rets.append(pl.result()["total_return"])
    assert rets[-1] < 0.9 * rets[0] and all(b <= a + 1e-7 for a, b in zip(rets, rets[1:]))   # monotone: step 0 is a candidate
    pl.close()


def test_cpp_ilqs_planner_matches_python_mirror():
    """Forces both branches: exploration 0 makes sampling fail to improve (-> iLQG iteration, then the trajectory ->
    spline conversion on the next call), exploration 0.1 lets sampling win again."""
    from mujoco_mpc_b200.engine import CppILQSPlanner, Engine
    from mujoco_mpc_b200.gradient import ILQSPlanner
    m = get_model("particle")
    H = 26
    cpp = CppILQSPlanner(m, H, num_trajectory=8, num_rollouts=6, fd_tolerance=1e-3, fd_mode=0)
    e1, e2 = Engine(m, 8, H), Engine(m, 8, H)
    py = ILQSPlanner(m, e1, e2, horizon=H, num_trajectory=8, num_rollouts=6, fd_tolerance=1e-3, fd_mode=0)
    cpp.reset(); cpp.set_state(np.zeros(4), 0.0, mocap_of(m)); py.set_state(np.zeros(4), 0.0, mocap_of(m))
    seq = []
    for it in range(10):
        sigma = 0.0 if 2 <= it < 7 else 0.1
        cpp.set_exploration(sigma); py.sampling.sigma = sigma
        cpp.optimize_policy(); py.optimize_policy()
        r = cpp.result()
        seq.append(r["active_policy"])
        assert r["active_policy"] == py.active_policy, (it, seq)
        assert r["sampling_winner"] == py.sampling.winner
        np.testing.assert_allclose(r["sampling_return"], float(py.sampling.returns[py.sampling.winner]), rtol=1e-5)
        np.testing.assert_allclose(r["ilqg_return"], py.ilqg.total_return, rtol=1e-5, atol=1e-9)
    assert 1 in seq and seq[0] == 0            # both policies were active at some point
    a = cpp.action_from_policy(0.05, np.zeros(4))
    np.testing.assert_allclose(a, py.action_from_policy(0.05, np.zeros(4)), atol=2e-5)
    cpp.close(); e1.close(); e2.close()


def test_ilqs_reaches_goal_on_particle():
    from mujoco_mpc_b200.engine import CppILQSPlanner
    m = get_model("particle")
    pl = CppILQSPlanner(m, 26, num_trajectory=16, num_rollouts=8, fd_tolerance=1e-3)
    pl.reset(); pl.set_state(np.zeros(4), 0.0, mocap_of(m))
    for _ in range(30):
        pl.optimize_policy()
    r = pl.result()
    assert min(r["sampling_return"], r["ilqg_return"] if r["ilqg_return"] > 0 else 1e9) < 0.05
    pl.close()


def test_agent_plan_iteration_glue():
    """Agent::PlanIteration (agent.cc:283-357) in C++: steps_ from horizon / timestep, timestep override, the task
    snapshot applied before planning, MakeDifferentiable only for gradient-based planners and restored afterwards."""
    from mujoco_mpc_b200.engine import CppAgent, CppSamplingPlanner, Engine, EngineError
    m = get_model("quadruped")
    state = np.concatenate([m.key_qpos[0], np.zeros(m.nv)])
    ag = CppAgent(m, "sampling", horizon=0.31, timestep=0.01, num_trajectory=16)
    assert ag.steps == 32
    ag.reset(); ag.set_state(state, 0.0, mocap_of(m))
    ag.plan_iteration(); ag.plan_iteration()
    direct = CppSamplingPlanner(m, 16, 32)
    direct.reset(); direct.set_state(state, 0.0, mocap_of(m))
    direct.optimize_policy(); r = direct.optimize_policy()
    np.testing.assert_allclose(ag.action_from_policy(0.02), direct.action_from_policy(0.02), atol=1e-12)
    # the residual snapshot is applied before planning: zero weights -> zero cost -> the nominal (candidate 0) always wins
    ag.set_task(weight=np.zeros(m.task_num_term))
    ag.plan_iteration()
    a0 = ag.action_from_policy(0.02)
    ag.plan_iteration()
    np.testing.assert_allclose(ag.action_from_policy(0.02), a0, atol=1e-12)
    ag.close(); direct.close()
    # timestep override reaches the kernels: half the step -> the same horizon in seconds needs twice the steps
    ag2 = CppAgent(m, "sampling", horizon=0.31, timestep=0.005, num_trajectory=8)
    assert ag2.steps == 63
    ag2.reset(); ag2.set_state(state, 0.0, mocap_of(m)); ag2.plan_iteration(); ag2.close()
    with pytest.raises(EngineError):
        CppAgent(m, "sampling", integrator=1)                      # only Euler is implemented: refused, not replaced
    # gradient-based planner: differentiable by default, and the engine is back to the plain model afterwards
    ag3 = CppAgent(m, "ilqg", horizon=0.15, timestep=0.01, ilqg_num_rollouts=6)
    ag3.reset(); ag3.set_state(state, 0.0, mocap_of(m))
    assert ag3.plan_iteration() >= 0
    a = ag3.action_from_policy(0.01, state)
    assert np.isfinite(a).all() and (np.abs(a) <= 1 + 1e-6).all()
    ag3.close()
    e = Engine(m, 4, 16)
    e.set_options(0.005)                                             # ABI-level check of the override
    q = state[: m.nq]; g1 = e.step_batch(q[None], np.zeros((1, m.nv)), np.zeros((1, m.nu)), mocap_of(m), [0.0])
    e.set_options(0.01); g2 = e.step_batch(q[None], np.zeros((1, m.nv)), np.zeros((1, m.nu)), mocap_of(m), [0.0])
    np.testing.assert_allclose(g1["next_qvel"] * 2, g2["next_qvel"], rtol=0.2, atol=1e-3)   # v' = dt * a (free fall at the home pose)
    e.close()
