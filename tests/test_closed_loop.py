"""GPU: closed-loop MPC, testspeed-style (mjpc/testspeed.cc:71-123): the planner runs on the device, the simulated
plant is the fp64 oracle (test infrastructure), one planning iteration per plant step.  The reference's own quality
metric is the average cost along the closed-loop trajectory; here it is compared between the device planner and
the same planner on the CPU oracle backend, both fed the same injected noise.

Task::Transition is not needed for this scenario: Quadruped mode, fixed Stand gait (auto gait switching only acts
above 0.02 m/s and only changes weights through Transition), goal mocap fixed at (0.3, 0, 0.26)."""
import numpy as np
import pytest

from conftest import OracleBackend, get_model, mocap_of

pytestmark = pytest.mark.gpu


def _closed_loop(m, backend, plant, steps, N, H):
    from mujoco_mpc_b200.planner import SamplingPlanner
    pl = SamplingPlanner(m, backend, num_trajectory=N, horizon=H)
    pl.reset(np.zeros(m.nu))
    q = m.key_qpos[0].copy(); v = np.zeros(m.nv)
    mocap = mocap_of(m)
    t, costs, heights = 0.0, [], []
    warm = None
    for k in range(steps):
        pl.set_state(np.concatenate([q, v]), t, mocap)
        pl.optimize_policy()
        u = pl.action_from_policy(t)
        r = plant.forward_debug(q, v, u, mocap, time=t, warmstart=warm)
        costs.append(plant.cost_value(r["residual"][: m.task_num_residual]))
        warm = r["qacc"]
        q, v = r["next_qpos"], r["next_qvel"]
        t += m.opt_timestep
        heights.append(q[2])
    return np.array(costs), np.array(heights), q


def test_quadruped_closed_loop_cost_matches_cpu_planner(oracle_lib):
    from mujoco_mpc_b200.blob import to_blob
    from mujoco_mpc_b200.engine import Engine
    m = get_model("quadruped")
    plant = oracle_lib.Oracle(to_blob(m), m, 64)
    steps, N, H = 60, 64, 32
    e = Engine(m, N, H)
    c_gpu, h_gpu, q_gpu = _closed_loop(m, e, plant, steps, N, H)
    c_cpu, h_cpu, q_cpu = _closed_loop(m, OracleBackend(m, threads=16), plant, steps, N, H)
    e.close()
    print("closed-loop average cost: device planner %.4f, CPU-oracle planner %.4f; final x %.3f / %.3f" % (
        c_gpu.mean(), c_cpu.mean(), q_gpu[0], q_cpu[0]))
    assert np.isfinite(c_gpu).all() and (h_gpu > 0.15).all() and (h_cpu > 0.15).all()     # never falls
    # same injected noise, same plant: the two loops follow each other until fp32/fp64 ranking ties split them;
    # the quality metric must agree closely and the first steps exactly
    assert abs(c_gpu.mean() - c_cpu.mean()) < 0.1 * c_cpu.mean()
    np.testing.assert_allclose(c_gpu[:5], c_cpu[:5], rtol=2e-3)
    # planning beats not planning: zero control for the same duration costs more
    q, v, cz = m.key_qpos[0].copy(), np.zeros(m.nv), []
    warm = None
    for k in range(steps):
        r = plant.forward_debug(q, v, np.zeros(m.nu), mocap_of(m), time=k * m.opt_timestep, warmstart=warm)
        cz.append(plant.cost_value(r["residual"][: m.task_num_residual])); warm = r["qacc"]
        q, v = r["next_qpos"], r["next_qvel"]
    assert c_gpu.mean() < np.mean(cz)


def test_humanoid_track_closed_loop(oracle_lib):
    """BASELINE config 3 loop at a small size: Tracking::TransitionLocked (host) moves the mocap markers and owns
    (mode, reference_time); every plan iteration snapshots that task state into the engine (mjpc_b200_set_task), the
    device plans, the oracle plant steps.  The tracked markers must stay close to the clip, and better than with
    zero control."""
    from mujoco_mpc_b200.blob import to_blob
    from mujoco_mpc_b200.engine import Engine
    from mujoco_mpc_b200.planner import SamplingPlanner
    from mujoco_mpc_b200.transition import HumanoidTrackTransition
    m = get_model("humanoid_track")
    plant = oracle_lib.Oracle(to_blob(m), m, 64)
    N, H, steps = 64, 41, 80

    def run(plan):
        e = Engine(m, N, H) if plan else None
        pl = SamplingPlanner(m, e, num_trajectory=N, horizon=H) if plan else None
        if plan:
            pl.reset(np.zeros(m.nu))
        tr = HumanoidTrackTransition(m)
        q, v, t, warm, errs, costs = m.qpos0.copy(), np.zeros(m.nv), 0.0, None, [], []
        for k in range(steps):
            q, v, mocap = tr.transition(t, q, v)
            plant.set_task(task_state=tr.task_state())
            u = np.zeros(m.nu)
            if plan:
                e.set_task(task_state=tr.task_state())
                pl.set_state(np.concatenate([q, v]), t, mocap)
                pl.optimize_policy()
                u = pl.action_from_policy(t)
            r = plant.forward_debug(q, v, u, mocap, time=t, warmstart=warm)
            res = r["residual"][:141]
            errs.append(np.abs(res[45:93]).max())          # mean-centred marker position errors [m]
            costs.append(plant.cost_value(res)); warm = r["qacc"]
            q, v = r["next_qpos"], r["next_qvel"]
            t += m.opt_timestep
        if plan:
            e.close()
        return np.array(errs), np.array(costs), q
    e_mpc, c_mpc, q_mpc = run(True)
    e_zero, c_zero, q_zero = run(False)
    print("humanoid track closed loop (0.4 s): mean cost MPC %.3f vs zero control %.3f; max marker error %.3f / %.3f m" % (
        c_mpc.mean(), c_zero.mean(), e_mpc.max(), e_zero.max()))
    assert np.isfinite(c_mpc).all() and q_mpc[2] > 0.8          # upright
    assert c_mpc.mean() < c_zero.mean()
    plant.set_task(task_state=np.asarray(m.task_state, float))
