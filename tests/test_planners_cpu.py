"""CPU: the host-side planner drivers on the oracle backend (BASELINE config 1 plumbing + the reference's
behavioural planner tests)."""
import numpy as np

from conftest import OracleBackend, get_model, mocap_of


def test_ilqg_particle_reaches_goal():
    # mjpc/test/ilqg_planner/ilqg_test.cc:49-126: 25 iterations, dt 0.1, horizon 2.5 -> 26 steps
    from mujoco_mpc_b200.ilqg import ILQGPlanner
    m = get_model("particle")            # agent_timestep = 0.1 applied, residual = ParticleTestTask, risk 1
    assert abs(m.opt_timestep - 0.1) < 1e-12
    steps = int(max(min(2.5 / 0.1 + 1, 512), 1))
    pl = ILQGPlanner(m, OracleBackend(m, threads=1), horizon=steps)
    pl.set_state(np.zeros(4), 0.0, mocap_of(m))
    for _ in range(25):
        pl.optimize_policy()
    goal = mocap_of(m)[:2]
    assert abs(pl.states[-1, 0] - goal[0]) < 1e-2 and abs(pl.states[-1, 1] - goal[1]) < 1e-2
    assert abs(pl.states[-1, 2]) < 0.1 and abs(pl.states[-1, 3]) < 0.1
    cr = np.asarray(m.actuator_ctrlrange).reshape(-1, 2)
    assert (pl.actions[:-1] <= cr[:, 1] + 1e-12).all() and (pl.actions[:-1] >= cr[:, 0] - 1e-12).all()


def test_sampling_particle_reaches_goal():
    # mjpc/test/sampling_planner/sampling_planner_test.cc:44-115 (behavioural): the particle reaches the mocap goal
    from mujoco_mpc_b200.planner import SamplingPlanner
    m = get_model("particle")
    pl = SamplingPlanner(m, OracleBackend(m, threads=2), num_trajectory=16, horizon=11)
    pl.reset()
    pl.set_state(np.zeros(4), 0.0, mocap_of(m))
    for _ in range(150):
        pl.optimize_policy()
    tr = pl.backend.fetch_trajectory(pl.winner)
    assert np.abs(tr["states"][-1, :2] - mocap_of(m)[:2]).max() < 0.1
    cr = np.asarray(m.actuator_ctrlrange).reshape(-1, 2)
    assert (np.abs(tr["actions"]) <= cr[:, 1] + 1e-9).all()
    assert pl.improvement >= 0


def test_config1_cartpole_threadpool():
    """BASELINE config 1: Cartpole, Predictive Sampling, 8 candidates x 32 steps on the CPU ThreadPool path."""
    from mujoco_mpc_b200.planner import SamplingPlanner
    m = get_model("cartpole")
    pl = SamplingPlanner(m, OracleBackend(m, threads=4))
    assert pl.num_trajectory == 8 and pl.horizon == 32 and pl.P == 10 and abs(pl.sigma - 0.5) < 1e-12 and pl.interp == 2
    pl.reset()
    pl.set_state(np.concatenate([m.key_qpos[0], np.zeros(m.nv)]), 0.0, mocap_of(m))
    r0, _ = pl.optimize_policy()
    first = float(r0[0])
    for _ in range(20):
        ret, fail = pl.optimize_policy()
        assert not fail.any()
    assert float(ret[pl.winner]) <= first + 1e-9        # the winner is never worse than the initial nominal
    assert ret[0] == ret.max() or pl.improvement >= 0     # candidate 0 is the un-noised nominal (planner.cc:374)


def test_philox_noise_is_reproducible():
    from mujoco_mpc_b200.planner import philox_normal
    a = philox_normal(3, 5, 3, 12); b = philox_normal(3, 5, 3, 12); c = philox_normal(4, 5, 3, 12)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    z = philox_normal(0, 256, 3, 12).ravel()
    assert abs(z.mean()) < 0.05 and abs(z.std() - 1) < 0.05


def test_cross_entropy_particle_reaches_goal():
    """Cross-Entropy planner (mjpc/planners/cross_entropy/planner.cc) on the oracle backend: same behavioural bar as
    the reference's sampling test, plus the elite statistics (mean/variance of the n_elite best knot sets)."""
    from mujoco_mpc_b200.planner import CrossEntropyPlanner
    m = get_model("particle")
    pl = CrossEntropyPlanner(m, OracleBackend(m, threads=2), num_trajectory=16, horizon=11, n_elite=4)
    pl.set_state(np.zeros(4), 0.0, mocap_of(m))
    assert np.allclose(pl.variance, pl.std_initial ** 2)          # Reset: variance = std_initial^2 (planner.cc:141-142)
    for it in range(60):
        ret, fail = pl.optimize_policy()
        assert len(ret) == 17 and not fail.any()                   # N noisy candidates + the nominal
        # elites: mean of the best n_elite knot sets is the installed policy
        assert (ret[pl.order[:4]] <= np.sort(ret[:16])[3] + 1e-12).all()
        assert pl.improvement >= 0
    tr = pl.backend.fetch_trajectory(pl.nominal_index)            # BestTrajectory() is the nominal (planner.cc:462)
    assert np.abs(tr["states"][-1, :2] - mocap_of(m)[:2]).max() < 0.1
    cr = np.asarray(m.actuator_ctrlrange).reshape(-1, 2)
    assert (np.abs(tr["actions"]) <= cr[:, 1] + 1e-9).all()
    assert (pl.variance >= 0).all() and pl.variance.shape == (pl.P, m.nu)
