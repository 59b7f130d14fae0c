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


def test_noisy_rollout_force_perturbation_equals_equivalent_controls():
    """NoisyRollout (mjpc/trajectory.cc:100-210): Ornstein-Uhlenbeck xfrc_applied noise from the injected Philox
    stream.  On the particle (two slide joints on one body, unit-gear motors on the same joints) a Cartesian force on
    the body is exactly a control: the noisy rollout with zero controls must reproduce the clean rollout driven by the
    recomputed noise sequence - this pins the noise definition, the OU recursion and the force -> joint-space map."""
    from mujoco_mpc_b200.blob import to_blob
    from mujoco_mpc_b200.planner import philox4x32
    from oracle import pyoracle
    m = get_model("particle")
    o = pyoracle.Oracle(to_blob(m), m, 64)
    H, std, rate_s, seed = 12, 0.05, 0.2, 1234
    body = m.body_names.index("pointmass") if "pointmass" in m.body_names else int(m.jnt_bodyid[0])
    assert int(m.jnt_bodyid[0]) == int(m.jnt_bodyid[1]) == body
    state = np.array([0.05, -0.02, 0.1, 0.0])
    kt = np.arange(H) * m.opt_timestep - 1e-9     # knots are reached robustly despite the accumulated rollout clock
    o.set_xfrc_noise(std, rate_s, seed)
    noisy = o.rollout_spline(state, 0.0, mocap_of(m), np.zeros((2, H, m.nu)), kt, 0, H)
    o.set_xfrc_noise(0.0)
    # recompute the perturbation of candidate (stream) 1: counter (step, stream, element, 'XFRC'), key (seed, 1)
    rate = np.exp(-m.opt_timestep / rate_s); scale = std * np.sqrt(1 - rate * rate)
    x = np.zeros(6 * m.nbody); forces = np.zeros((H, 2))
    for t in range(H - 1):
        ctr = np.array([[t, 1, e, 0x58465243] for e in range(6 * m.nbody)], np.uint32)
        r = philox4x32(ctr, (seed, 1))
        u1 = (r[:, 0].astype(np.float64) + 0.5) / 4294967296.0; u2 = (r[:, 1].astype(np.float64) + 0.5) / 4294967296.0
        x = rate * x + scale * np.sqrt(-2 * np.log(u1)) * np.cos(2 * np.pi * u2)
        forces[t] = x[6 * body: 6 * body + 2]
    assert np.abs(forces).max() < 1.0                      # inside the control range, so the clamp is inactive
    clean = o.rollout_spline(state, 0.0, mocap_of(m), forces[None], kt, 0, H)
    np.testing.assert_allclose(noisy["states"][1], clean["states"][0], atol=1e-12)
    assert np.abs(noisy["states"][1] - noisy["states"][0]).max() > 1e-5      # streams differ per candidate


def test_robust_planner_cpu():
    """Robust planner (robust_planner.cc:91-157) on the oracle backend: the installed candidate minimises the mean of
    its noisy-rollout returns among the top candidates; with zero force noise it reduces to the sampling planner."""
    from mujoco_mpc_b200.planner import RobustPlanner, SamplingPlanner
    m = get_model("particle")
    st, mc = np.array([0.1, -0.1, 0.0, 0.0]), mocap_of(m)
    rp = RobustPlanner(m, OracleBackend(m, threads=2), num_trajectory=20, horizon=11, ncandidates=4, nrepetitions=3,
                       xfrc_std=0.2, xfrc_rate=0.1)
    assert rp.ncandidates == 4 and rp.nrepetitions == 3
    rp.reset(); rp.set_state(st, 0.0, mc)
    for _ in range(5):
        ret, fail = rp.optimize_policy()
        assert rp.scores.shape == (4,) and rp.winner in list(np.argsort(ret, kind="stable")[:4])
        assert rp.scores.min() == rp.scores[list(np.argsort(ret, kind="stable")[:4]).index(rp.winner)]
    # defaults from the reference: 5 repetitions, candidates = trajectories / repetitions, std 0.1, rate 0.1
    d = RobustPlanner(m, OracleBackend(m, threads=1), num_trajectory=20, horizon=11)
    assert (d.nrepetitions, d.ncandidates, d.xfrc_std, d.xfrc_rate) == (5, 4, 0.1, 0.1)
    # no force noise: every repetition reproduces the clean return -> same winner as Predictive Sampling
    r0 = RobustPlanner(m, OracleBackend(m, threads=2), num_trajectory=20, horizon=11, ncandidates=4, nrepetitions=2, xfrc_std=0.0)
    sp = SamplingPlanner(m, OracleBackend(m, threads=2), num_trajectory=20, horizon=11)
    for p in (r0, sp):
        p.reset(); p.set_state(st, 0.0, mc)
    for _ in range(3):
        r0.optimize_policy(); sp.optimize_policy()
        assert r0.winner == sp.winner
        np.testing.assert_allclose(r0.delegate.values, sp.values, atol=1e-12)


def test_host_ilqg_policy_action_matches_oracle():
    """iLQGPolicy::Action on the host (csrc/host/ilqg_planner.cc, what the physics thread calls between plans) vs the
    oracle's policy: zero / linear / cubic interpolation of nominal actions, states (free-joint quaternion
    renormalised) and gains, tangent-space state difference, feedback scaling, clamp; before / inside / after the plan."""
    import ctypes as C

    from mujoco_mpc_b200.blob import to_blob
    from mujoco_mpc_b200.engine import ModelBlob, load_library
    from oracle import pyoracle
    lib = load_library()
    m = get_model("quadruped")
    blob = to_blob(m)
    buf = C.create_string_buffer(blob, len(blob))
    mb = ModelBlob(C.cast(buf, C.c_void_p), len(blob))
    o = pyoracle.Oracle(blob, m, 64)
    rng = np.random.default_rng(4)
    H, nu, ds, n = 12, m.nu, m.nq + m.nv, 2 * m.nv
    t = 0.5 + np.cumsum(rng.uniform(0.005, 0.02, H))
    u = rng.uniform(-0.6, 0.6, (H, nu)).astype(np.float32)
    x = np.tile(np.concatenate([m.key_qpos[0], np.zeros(m.nv)]), (H, 1)) + 0.05 * rng.standard_normal((H, ds))
    for k in range(H):
        x[k, 3:7] /= np.linalg.norm(x[k, 3:7])
    x = x.astype(np.float32)
    K = (0.3 * rng.standard_normal((H, nu, n))).astype(np.float32)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for mode in (0, 1, 2):
        for time in (t[0] - 0.1, t[0], 0.5 * (t[3] + t[4]), t[7] + 1e-4, t[-2] + 0.001, t[-1] + 1.0):
            for scale in (1.0, 0.37):
                state = x[5].astype(np.float64) + 0.02 * rng.standard_normal(ds)
                state[3:7] /= np.linalg.norm(state[3:7])
                for st in (state, None):
                    out = np.zeros(nu)
                    rc = lib.mjpc_b200_host_ilqg_policy_action(C.byref(mb), fp(u), fp(x), dp(t), fp(K), H, mode, C.c_double(scale),
                                                               dp(st) if st is not None else None, C.c_double(time), dp(out))
                    assert rc == 0
                    if st is not None:
                        ref = o.ilqg_policy_action(u, x, t, K, mode, scale, st, time)
                    else:   # open loop: the oracle's policy with the nominal state itself (zero feedback) at zero scale
                        ref = o.ilqg_policy_action(u, x, t, K, mode, 0.0, state, time)
                    np.testing.assert_allclose(out, ref, atol=1e-9, err_msg="mode %d time %.4f" % (mode, time))
                    assert (np.abs(out) <= 1 + 1e-12).all()


def test_gradient_and_ilqs_mirrors_on_oracle():
    """The Gradient / iLQS host logic (mujoco_mpc_b200/gradient.py) on the CPU oracle backend: the spline mapping is the
    exact linear operator of the spline sampler, the gradient planner descends monotonically, iLQS switches to iLQG when
    sampling cannot improve and converts back (ilqs/planner.cc:98-172)."""
    from conftest import OracleBackend
    from mujoco_mpc_b200.gradient import GradientPlanner, ILQSPlanner, gradient_sweep, spline_mapping
    from mujoco_mpc_b200.planner import sample_spline
    rng = np.random.default_rng(0)
    for rep in (0, 1, 2):
        ti = np.linspace(0.2, 0.7, 6); to = np.linspace(0.2, 0.69, 30); p = rng.standard_normal((6, 2))
        W = spline_mapping(ti, to, rep)
        np.testing.assert_allclose(W @ p, np.stack([sample_spline(ti, p, rep, t) for t in to]), atol=1e-12)
    # gradient sweep against brute-force differentiation of a random linear-quadratic chain
    T, n, mm = 5, 3, 2
    A = rng.standard_normal((T, n, n)) * 0.3; B = rng.standard_normal((T, n, mm)); cx = rng.standard_normal((T, n)); cu = rng.standard_normal((T, mm))
    k, dV0 = gradient_sweep(A, B, cx, cu)
    def total(du):          # first-order cost change of action perturbations du [T-1][m] through the linear dynamics
        dx = np.zeros(n); c = 0.0
        for t in range(T - 1):
            c += cx[t] @ dx + cu[t] @ du[t]; dx = A[t] @ dx + B[t] @ du[t]
        return c + cx[T - 1] @ dx
    for t in range(T - 1):
        for j in range(mm):
            du = np.zeros((T - 1, mm)); du[t, j] = 1.0
            assert abs(total(du) + k[t, j]) < 1e-12          # k = -dJ/du
    m = get_model("particle")
    pl = GradientPlanner(m, OracleBackend(m, threads=2), horizon=26, num_trajectory=8, num_spline_points=5, representation=1, fd_tolerance=1e-5)
    pl.set_state(np.zeros(4), 0.0, mocap_of(m))
    rets = []
    for _ in range(12):
        pl.optimize_policy(); rets.append(pl.total_return)
    assert rets[-1] < rets[0] and all(b <= a + 1e-12 for a, b in zip(rets, rets[1:]))
    il = ILQSPlanner(m, OracleBackend(m, threads=2), OracleBackend(m, threads=2), horizon=26, num_trajectory=8, num_rollouts=6, fd_tolerance=1e-5)
    il.set_state(np.zeros(4), 0.0, mocap_of(m))
    seq = []
    for it in range(8):
        il.sampling.sigma = 0.0 if 2 <= it < 6 else 0.1
        il.optimize_policy(); seq.append(il.active_policy)
    assert seq[:2] == [0, 0] and 1 in seq
    assert il.ilqg.total_return < float(il.sampling.returns[0]) * 1.5
