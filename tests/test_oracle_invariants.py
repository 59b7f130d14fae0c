"""CPU: physical invariants of the oracle's forward dynamics that do not depend on any reference vectors.

The oracle and the CUDA kernels are restated from the same description of MuJoCo's pipeline, so device <-> oracle
parity cannot catch an error they share.  Conservation laws can: with gravity off and no contacts every force in the
model (actuators, joint springs and dampers, joint / tendon limits) is internal, so total linear and angular momentum -
computed here from an independent numpy restatement of the kinematics and body Jacobians (refmath.py) - are constants
of the continuous dynamics.  The semi-implicit Euler step is first order: the drift over a fixed duration must halve
when the time step halves (a wrong Coriolis / centrifugal term or mass matrix leaves a dt-independent residual)."""
import numpy as np
import pytest

from mujoco_mpc_b200 import models
from mujoco_mpc_b200 import task as T
from mujoco_mpc_b200.blob import to_blob
from mujoco_mpc_b200.mjcf import compile_xml
from mujoco_mpc_b200.refmath import body_jacobian, kinematics


def _momentum(m, q, v):
    kin = kinematics(m, q)
    P, L = np.zeros(3), np.zeros(3)
    for b in range(1, m.nbody):
        J = body_jacobian(m, kin, b, kin["xipos"][b])
        vl, w = J[:3] @ v, J[3:] @ v
        Iw = kin["ximat"][b] @ np.diag(m.body_inertia[b]) @ kin["ximat"][b].T
        P += m.body_mass[b] * vl
        L += np.cross(kin["xipos"][b], m.body_mass[b] * vl) + Iw @ w
    return P, L


def _floating(name, dt):
    """The task model with gravity switched off and the given time step."""
    if name == "humanoid":
        xml = models.humanoid_stand_xml().replace('<mujoco model="Humanoid">',
                                                  '<mujoco model="Humanoid">\n  <option gravity="0 0 0" timestep="%g"/>' % dt)
        m = compile_xml(xml, pair_filter=models._robot_vs_world_only)
        m.task_residual_id = T.RESIDUAL_HUMANOID_STAND
        ids = np.zeros(T.HI_SIZE, np.int32)
        ids[0], ids[1] = m.body_names.index("torso"), m.body_names.index("head")
        for k in range(4):
            ids[2 + k] = m.site_names.index("sp%d" % k)
        m.task_ids, m.task_state = ids, np.zeros(1)
    else:
        m = models.load("quadruped")                           # a fresh compile: safe to edit in place
        m.opt_gravity = np.zeros(3)
        m.opt_timestep = dt
    if "ray_geoms" not in m:
        m.ray_geoms = np.zeros(0, np.int32)
    return m


@pytest.mark.parametrize("name", ["humanoid", "quadruped"])
def test_momentum_conservation_first_order(oracle_lib, name):
    drift = {}
    for dt in (0.002, 0.001, 0.0005):
        m = _floating(name, dt)
        o = oracle_lib.Oracle(to_blob(m), m, 64)
        rng = np.random.default_rng(0)
        q = (m.qpos0 if name == "humanoid" else m.key_qpos[0]).copy()
        q[2] = 5.0                                               # far above the floor: no contacts
        q[7:] += 0.2 * rng.standard_normal(m.nq - 7)
        v = np.zeros(m.nv); v[6:] = rng.standard_normal(m.nv - 6)
        u = rng.uniform(-0.3, 0.3, m.nu)
        mocap = np.zeros(7 * m.nmocap)
        if m.nmocap:
            mocap = np.concatenate([np.asarray(m.mocap_pos0, float), np.asarray(m.mocap_quat0, float)], 1).reshape(-1)
        P0, L0 = _momentum(m, q, v)
        warm = None
        for k in range(int(round(0.08 / dt))):
            r = o.forward_debug(q, v, u, mocap, time=k * dt, warmstart=warm)
            assert r["ncon"] == 0
            q, v, warm = r["next_qpos"], r["next_qvel"], r["qacc"]
        P, L = _momentum(m, q, v)
        drift[dt] = (np.linalg.norm(P - P0), np.linalg.norm(L - L0))
    for k in (0, 1):                                             # linear, angular
        r1 = drift[0.002][k] / drift[0.001][k]
        r2 = drift[0.001][k] / drift[0.0005][k]
        assert 1.8 < r1 < 2.2 and 1.8 < r2 < 2.2, (name, k, drift)


def test_resting_sphere_penetration_closed_form(oracle_lib):
    """A frictionless sphere resting on the plane: the soft-contact model has a closed-form equilibrium.  With
    r = penetration, imp(r) the solimp impedance, k = 1/(dmax^2 tc^2 dr^2), R = (1-imp)/imp * (1/m):
    force = aref / R = k imp r * imp m/(1-imp) = m g   =>   r = g (1 - imp(r)) / (k imp(r)^2).
    Pins the impedance curve, the reference acceleration, the regulariser and the force law at steady state."""
    xml = """
<mujoco model="sphere">
  <option timestep="0.002"/>
  <custom><numeric name="agent_planner" data="0"/><numeric name="agent_horizon" data="0.1"/></custom>
  <worldbody>
    <geom name="floor" type="plane" size="5 5 .1" condim="1"/>
    <body name="ball" pos="0 0 0.1"><freejoint/><geom name="ball" type="sphere" size="0.1" mass="1.5" condim="1"/></body>
  </worldbody>
  <sensor><user name="Dummy" dim="13" user="0 1 0 1"/></sensor>
</mujoco>"""
    m = compile_xml(xml)
    m.task_residual_id = T.RESIDUAL_PARTICLE_COPY           # any residual: only the state is inspected
    m.task_ids, m.task_state, m.ray_geoms = np.zeros(1, np.int32), np.zeros(1), np.zeros(0, np.int32)
    o = oracle_lib.Oracle(to_blob(m), m, 64)
    q, v, warm = m.qpos0.copy(), np.zeros(m.nv), None
    for k in range(3000):                                   # 6 s: critically damped contact settles
        r = o.forward_debug(q, v, np.zeros(0), np.zeros(0), time=0.0, warmstart=warm)
        q, v, warm = r["next_qpos"], r["next_qvel"], r["qacc"]
    assert np.abs(v).max() < 1e-9 and r["ncon"] == 1 and r["nefc"] == 1
    depth = 0.1 - q[2]
    g, d0, dmax, width, mid, power, tc, dr = 9.81, 0.9, 0.95, 0.001, 0.5, 2.0, 0.02, 1.0
    kk = 1.0 / (dmax * dmax * tc * tc * dr * dr)

    def imp(x):
        a = min(x / width, 1.0)
        y = a ** power / mid ** (power - 1) if a <= mid else 1 - (1 - a) ** power / (1 - mid) ** (power - 1)
        return d0 + y * (dmax - d0)
    x = 1e-4
    for _ in range(200):
        x = g * (1 - imp(x)) / (kk * imp(x) ** 2)
    assert abs(depth - x) < 1e-9 * max(1.0, 1 / x) and 1e-5 < x < 1e-3, (depth, x)
    assert abs(r["efc_force"][0] - 1.5 * g) < 1e-8          # the contact carries exactly the weight


@pytest.mark.parametrize("cone", ["elliptic", "pyramidal"])
def test_sliding_box_coulomb_friction(oracle_lib, cone):
    """A box sliding on the plane decelerates at mu*g while it slides and then sticks (velocity exactly killed by the
    friction rows, no creep): exercises the cone zones of the Newton solver - sliding = on the cone boundary ("middle
    zone" for the elliptic cone, one active edge pair for the pyramid), stuck = inside the cone (quadratic zone)."""
    mu = 0.4
    xml = f"""
<mujoco model="box">
  <option timestep="0.002" cone="{cone}" impratio="1"/>
  <custom><numeric name="agent_planner" data="0"/><numeric name="agent_horizon" data="0.1"/></custom>
  <worldbody>
    <geom name="floor" type="plane" size="5 5 .1" friction="{mu} 0.005 0.0001"/>
    <body name="box" pos="0 0 0.05"><freejoint/><geom name="box" type="box" size="0.1 0.1 0.05" mass="2" friction="{mu} 0.005 0.0001"/></body>
  </worldbody>
  <sensor><user name="Dummy" dim="13" user="0 1 0 1"/></sensor>
</mujoco>"""
    m = compile_xml(xml)
    m.task_residual_id = T.RESIDUAL_PARTICLE_COPY
    m.task_ids, m.task_state, m.ray_geoms = np.zeros(1, np.int32), np.zeros(1), np.zeros(0, np.int32)
    o = oracle_lib.Oracle(to_blob(m), m, 64)
    q, v, warm = m.qpos0.copy(), np.zeros(m.nv), None
    for k in range(500):                                    # settle on the floor
        r = o.forward_debug(q, v, np.zeros(0), np.zeros(0), warmstart=warm)
        q, v, warm = r["next_qpos"], r["next_qvel"], r["qacc"]
    assert r["ncon"] == 4 and np.abs(v).max() < 1e-6
    v[0] = 1.0                                              # shove it along x
    vx = []
    for k in range(200):
        r = o.forward_debug(q, v, np.zeros(0), np.zeros(0), warmstart=warm)
        q, v, warm = r["next_qpos"], r["next_qvel"], r["qacc"]
        vx.append(v[0])
    vx = np.array(vx)
    t = 0.002 * (np.arange(200) + 1)
    sliding = vx > 0.2
    decel = -np.polyfit(t[sliding], vx[sliding], 1)[0]
    assert abs(decel - mu * 9.81) < 0.03 * mu * 9.81, (cone, decel, mu * 9.81)
    assert abs(vx[-1]) < 1e-4 and abs(v[1]) < 1e-6 and abs(q[2] - 0.05) < 1e-3     # stopped, stuck, still on the floor
    # the last few cm/s decay smoothly (soft constraint): compare the time to shed 95 % of the speed
    t95 = 0.95 / (mu * 9.81)
    assert abs(t[np.argmax(vx < 0.05)] - t95) < 0.05 * t95, (cone, t[np.argmax(vx < 0.05)], t95)


def test_capsule_capsule_closed_form(oracle_lib):
    """[EXT] mjc_CapsuleCapsule restated (oracle/physics.h collide_capsule_capsule): crossed capsules touch at one point
    midway between the axes; parallel overlapping capsules produce two contacts at the ends of the overlap; contacts
    between two MOVING bodies push them apart with equal and opposite forces (momentum conserved)."""
    from mujoco_mpc_b200.blob import to_blob
    from mujoco_mpc_b200.mjcf import compile_xml
    xml = """
    <mujoco>
      <option timestep="0.002" gravity="0 0 0"/>
      <custom><numeric name="agent_planner" data="0"/><numeric name="agent_horizon" data="0.1"/></custom>
      <worldbody>
        <body name="a" pos="0 0 0"><freejoint/><geom name="ga" type="capsule" size="0.05 0.3" quat="0.7071068 0 0.7071068 0" mass="1" condim="1"/></body>
        <body name="b" pos="0 0 0"><freejoint/><geom name="gb" type="capsule" size="0.04 0.2" quat="0.7071068 0.7071068 0 0" mass="2" condim="1"/></body>
      </worldbody>
      <sensor><user name="Dummy" dim="26" user="0 1 0 1"/></sensor>
    </mujoco>"""
    m = compile_xml(xml)
    m.task_residual_id = T.RESIDUAL_PARTICLE_COPY           # any residual: only the state is inspected
    m.task_ids, m.task_state, m.ray_geoms = np.zeros(1, np.int32), np.zeros(1), np.zeros(0, np.int32)
    o = oracle_lib.Oracle(to_blob(m), m, 64)
    ident = [1.0, 0, 0, 0]
    # crossed: A along x at z = 0, B along y at z = 0.08 -> penetration 0.01 along z at (0, 0, ~0.045)
    q = np.array([0, 0, 0] + ident + [0, 0, 0.08] + ident, float)
    r = o.forward_debug(q, np.zeros(12), np.zeros(0), np.zeros(0))
    assert r["ncon"] == 1
    c = r["contact"][0]
    np.testing.assert_allclose(c[0], 0.08 - 0.05 - 0.04, atol=1e-12)
    np.testing.assert_allclose(c[1:4], [0, 0, 0.05 + 0.5 * (0.08 - 0.09)], atol=1e-12)
    np.testing.assert_allclose(np.abs(c[4:7]), [0, 0, 1], atol=1e-12)
    # equal and opposite: total linear momentum change is zero (free bodies, no gravity)
    acc = r["qacc"]
    np.testing.assert_allclose(1.0 * acc[0:3] + 2.0 * acc[6:9], 0, atol=1e-9)
    assert acc[2] < 0 < acc[8]                                   # A pushed down, B pushed up
    # parallel: both along x (B rotated onto x), B shifted by 0.25 in x and 0.085 in z -> overlap x in [0.05, 0.3]
    s = np.sqrt(0.5)
    qb = [s, 0, 0, -s]                                           # undo the geom's y-orientation: rotate -90 deg about z
    q = np.array([0, 0, 0] + ident + [0.25, 0, 0.085] + qb, float)
    r = o.forward_debug(q, np.zeros(12), np.zeros(0), np.zeros(0))
    assert r["ncon"] == 2
    xs = sorted(r["contact"][k][1] for k in range(2))
    np.testing.assert_allclose(xs, [0.05, 0.3], atol=1e-9)       # ends of the overlap: B's lower end, A's upper end
    for k in range(2):
        np.testing.assert_allclose(r["contact"][k][0], 0.085 - 0.09, atol=1e-12)
    # separated beyond the margin: nothing
    q = np.array([0, 0, 0] + ident + [0, 0, 0.2] + ident, float)
    assert o.forward_debug(q, np.zeros(12), np.zeros(0), np.zeros(0))["ncon"] == 0
