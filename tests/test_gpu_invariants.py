"""GPU: the physical invariants of tests/test_oracle_invariants.py evaluated on the DEVICE physics (step_debug through
the C ABI): they do not involve the oracle at all, so they catch an error the oracle and the kernels could share."""
import numpy as np
import pytest

from test_oracle_invariants import _floating, _momentum

pytestmark = pytest.mark.gpu


def _simple(xml):
    from mujoco_mpc_b200 import task as T
    from mujoco_mpc_b200.mjcf import compile_xml
    m = compile_xml(xml)
    m.task_residual_id = T.RESIDUAL_PARTICLE_COPY
    m.task_ids, m.task_state, m.ray_geoms = np.zeros(1, np.int32), np.zeros(1), np.zeros(0, np.int32)
    return m


def _run(e, q, v, steps, nu=0):
    warm, out = None, []
    for k in range(steps):
        r = e.step_debug(q, v, np.zeros(nu), np.zeros(0), warmstart=warm)
        q, v, warm = r["next_qpos"].astype(float), r["next_qvel"].astype(float), r["qacc"]
        out.append((q.copy(), v.copy()))
    return q, v, r, out


def test_resting_sphere_penetration_on_device():
    from mujoco_mpc_b200.engine import Engine
    m = _simple("""
<mujoco model="sphere">
  <option timestep="0.002"/>
  <custom><numeric name="agent_planner" data="0"/><numeric name="agent_horizon" data="0.1"/></custom>
  <worldbody>
    <geom name="floor" type="plane" size="5 5 .1" condim="1"/>
    <body name="ball" pos="0 0 0.1"><freejoint/><geom name="ball" type="sphere" size="0.1" mass="1.5" condim="1"/></body>
  </worldbody>
  <sensor><user name="Dummy" dim="13" user="0 1 0 1"/></sensor>
</mujoco>""")
    e = Engine(m, 4, 4)
    q, v, r, _ = _run(e, m.qpos0.copy(), np.zeros(m.nv), 2500)
    e.close()
    g, d0, dmax, width, mid, power, tc = 9.81, 0.9, 0.95, 0.001, 0.5, 2.0, 0.02
    kk = 1.0 / (dmax * dmax * tc * tc)

    def imp(x):
        a = min(x / width, 1.0)
        y = a ** power / mid ** (power - 1) if a <= mid else 1 - (1 - a) ** power / (1 - mid) ** (power - 1)
        return d0 + y * (dmax - d0)
    x = 1e-4
    for _ in range(200):
        x = g * (1 - imp(x)) / (kk * imp(x) ** 2)
    depth = 0.1 - q[2]
    assert r["ncon"] == 1 and np.abs(v).max() < 1e-4
    assert abs(depth - x) < 0.02 * x, (depth, x)             # fp32: position resolution 1e-8 m at z = 0.1


@pytest.mark.parametrize("cone", ["elliptic", "pyramidal"])
def test_sliding_box_coulomb_friction_on_device(cone):
    from mujoco_mpc_b200.engine import Engine
    mu = 0.4
    m = _simple(f"""
<mujoco model="box">
  <option timestep="0.002" cone="{cone}" impratio="1"/>
  <custom><numeric name="agent_planner" data="0"/><numeric name="agent_horizon" data="0.1"/></custom>
  <worldbody>
    <geom name="floor" type="plane" size="5 5 .1" friction="{mu} 0.005 0.0001"/>
    <body name="box" pos="0 0 0.05"><freejoint/><geom name="box" type="box" size="0.1 0.1 0.05" mass="2" friction="{mu} 0.005 0.0001"/></body>
  </worldbody>
  <sensor><user name="Dummy" dim="13" user="0 1 0 1"/></sensor>
</mujoco>""")
    e = Engine(m, 4, 4)
    q, v, r, _ = _run(e, m.qpos0.copy(), np.zeros(m.nv), 500)
    assert r["ncon"] == 4
    v[0] = 1.0
    q, v, r, hist = _run(e, q, v, 200)
    e.close()
    vx = np.array([h[1][0] for h in hist]); t = 0.002 * (np.arange(200) + 1)
    sliding = vx > 0.2
    decel = -np.polyfit(t[sliding], vx[sliding], 1)[0]
    assert abs(decel - mu * 9.81) < 0.04 * mu * 9.81, (cone, decel)
    assert abs(vx[-1]) < 1e-3 and abs(q[2] - 0.05) < 1e-3


def test_momentum_conservation_on_device():
    from mujoco_mpc_b200.engine import Engine
    drift = {}
    for dt in (0.004, 0.002):
        m = _floating("quadruped", dt)
        e = Engine(m, 4, 4)
        rng = np.random.default_rng(0)
        q = m.key_qpos[0].copy(); q[2] = 5.0; q[7:] += 0.2 * rng.standard_normal(m.nq - 7)
        v = np.zeros(m.nv); v[6:] = rng.standard_normal(m.nv - 6)
        u = rng.uniform(-0.3, 0.3, m.nu)
        mocap = np.concatenate([np.asarray(m.mocap_pos0, float), np.asarray(m.mocap_quat0, float)], 1).reshape(-1)
        P0, L0 = _momentum(m, q, v)
        warm = None
        for k in range(int(round(0.08 / dt))):
            r = e.step_debug(q, v, u, mocap, time=k * dt, warmstart=warm)
            q, v, warm = r["next_qpos"].astype(float), r["next_qvel"].astype(float), r["qacc"]
        e.close()
        P, L = _momentum(m, q, v)
        drift[dt] = (np.linalg.norm(P - P0), np.linalg.norm(L - L0))
    for k in (0, 1):
        ratio = drift[0.004][k] / drift[0.002][k]
        assert 1.7 < ratio < 2.3, (k, drift)
