import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def mocap_of(m):
    if m.nmocap == 0:
        return np.zeros(0)
    return np.concatenate([np.concatenate([m.mocap_pos0[i], m.mocap_quat0[i]]) for i in range(m.nmocap)])


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


_MODELS = {}


def get_model(name, agent_timestep=True):
    from mujoco_mpc_b200 import models
    key = (name, agent_timestep)
    if key not in _MODELS:
        _MODELS[key] = models.load(name, agent_timestep=agent_timestep)
    return _MODELS[key]


@pytest.fixture(scope="session")
def quadruped():
    return get_model("quadruped")


def quadruped_inputs(m, N=16, H=64, seed=0, sigma=0.04, iteration=0):
    """Seeded synthetic planner inputs: home keyframe, zero nominal, Philox noise."""
    from mujoco_mpc_b200.planner import candidate_knots
    state = np.concatenate([m.key_qpos[0], np.zeros(m.nv)])
    P = 3
    T = (H - 1) * m.opt_timestep
    kt = np.arange(P) * T / (P - 1)
    cr = np.asarray(m.actuator_ctrlrange).reshape(-1, 2)
    knots = candidate_knots(np.zeros((P, m.nu)), sigma, cr, iteration, N, seed=0x5EED + seed)
    return state, mocap_of(m), knots, kt
