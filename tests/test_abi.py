"""CPU: the C-ABI library loads and exports every symbol include/mjpc_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest


def test_header_symbols_exported():
    from mujoco_mpc_b200 import build
    from mujoco_mpc_b200.engine import EXPORTS
    so = build.build()
    lib = ctypes.CDLL(so)
    hdr = open(os.path.join(os.path.dirname(build.HERE), "include", "mjpc_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mjpc_b200_[a-z_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(EXPORTS) == declared
    lib.mjpc_b200_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.mjpc_b200_version()


def test_no_cpu_fallback():
    """Without a CUDA device create() must fail loudly with MJPC_B200_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from conftest import get_model
    from mujoco_mpc_b200.engine import Engine, EngineError
    with pytest.raises(EngineError) as e:
        Engine(get_model("particle"), 4, 8)
    assert "-4" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "mujoco_mpc_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and '"oracle/' not in txt and "../oracle" not in txt, f
