"""BASELINE config 3 as specified: Humanoid Track, Predictive Sampling, 1024 candidates x 128 steps, 16 cubic knots,
dt 0.005, candidates sharded over the GPUs of one node (128 per GPU at 8), one all-gather of the returns per planning
iteration.  Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 profiles/humanoid_multi.py"""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from conftest import get_model
from mujoco_mpc_b200.engine import Engine
from mujoco_mpc_b200.planner import philox_normal
from mujoco_mpc_b200.sharding import shard_bounds
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m = get_model("humanoid_track")
NTOT, H, P, STEPS, WARM = 1024, 128, 16, 10, 3
lo, hi = shard_bounds(NTOT, world)[rank]
N = hi - lo
e = Engine(m, N, H, device=local)
mocap = np.concatenate([m.key_mpos[0].reshape(-1, 3), np.tile([1.0, 0, 0, 0], (m.nmocap, 1))], 1).reshape(-1)
state = np.concatenate([m.key_qpos[0], np.zeros(m.nv)])
kt = np.arange(P) * (H - 1) * 0.005 / (P - 1)
gathered = torch.empty(NTOT, dtype=torch.float32, device="cuda")
local_ret = torch.empty(N, dtype=torch.float32, device="cuda")
ms = []
for it in range(STEPS + WARM):
    z = philox_normal(it, NTOT, P, m.nu)[lo:hi]                       # the global noise stream, this rank's slice
    knots = np.clip(0.15 * z, -1, 1)
    if rank == 0:
        knots[0] = 0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ret, fail, _ = e.rollout_spline(state, 0.0, mocap, knots, kt, 2, H)
    k_ms = e.last_kernel_ms
    c_ms = 0.0
    if world > 1:
        local_ret.copy_(torch.from_numpy(ret))
        e0.record(); dist.all_gather_into_tensor(gathered, local_ret); e1.record()
        torch.cuda.synchronize(); c_ms = e0.elapsed_time(e1)
    if it >= WARM:
        ms.append(k_ms + c_ms)
t = torch.tensor([float(np.mean(ms))], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": "Humanoid Track PS, 1024 candidates x 128 steps, 16 cubic knots, fp32, synthetic clips",
                      "n_gpus": world, "candidates_per_gpu": N, "ms_per_iteration": float(t.item()),
                      "env_steps_per_s": NTOT * H / (float(t.item()) * 1e-3), "static_kernel": bool(e.last_kernel_static),
                      "timing": "device kernel time + all-gather (CUDA events), max over ranks"}))
if world > 1:
    dist.destroy_process_group()
