"""BASELINE config 3 as specified: Humanoid Track (the reference's 1889 keyframes), Predictive Sampling, 1024 candidates
x 128 steps, 16 cubic knots, dt 0.005, ONE planning problem sharded over the GPUs of one node inside the library
(mjpc_b200_rollout_spline_sharded: contiguous candidate ranges, one ncclAllGather of the returns on the engine stream,
ranking on every rank; winner trajectory by ncclBroadcast).
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 profiles/humanoid_multi.py"""
import json, os, sys, time
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
e = Engine(m, hi - lo if world > 1 else NTOT, H, device=local)
if world > 1:
    e.comm_init_torch(dist)
mocap = np.concatenate([m.key_mpos[0].reshape(-1, 3), np.tile([1.0, 0, 0, 0], (m.nmocap, 1))], 1).reshape(-1)
state = np.concatenate([m.key_qpos[0], m.key_qvel[0]])
kt = np.arange(P) * (H - 1) * 0.005 / (P - 1)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


ms, e2e = [], []
for it in range(STEPS + WARM):
    knots = np.clip(0.15 * philox_normal(it, NTOT, P, m.nu), -1, 1).astype(np.float32)   # the GLOBAL candidate set, same on every rank
    knots[0] = 0
    barrier()
    t0 = time.perf_counter()
    if world > 1:
        ret, fail, order = e.rollout_spline_sharded(state, 0.0, mocap, knots, kt, 2, H)
        best = e.fetch_trajectory_sharded(int(order[0]))
    else:
        ret, fail, order = e.rollout_spline(state, 0.0, mocap, knots, kt, 2, H)
        best = e.fetch_trajectory(int(order[0]))
    dt = time.perf_counter() - t0
    if it >= WARM:
        ms.append(e.last_kernel_ms); e2e.append(dt * 1e3)


def maxr(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


dev_ms, e2e_ms = maxr(float(np.mean(ms))), maxr(float(np.mean(e2e)))
chk = torch.tensor([float(np.sum(ret.astype(np.float64) * np.arange(1, NTOT + 1))), float(order[0]), float(np.abs(best["states"]).sum())],
                   dtype=torch.float64, device="cuda")
cmin, cmax = chk.clone(), chk.clone()
if world > 1:
    dist.all_reduce(cmin, op=dist.ReduceOp.MIN); dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": "Humanoid Track PS, 1024 candidates x 128 steps, 16 cubic knots, dt 0.005, fp32, one problem",
                      "keyframes": getattr(m, "key_source", "?"), "n_gpus": world, "candidates_per_gpu": hi - lo,
                      "device_ms_per_iteration": dev_ms, "env_steps_per_s": NTOT * H / (dev_ms * 1e-3),
                      "e2e_ms_per_iteration": e2e_ms, "e2e_env_steps_per_s": NTOT * H / (e2e_ms * 1e-3),
                      "failures": int(fail.sum()), "identical_on_all_ranks": bool(torch.equal(cmin, cmax)),
                      "static_kernel": bool(e.last_kernel_static),
                      "timing": "device: CUDA events around kernel + ncclAllGather + ranking on the engine stream; e2e: host wall clock "
                                "around rollout_spline_sharded + fetch_trajectory_sharded with host buffers; both max over ranks"}))
if world > 1:
    dist.destroy_process_group()
