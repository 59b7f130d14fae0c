#!/usr/bin/env python
"""Benchmark of the hot path: Quadruped (flat) Predictive Sampling, 256 candidates x 64-step horizon, fp32.

A "step" is one planning iteration's rollout batch (SamplingPlanner::Rollouts + ranking): 256 x 64 = 16384
simulated environment steps per GPU.  Metric: env-steps/sec (BASELINE.json).  With --gpus N every rank rolls out
its own shard of 256 candidates (weak scaling, candidates are independent units) and the per-candidate returns
are exchanged with one NCCL all-gather per iteration.

  value      device-timed throughput, inputs resident in HBM (CUDA events on the engine's stream, per launch)
  e2e        same metric through the public call (Engine.rollout_spline + winner fetch) with HOST buffers:
             H2D of state/mocap/knots and D2H of returns/order/winner trajectory inside the timed region
  roofline   dominant kernel (rollout_kernel) vs the measured HBM copy peak; algorithmic bytes per env-step are
             SURVEY.md 8(d)'s figure.  The kernel is latency-bound by construction.
  cpu_baseline  the CPU oracle (a port, the reference binary cannot be built offline) on this box's host cores.

--impl reference times that CPU path as its own arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_CAND, HORIZON, INTERP = 256, 64, 2
METRIC, UNIT = "env-steps/sec", "env-steps/s"
WORKLOAD = "Quadruped (flat) Predictive Sampling, 256 candidates x 64-step horizon, fp32"


BURN_IN = 30  # planning iterations (untimed, part of set-up) that take the zero policy to the steady-state nominal


class OracleBackend:
    """CPU oracle behind the same rollout_spline signature (only used by the CPU arms)."""

    def __init__(self, m, threads):
        from mujoco_mpc_b200.blob import to_blob
        from oracle import pyoracle
        self.o, self.threads = pyoracle.Oracle(to_blob(m), m, 64), threads

    def rollout_spline(self, state, time, mocap, knots, kt, interp, H):
        r = self.o.rollout_spline(state, time, mocap, knots, kt, interp, H, nthreads=self.threads, full=False)
        return r["returns"], r["failure"], np.argsort(r["returns"], kind="stable")


def load_inputs(backend, n_iter, rank=0, n_cand=N_CAND):
    """Model at the home keyframe (testspeed.cc:71-76); nominal spline = the planner's steady state: BURN_IN
    Predictive-Sampling iterations starting from the repeated initial action; candidates of timed
    iteration i = nominal + Philox noise with counter (BURN_IN + i, candidate, knot, dof), candidate 0 un-noised."""
    from conftest import get_model, mocap_of
    from mujoco_mpc_b200.planner import SamplingPlanner, candidate_knots
    m = get_model("quadruped")
    state = np.concatenate([m.key_qpos[0], np.zeros(m.nv)])
    mocap = mocap_of(m)
    pl = SamplingPlanner(m, backend, num_trajectory=n_cand, horizon=HORIZON, seed=0x5EED + 7919 * rank)
    pl.reset()
    pl.set_state(state, 0.0, mocap)
    for _ in range(BURN_IN):
        pl.optimize_policy()
    pl.make_candidates()  # resample the winner onto the knot grid
    knots = [candidate_knots(pl.values, pl.sigma, pl.ctrlrange, BURN_IN + it, n_cand, seed=pl.seed).astype(np.float32)
             for it in range(n_iter)]
    return m, state, mocap, knots, pl.times.copy(), float(np.min(pl.returns))


def algorithmic_bytes_per_env_step(m, P):
    ds, nu, nr, ntr = m.nq + m.nv, m.nu, m.task_num_residual, m.task_num_trace
    return 4 * (ds + nu + nr + 3 * ntr + 2) + 4 * (P * nu + ds + 7 * m.nmocap + m.nuserdata) / HORIZON


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index, self.t_mark = [], None, index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def mark(self):
        """Start of the timed region: only samples taken from here on are reported."""
        self.t_mark = time.time()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
        rows = [r for t, r in self.rows if self.t_mark is None or t >= self.t_mark]
        if not rows and self.rows:
            rows = [self.rows[-1][1]]

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [num(r[0]) for r in rows if r and num(r[0]) is not None]
        mx = [num(r[1]) for r in rows if len(r) > 1 and num(r[1]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_baseline_run(steps, warmup, threads):
    """The reference's CPU ThreadPool path restated by the oracle (fp64 = the reference's arithmetic)."""
    from conftest import get_model
    be = OracleBackend(get_model("quadruped"), threads)
    m, state, mocap, knots, kt, _ = load_inputs(be, steps + warmup)
    times = []
    for it in range(steps + warmup):
        t0 = time.perf_counter()
        ret, _, _ = be.rollout_spline(state, 0.0, mocap, knots[it], kt, INTERP, HORIZON)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return float(np.mean(times)), (state, mocap, knots[-1], kt, ret)


def ilqg_probe(m, eng, mocap):
    """BASELINE config 4 (Quadruped iLQG, H=64, 10 line-search rollouts): wall time of each C-ABI sweep with host
    buffers (H2D/D2H included), after 3 warm-up planning iterations.  Reported beside the headline, not part of it."""
    from mujoco_mpc_b200.ilqg import ILQGPlanner
    pl = ILQGPlanner(m, eng, horizon=HORIZON, num_rollouts=10, fd_tolerance=1e-3)
    pl.set_state(np.concatenate([m.key_qpos[0], np.zeros(m.nv)]), 0.0, mocap)
    for _ in range(3):
        pl.optimize_policy()

    def tm(f, reps=5):
        f(); t0 = time.perf_counter()
        for _ in range(reps):
            out = f()
        return (time.perf_counter() - t0) / reps * 1e3, out
    t_fd, (A, B, C, D) = tm(lambda: eng.model_derivatives(pl.states, pl.actions, pl.times, pl.mocap, 1e-3))
    t_cd, cd = tm(lambda: eng.cost_derivatives(pl.residual, C, D))
    t_bp, bp = tm(lambda: eng.backward_pass(A, B, cd[0], cd[1], cd[2], cd[4], cd[3], pl.actions, mu=pl.regularization))
    t_ro, _ = tm(lambda: eng.rollout_feedback(pl.state, 0.0, pl.mocap, pl.actions, pl.states, pl.times, bp["K"], bp["du"],
                                              pl._steps(), 3))
    t_it, _ = tm(lambda: pl.optimize_policy())
    fd_steps = HORIZON * (1 + m.nu + 2 * m.nv)
    return {"workload": "Quadruped (flat) iLQG, H=64, 10 line-search rollouts, one-sided FD", "fd_sweep_ms": t_fd,
            "fd_mj_step_equivalents": fd_steps, "fd_steps_per_s": fd_steps / (t_fd * 1e-3), "cost_derivatives_ms": t_cd,
            "backward_pass_ms": t_bp, "line_search_rollouts_ms": t_ro, "optimize_policy_ms": t_it,
            "timing": "host wall clock around each C-ABI call, host buffers"}


def humanoid_probe():
    """BASELINE config 3 task (Humanoid Track PS, H=128, 16 cubic knots, dt 0.005) at its per-GPU share of the 8-GPU
    configuration (128 of 1024 candidates): device-timed kernel of one planning iteration.  Reported beside the
    headline; the keyframes are synthetic clips (models.synth_mocap)."""
    from conftest import get_model
    from mujoco_mpc_b200.engine import Engine
    m = get_model("humanoid_track")
    N, H, P = 128, 128, 16
    e = Engine(m, N, H)
    mocap = np.concatenate([m.key_mpos[0].reshape(-1, 3), np.tile([1.0, 0, 0, 0], (m.nmocap, 1))], 1).reshape(-1)
    state = np.concatenate([m.key_qpos[0], np.zeros(m.nv)])
    kt = np.arange(P) * (H - 1) * 0.005 / (P - 1)
    knots = np.clip(0.15 * np.random.default_rng(0).standard_normal((N, P, m.nu)), -1, 1); knots[0] = 0
    ms = []
    for i in range(6):
        ret, fail, _ = e.rollout_spline(state, 0.0, mocap, knots, kt, 2, H)
        if i >= 2:
            ms.append(e.last_kernel_ms)
    out = {"workload": "Humanoid Track PS, 128 candidates (1/8 of 1024) x 128 steps, 16 cubic knots, dt 0.005, fp32",
           "kernel_ms": float(np.mean(ms)), "env_steps_per_s_per_gpu": N * H / (float(np.mean(ms)) * 1e-3),
           "static_kernel": bool(e.last_kernel_static), "failures": int(fail.sum())}
    e.close()
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ms, _ = cpu_baseline_run(args.steps, args.warmup, threads)
    value = N_CAND * HORIZON / ms
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "backend": "CPU oracle (restatement of mj_step + Trajectory::Rollout, ThreadPool dispatch); "
                       "the reference binary cannot be built offline (MuJoCo is fetched at configure time)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "full workload: 256 candidates x 64 steps per step"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mujoco_mpc_b200 import build
    from mujoco_mpc_b200.engine import Engine
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    n_iter = args.steps + args.warmup
    from conftest import get_model
    eng = Engine(get_model("quadruped"), N_CAND, HORIZON, device=local)
    m, state, mocap, knots, kt, nominal_return = load_inputs(eng, n_iter, rank=rank)
    P = knots[0].shape[1]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    gathered = torch.empty(world * N_CAND, dtype=torch.float32, device="cuda") if world > 1 else None
    local_ret = torch.empty(N_CAND, dtype=torch.float32, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-timed region: inputs resident, one event pair per launch, L2 flushed in between
    eng.upload_spline_inputs(state, 0.0, mocap, knots[0], kt, INTERP, HORIZON)
    eng.sync()
    clocks = ClockSampler(local)
    clocks.start()          # nvidia-smi needs ~0.5 s to produce its first row: start before the warm-up
    kern_ms, coll_ms = [], []
    launches0 = 0
    for it in range(n_iter):
        if it == args.warmup:
            barrier()
            clocks.mark()
            launches0 = eng.launch_count
            t_wall0 = time.perf_counter()
        flush.zero_()
        torch.cuda.synchronize()
        eng.launch_resident()
        eng.sync()
        k_ms = eng.last_kernel_ms
        c_ms = 0.0
        if world > 1:
            ret, _, _ = eng.read_returns()
            local_ret.copy_(torch.from_numpy(ret))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_gather_into_tensor(gathered, local_ret)
            e1.record()
            torch.cuda.synchronize()
            c_ms = e0.elapsed_time(e1)
        if it >= args.warmup:
            kern_ms.append(k_ms); coll_ms.append(c_ms)
    barrier()
    wall = time.perf_counter() - t_wall0
    clk = clocks.stop()
    gpu_launches = eng.launch_count - launches0
    total_ms = float(np.sum(kern_ms) + np.sum(coll_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * N_CAND * HORIZON / (ms_per_step * 1e-3)

    # ---------------- end-to-end through the public call with host buffers
    ds, nu, nr, ntr = eng.ds, eng.nu, eng.nr, eng.ntr
    h2d = 4 * (ds + 7 * m.nmocap + eng.info.task_state_size + N_CAND * P * nu + P)
    d2h = N_CAND * (4 + 4 + 1) + HORIZON * (4 * (ds + nu + nr + ntr + 1) + 8)
    for it in range(args.warmup):
        eng.rollout_spline(state, 0.0, mocap, knots[it], kt, INTERP, HORIZON)
    barrier()
    t0 = time.perf_counter()
    for it in range(args.steps):
        ret, fail, order = eng.rollout_spline(state, 0.0, mocap, knots[args.warmup + it], kt, INTERP, HORIZON)
        best = eng.fetch_trajectory(int(order[0]))
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * N_CAND * HORIZON / float(te.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---------------- parity + roofline + CPU baseline (rank 0)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bytes_per_launch = algorithmic_bytes_per_env_step(m, P) * N_CAND * HORIZON
    kernel_ms = float(np.mean(kern_ms))
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": "rollout_kernel_quadruped" if eng.last_kernel_static else "rollout_kernel", "kernel_ms": kernel_ms, "peak_source": peak_src,
                "algorithmic_bytes_per_env_step": algorithmic_bytes_per_env_step(m, P),
                "note": "latency/occupancy-bound by construction: 256 warps, 64 dependent steps each (DESIGN.md)"}
    prof = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(prof):
        pj = json.load(open(prof))
        roofline["traffic"] = pj.get("dram_bytes_per_launch")
        # what actually bounds the kernel (from the committed ncu capture of the same launch): issue-slot use and stalls
        roofline["latency_bound_evidence"] = {k: pj[k] for k in ("smsp__issue_active_pct", "sm__warps_active_pct_of_peak",
                                                                 "warp_instructions_per_env_step", "stall_mix_pct",
                                                                 "counters_from") if k in pj}
    ilqg = ilqg_probe(m, eng, mocap) if world == 1 else None
    config3 = humanoid_probe() if world == 1 else None
    cpu = None
    parity = None
    if not args.no_cpu_baseline and world == 1:     # the CPU arm is reported at N = 1 only
        threads = os.cpu_count() or 1
        cpu_s, (c_state, c_mocap, c_knots, c_kt, cpu_ret) = cpu_baseline_run(3, 1, threads)
        cpu = {"value": N_CAND * HORIZON / cpu_s, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "3 full steps (256 candidates x 64 steps each) after its own %d-iteration burn-in, fp64 oracle, "
                         "ThreadPool over all host threads" % BURN_IN}
        # parity on identical inputs: the CPU arm's last candidate set through the device path
        gret, _, _ = eng.rollout_spline(c_state, 0.0, c_mocap, c_knots, c_kt, INTERP, HORIZON)
        # exact operation count of the (dense, unoptimised) oracle on a 4-candidate sample of the same inputs
        from mujoco_mpc_b200.blob import to_blob
        from oracle import pyoracle
        ops, _ = pyoracle.count_flops(to_blob(m), c_state, 0.0, c_mocap, c_knots[:4], c_kt, INTERP, HORIZON)
        roofline["oracle_ops_per_env_step"] = ops
        roofline["achieved_tflops_at_oracle_op_count"] = ops * value / world / 1e12
        roofline["fp32_note"] = "operation count of the dense CPU restatement (instrumented scalar); the kernel exploits the dof-tree sparsity and executes fewer"
        rel = np.abs(gret - cpu_ret) / np.maximum(np.abs(cpu_ret), 1e-12)
        # the same candidates through the oracle instantiated in fp32: how far apart two correct implementations of
        # the same arithmetic land on these inputs (contact make/break amplifies rounding), i.e. the noise floor
        o32 = pyoracle.Oracle(to_blob(m), m, 32)
        r32 = o32.rollout_spline(c_state, 0.0, c_mocap, c_knots, c_kt, INTERP, HORIZON, nthreads=threads, full=False)["returns"]
        rel32 = np.abs(gret - r32) / np.maximum(np.abs(r32), 1e-12)
        floor = np.abs(r32 - cpu_ret) / np.maximum(np.abs(cpu_ret), 1e-12)
        parity = {"max_rel_return_err_vs_fp64_oracle": float(rel.max()), "mean_rel": float(rel.mean()),
                  "median_rel_vs_fp64_oracle": float(np.median(rel)),
                  "candidates_above_1e-4_vs_fp64": int((rel > 1e-4).sum()),
                  "max_rel_vs_nearer_oracle_precision": float(np.minimum(rel, rel32).max()),
                  "fp32_oracle_vs_fp64_oracle": {"max_rel": float(floor.max()), "median_rel": float(np.median(floor)),
                                                 "candidates_above_1e-4": int((floor > 1e-4).sum())},
                  "argmin_agrees": bool(int(np.argmin(gret)) == int(np.argmin(cpu_ret))),
                  "note": "inputs = steady-state candidates (feet in sustained contact); per-step parity is in tests/"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "candidates_per_gpu": N_CAND, "horizon": HORIZON, "spline_points": P,
                       "nominal": "steady-state policy after %d planning iterations from the zero policy (return %.4f)" % (BURN_IN, nominal_return),
                       "l2": "flushed between timed iterations (256 MB memset)", "sharding": "candidates, %d per GPU" % N_CAND,
                       "e2e_call": "Engine.rollout_spline (mjpc_b200_rollout_spline) + fetch_trajectory(winner), host buffers"},
            "clocks": clk, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(gpu_launches), "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "ilqg": ilqg, "humanoid_track": config3,
            "wall_s_timed_region": wall}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
