#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (env step + 84x84 render + auto-reset).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs-per-gpu E]

Metric (BASELINE.json): env-steps/sec with 84x84 obs; workload = configs[1]: autoturn, 4096 batched envs per
GPU, 84x84 grayscale obs, random actions (weak scaling: every rank owns its own slab of 4096 envs; there is
no data-path collective, only the episode-stat all-reduce after the timed region).

One "step" = one pass of the hot path over the whole batch (one env-step of every env). The timed region
is ONE launch of the fused multi-step kernel (sf_rollout, T = K steps; actions come from the on-device
counter-hash policy, so inputs are resident), bracketed by barrier + synchronize and timed with CUDA events
on the launching stream; max over ranks. `e2e` is the same metric through the numpy drop-in API
(SFVecEnv.step(np.ndarray) -> sf_step_host): host actions in, host observations out, copies inside the
timed region. `roofline` is for the dominant (only) kernel. `cpu_baseline` / `--impl reference` time the
reference's own CPU implementation on the host cores (see reference_arm()).
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 4096
GAMETYPE = "autoturn"
METRIC = "env-steps/sec with 84x84 obs"
B_RENDER = 84 * 84 + 10 + 2 * 648  # obs + action/reward/done/info + state read+write = 8362 B
WORKLOAD = "autoturn, 4096 batched envs per GPU, 84x84 grayscale obs, random actions (BASELINE.json configs[1])"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU implementation of the path on the host cores
# ------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, gametype, seed, steps = args
    import numpy as np
    from oracle.oracle import OracleEnv, RefEnv
    helper = OracleEnv(gametype, 1)
    rng = np.random.RandomState(seed)
    km = np.array([helper.keymask(a) for a in rng.randint(0, helper.num_actions(1), steps)], np.uint8)
    env = RefEnv(gametype, 1) if kind == "reference" else OracleEnv(gametype, 1)
    t0 = time.perf_counter()
    env.run(km, render=True)
    return time.perf_counter() - t0


def cpu_arm(steps_per_core, cores=None):
    """One process per host core (the SubprocVecEnv layout of rl/train.py:30-34, without the per-step pipe
    round trip), each running `steps_per_core` env steps of: the UNMODIFIED reference core compiled from
    /root/reference (oracle/_ref: key events, Game::stepOneTick, shaping, auto-reset) + a frame per step.
    libcairo cannot be built offline, so the frame is drawn by the restated renderer (oracle/sf_draw_oracle.c,
    with the static background and glyph masks cached like cairo's own caches) + INTER_AREA 84x84."""
    from oracle.oracle import ref_available
    kind = "reference" if ref_available() else "port"
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(kind, GAMETYPE, 0, 64)] * cores)  # warm-up: page in the libraries
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(kind, GAMETYPE, 100 + i, steps_per_core) for i in range(cores)])
        wall = time.perf_counter() - t0
    return dict(value=steps_per_core * cores / wall, unit="env-steps/s", cores=cores, kind=kind,
                sample="%d procs x %d steps of %s, tick by the %s, frame by the restated renderer (no cairo offline) + INTER_AREA; %.1f s wall"
                       % (cores, steps_per_core, GAMETYPE, "compiled reference core" if kind == "reference" else "oracle port", wall)), wall


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    # each "step" is a bounded sample of the workload: `sample` envs stepped once per core-process
    per_core = 3000
    vals = []
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_arm(200, cores)
    t_all = time.perf_counter()
    for _ in range(max(1, min(args.steps, 3))):
        b, wall = cpu_arm(per_core, cores)
        vals.append(b)
    best = max(vals, key=lambda b: b["value"])
    line = {
        "impl": "reference", "metric": METRIC, "value": best["value"], "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * ENVS_PER_GPU * args.gpus / best["value"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs": ENVS_PER_GPU * args.gpus, "gametype": GAMETYPE},
        "cpu_baseline": best,
        "e2e": {"value": best["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = os.path.join(ROOT, "gpurun_out", "clocks_rank%d.csv" % index)

    def start(self):
        try:
            os.makedirs(os.path.dirname(self.path), exist_ok=True)
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, reasons, smmax = [], set(), None
        for ln in open(self.path):
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smmax = float(p[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=smmax, reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def ours(args):
    import numpy as np
    import torch
    from spacefortress_b200 import SFVecEnv
    from spacefortress_b200 import dist as sfdist
    import torch.distributed as dist

    # Native libraries print to stdout (NCCL's version banner at init, for one): stdout is for the ONE JSON line, so
    # everything else goes to stderr until the line is printed
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    rank, local, world = sfdist.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n = args.envs_per_gpu
    K, Wm = args.steps, max(args.warmup, 3)
    first = rank * n
    env = SFVecEnv(GAMETYPE, num_envs=n, device=local, first_global_env=first)
    env.reset(to_numpy=False)
    # All envs start an episode together; the first ticks after a reset (no dead ships, no missiles in flight)
    # are cheaper than the long-run mix. Advance the state (render off) so the timed steps see a
    # representative mid-episode population. Not part of W or K.
    if args.presteps > 0:
        env.rollout(args.presteps, want=("reward",), action_seed=args.seed + 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: one fused launch of K steps -----------------------------------------
    obs_buf = torch.empty((K, n, 1, 84, 84), dtype=torch.uint8, device=dev)  # K * 28.9 MB >> L2 for K >= 8
    out = {"obs": obs_buf, "reward": torch.empty((K, n), dtype=torch.int32, device=dev),
           "done": torch.empty((K, n), dtype=torch.uint8, device=dev), "kill": torch.empty((K, n), dtype=torch.uint8, device=dev)}
    wout = {k: v[:Wm] if Wm <= K else torch.empty((Wm,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev) for k, v in out.items()}
    env.rollout(Wm, out=wout, action_seed=args.seed)  # W untimed warm-up steps
    clocks = ClockSampler(local)
    clocks.start()
    reps = []
    for rep in range(args.repeats):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        env.rollout(K, out=out, action_seed=args.seed)  # exactly K steps, 1 kernel launch
        e1.record()
        barrier()
        reps.append(e0.elapsed_time(e1))
    ms = sorted(reps)[len(reps) // 2]  # median launch: every launch times exactly K steps of the evolving batch
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = n * world * K / (ms * 1e-3)

    # ---- e2e through the numpy drop-in API (host actions -> host observations) ------------------------
    rng = np.random.RandomState(1 + rank)
    acts = rng.randint(0, env.num_actions, size=(args.e2e_steps + 3, n)).astype(np.int32)
    for t in range(3):
        env.step(acts[t])
    barrier()
    t0 = time.perf_counter()
    for t in range(args.e2e_steps):
        o, r, d, info = env.step(acts[3 + t])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = n * world * args.e2e_steps / e2e_s
    clk = clocks.stop()

    stats = env.episode_stats(reset=False)  # the one collective of the path (NCCL all-reduce when world > 1)
    peak, peak_src = measured_peak_gbs()
    achieved = n * K * B_RENDER / (ms * 1e-3) / 1e9  # this rank's kernel: algorithmic bytes per launch / launch duration
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum per env-step of the committed `ncu --set full` capture, x the env-steps of this launch
    try:
        import glob
        prof = json.load(open(sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_traffic.json")))[-1]))  # latest capture
        traffic = prof["dram_bytes_per_env_step"] * n * K
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": n, "envs": n * world, "gametype": GAMETYPE, "kernel": "sf_rollout_kernel, T=K steps per launch (one block of 24 warps per SM: a stepping warp up to two ticks ahead + 23 drawing warps in a block-cooperative frame pipeline)", "presteps": args.presteps,
                   "l2": "obs output %.1f MB/step streams into a K-step buffer (%.0f MB) larger than L2; env state (%.1f MB) is intentionally cache resident"
                         % (n * 7056 / 1e6, K * n * 7056 / 1e6, env.state_bytes() / 1e6),
                   "timing": "median of %d launches of K steps each, CUDA events on the launching stream, max over ranks" % args.repeats, "launch_ms_all": [round(x, 4) for x in reps]},
        "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]},
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": n * 4, "d2h_bytes_per_step": n * (7056 + 4 + 1 + 1 + 4),
                "api": "SFVecEnv.step(np.ndarray) -> sf_step_host, actions from and results into page-locked numpy buffers, %d steps" % args.e2e_steps},
        "gpu_launches": 1,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "bytes_per_env_step": B_RENDER, "env_steps_per_launch": n * K, "launch_ms": ms},
        "episode_stats": {k: stats[k] for k in ("episodes", "sum_return", "fort_kills")},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            b, _ = cpu_arm(args.cpu_steps_per_core)
            line["cpu_baseline"] = b
        except Exception as ex:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": "env-steps/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (ex,)}
    sys.stdout.flush()
    os.dup2(json_fd, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    env.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--repeats", type=int, default=5)
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--cpu-steps-per-core", type=int, default=40000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--presteps", type=int, default=400, help="untimed state-only ticks before W and K (mid-episode mix)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
