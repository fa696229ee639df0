"""Quick throughput probe (steady state: warm-up long enough that ships die and missiles fly)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spacefortress_b200 import SFVecEnv
def run(gametype, n, T, warm=200, render=True):
    env = SFVecEnv(gametype, num_envs=n, device=0, render=render)
    env.reset(to_numpy=False)
    env.rollout(warm, want=("reward",))  # state-only warm-up is fine: it advances the same state
    out = {"obs": torch.empty((T, n, 1, 84, 84), dtype=torch.uint8, device="cuda")} if render else {"reward": torch.empty((T, n), dtype=torch.int32, device="cuda")}
    env.rollout(T, out=out)  # warm: fills the explosion caches of the ships that are dead right now
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); env.rollout(T, out=out); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    sps = n * T / best * 1e3
    print("%-9s n=%-7d T=%-3d render=%d  %.3f ms  %.3e steps/s  %.1f GB/s (%.1f%% of 6556)" % (gametype, n, T, render, best, sps, sps * (8362 if render else 1306) / 1e9, sps * (8362 if render else 1306) / 6556.2e7))
    env.close()
for gt in ("autoturn", "youturn"):
    run(gt, 4096, 64)
    run(gt, 65536, 16)
run("youturn", 262144, 8)
run("autoturn", 131072, 64, render=False)
