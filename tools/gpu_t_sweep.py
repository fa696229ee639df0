"""Throughput of the fused rollout for several (gametype, n, T): all repetitions printed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spacefortress_b200 import SFVecEnv
def run(gt, n, T, warm=300):
    env = SFVecEnv(gt, num_envs=n, device=0); env.reset(to_numpy=False)
    env.rollout(warm, want=("reward",))
    out = {"obs": torch.empty((T, n, 1, 84, 84), dtype=torch.uint8, device="cuda")}
    env.rollout(T, out=out); torch.cuda.synchronize()
    ms = []
    for _ in range(4):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); env.rollout(T, out=out); e.record(); torch.cuda.synchronize(); ms.append(s.elapsed_time(e))
    print("%-9s n=%-7d T=%-3d  %s ms  best %.3e steps/s" % (gt, n, T, " ".join("%.2f" % m for m in ms), n * T / min(ms) * 1e3), flush=True)
    env.close()
for gt in ("youturn", "autoturn"):
    for n, T in ((65536, 16), (65536, 64), (4096, 16), (4096, 64), (262144, 8), (262144, 32)):
        run(gt, n, T)
