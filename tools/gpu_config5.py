"""BASELINE.json configs[4] on one GPU: youturn, N envs (default 262 144), render + auto-reset under high
episode-length variance (clocks staggered uniformly over the whole episode, so ~N/5295 envs finish per step),
episode statistics read once per rollout. usage: python tools/gpu_config5.py [N] [T]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spacefortress_b200 import SFVecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
env = SFVecEnv("youturn", num_envs=n, device=0); env.reset(to_numpy=False)
env.rollout(300, want=("reward",))
env.set_ticks(np.random.RandomState(0).randint(300, 5295, size=n))
out = {"obs": torch.empty((T, n, 1, 84, 84), dtype=torch.uint8, device="cuda"), "done": torch.empty((T, n), dtype=torch.uint8, device="cuda")}
env.rollout(T, out=out); torch.cuda.synchronize()
env.episode_stats()
ms = []
for _ in range(3):
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); env.rollout(T, out=out); st = env.episode_stats(); e.record(); torch.cuda.synchronize(); ms.append(s.elapsed_time(e))
print(json.dumps({"config": "youturn, %d envs, render + auto-reset, staggered episode ends, %d-step rollouts + episode stats" % (n, T),
                  "env_steps_per_s": n * T / min(ms) * 1e3, "ms": ms, "episodes_finished_in_last_rollout": int(st["episodes"]),
                  "dones_in_buffer": int(out["done"].sum())}))
