"""Target for ncu captures of the host-buffer step (sf_rollout_kernel<true>, T = 1 with the delta update inside):
   ncu --set full --import-source on -k regex:sf_rollout_kernel -s 7 -c 1 python tools/prof_host_step.py
(launches 0-3 are the four slices of the first, whole-frame step; from launch 4 on every step is one launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spacefortress_b200 import SFVecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = SFVecEnv(sys.argv[2] if len(sys.argv) > 2 else "autoturn", num_envs=n, device=0)
env.reset()
env.rollout(400, want=("reward",), action_seed=7)   # state-only kernel: a mid-episode population
acts = np.random.RandomState(0).randint(0, env.num_actions, size=(12, n)).astype(np.int32)
for t in range(12):
    env.step(acts[t])
print(env.host_delta_stats())
