"""Profiling target: steady-state fused rollout (few steps) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spacefortress_b200 import SFVecEnv
gt = sys.argv[1] if len(sys.argv) > 1 else "autoturn"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
T = int(sys.argv[3]) if len(sys.argv) > 3 else 4
env = SFVecEnv(gt, num_envs=n, device=0)
env.reset(to_numpy=False)
env.rollout(300, want=("reward",))          # state-only warm-up into steady state (sf_step_only_kernel)
obs = torch.empty((T, n, 1, 84, 84), dtype=torch.uint8, device="cuda")
env.rollout(T, out={"obs": obs})            # warm (fills explosion memos)
env.rollout(T, out={"obs": obs})            # <- profile this one (2nd sf_rollout_kernel launch)
torch.cuda.synchronize()
print("ok")
