"""Soak check: the fused multi-tick rollout (two-tick stages, stepping warp ahead of the drawing warps) against
single steps (one tick per stage) from the same start, frame by frame, for many envs and ticks. A race in the
pipelined kernel would show up as a sporadic mismatch. usage: python tools/gpu_soak.py [gametype] [n] [ticks] [T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spacefortress_b200 import SFVecEnv
gt = sys.argv[1] if len(sys.argv) > 1 else "autoturn"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ticks = int(sys.argv[3]) if len(sys.argv) > 3 else 512
T = int(sys.argv[4]) if len(sys.argv) > 4 else 64
a = SFVecEnv(gt, num_envs=n, device=0); b = SFVecEnv(gt, num_envs=n, device=0)
a.reset(to_numpy=False); b.reset(to_numpy=False)
bad = 0
g = torch.Generator(device="cuda"); g.manual_seed(7)
for t0 in range(0, ticks, T):
    acts = torch.randint(0, a.num_actions, (T, n), generator=g, device="cuda", dtype=torch.int32)
    out = a.rollout(T, actions=acts)
    for t in range(T):
        obs, rew, done, info = b.step(acts[t])
        if not (torch.equal(out["obs"][t], obs) and torch.equal(out["reward"][t], rew) and torch.equal(out["done"][t].bool(), done)):
            bad += 1
            d = (out["obs"][t] != obs).flatten(1).any(1).nonzero().flatten()[:8].tolist()
            print("MISMATCH tick", t0 + t, "envs", d, flush=True)
same_state = bytes(a.get_state()) == bytes(b.get_state())
print("%s n=%d ticks=%d T=%d: %d mismatching ticks, final states equal: %s" % (gt, n, ticks, T, bad, same_state))
sys.exit(1 if bad or not same_state else 0)
