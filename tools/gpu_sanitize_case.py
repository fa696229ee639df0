"""A tiny case for compute-sanitizer (memcheck): every kernel of the library once, at sizes that exercise multi-group blocks,
the group hand-out, the bulk-copied tables, explosions, auto-reset and the feature / policy-input kernels.
usage: compute-sanitizer --tool memcheck python tools/gpu_sanitize_case.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spacefortress_b200 import SFVecEnv
for gt, n in (("autoturn", 70), ("youturn", 5000)):
    env = SFVecEnv(gt, num_envs=n, device=0)
    env.reset()
    env.set_ticks(np.random.RandomState(0).randint(5280, 5295, size=n))   # episode ends inside the run
    env.rollout(40, want=("reward",))
    out = env.rollout(6)
    o, r, d, k = env.step(np.zeros(n, np.int64))
    o2 = env.step(torch.ones(n, dtype=torch.int32, device="cuda"))
    f = env.features("normalized-features")
    st = env.episode_stats()
    fr = env.render_frames(native=True)
    torch.cuda.synchronize()
    print(gt, n, "ok", int(out["done"].sum()), st["episodes"])
    env.close()
print("sanitize case done")
