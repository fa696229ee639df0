"""Long parity run of the FUSED rollout against the oracle: every frame and reward of `n` envs over `steps` ticks
(chunks of T ticks per launch, random actions, different seeds), for both training game types.
usage: python tools/gpu_long_parity.py [n] [steps] [T]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spacefortress_b200 import SFVecEnv
from oracle.oracle import OracleEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
T = int(sys.argv[3]) if len(sys.argv) > 3 else 50
bad = 0
for gt in ("autoturn", "youturn"):
    seeds = np.arange(1, n + 1)
    env = SFVecEnv(gt, num_envs=n, device=0, seeds=seeds); env.reset()
    orc = [OracleEnv(gt, int(s)) for s in seeds]
    rng = np.random.RandomState(5)
    t0 = time.time(); frames = 0; worst = 0
    for c in range(0, steps, T):
        acts = rng.randint(0, env.num_actions, size=(T, n)).astype(np.int32)
        out = env.rollout(T, actions=torch.from_numpy(acts).cuda())
        obs = out["obs"].cpu().numpy(); rew = out["reward"].cpu().numpy(); done = out["done"].cpu().numpy()
        for t in range(T):
            for i in range(n):
                r, d, k, _ = orc[i].step(orc[i].keymask(int(acts[t, i])))
                if d: orc[i].reset()
                if r != int(rew[t, i]) or bool(d) != bool(done[t, i]):
                    bad += 1; print("step mismatch", gt, c + t, i)
                o = orc[i].obs()
                if not np.array_equal(o, obs[t, i, 0]):
                    bad += 1; worst = max(worst, int(np.abs(o.astype(int) - obs[t, i, 0]).max()))
                    if bad < 10: print("frame mismatch", gt, c + t, i, int((o != obs[t, i, 0]).sum()), "px")
                frames += 1
    print("%s: %d frames + rewards compared in %.0f s, mismatches so far %d (max |delta| %d)" % (gt, frames, time.time() - t0, bad, worst), flush=True)
    env.close()
sys.exit(1 if bad else 0)
