import os, sys, subprocess
code = r'''
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from spacefortress_b200 import SFVecEnv
n=int(sys.argv[1]); T=int(sys.argv[2]); gt=sys.argv[3]
env = SFVecEnv(gt, num_envs=n, device=0); env.reset(to_numpy=False)
env.rollout(200, want=("reward",))
out = {"obs": torch.empty((T, n, 1, 84, 84), dtype=torch.uint8, device="cuda")}
env.rollout(2, out={k: v[:2] for k, v in out.items()}); torch.cuda.synchronize()
best=1e9
for _ in range(3):
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record(); env.rollout(T, out=out); e.record(); torch.cuda.synchronize(); best=min(best,s.elapsed_time(e))
print("%s n=%d E=%s: %.3e steps/s" % (gt, n, os.environ.get("SF_ENVS_PER_WARP","auto"), n*T/best*1e3))
'''
for n, T, es in ((4096, 64, (1, 2, 4)), (65536, 16, (4, 8, 16, 32)), (262144, 8, (16, 32))):
    for e in es:
        env = dict(os.environ, SF_ENVS_PER_WARP=str(e))
        subprocess.run([sys.executable, "-c", code, str(n), str(T), "autoturn"], env=env)
