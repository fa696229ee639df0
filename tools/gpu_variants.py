"""Build variants of the library with different -D knobs and time the 4096-env rollout for each (GPU box)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from spacefortress_b200 import SFVecEnv
n=int(sys.argv[1]); T=int(sys.argv[2]); gt=sys.argv[3]
env = SFVecEnv(gt, num_envs=n, device=0); env.reset(to_numpy=False)
env.rollout(300, want=("reward",))
out = {"obs": torch.empty((T, n, 1, 84, 84), dtype=torch.uint8, device="cuda")}
env.rollout(T, out=out); torch.cuda.synchronize()
best=1e9
for _ in range(3):
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record(); env.rollout(T, out=out); e.record(); torch.cuda.synchronize(); best=min(best,s.elapsed_time(e))
print("%s n=%d E=%s %s: %.3e steps/s" % (gt, n, os.environ.get("SF_ENVS_PER_WARP","auto"), os.environ.get("SF_NVCC_DEFS",""), n*T/best*1e3), flush=True)
'''
variants = [v for v in sys.argv[1:]] or [""]
for v in variants:
    defs, _, runenv = v.partition("|")
    env = dict(os.environ, SF_NVCC_DEFS=defs)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "spacefortress_b200", "build.py"), "--force"], env=env, capture_output=True, text=True)
    if r.returncode: print("BUILD FAILED", defs, r.stderr[-500:]); continue
    for kv in runenv.split():
        k, _, val = kv.partition("="); env[k] = val
    for n, T in ((4096, 64), (65536, 16)):
        subprocess.run([sys.executable, "-c", code, str(n), str(T), "autoturn"], env=env, cwd=ROOT)
subprocess.run([sys.executable, os.path.join(ROOT, "spacefortress_b200", "build.py"), "--force"], env=dict(os.environ, SF_NVCC_DEFS=""))
