"""Experiment: time the render-only kernel (no step code in it) on a steady-state batch."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spacefortress_b200 import SFVecEnv, _lib
for n in (4096, 16384, 65536):
    env = SFVecEnv("autoturn", num_envs=n, device=0); env.reset(to_numpy=False)
    env.rollout(300, want=("reward",))
    o = torch.empty((n, 84, 84), dtype=torch.uint8, device="cuda")
    env.rollout(2, want=("obs",))  # builds the explosion sprites
    L = env.L
    for _ in range(3): _lib.check(L.sf_render(env.h, C.c_void_p(o.data_ptr()), 0, None))
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    R = 20
    s.record()
    for _ in range(R): _lib.check(L.sf_render(env.h, C.c_void_p(o.data_ptr()), 0, None))
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / R
    print("render-only n=%d: %.3f ms per frame batch -> %.3e frames/s" % (n, ms, n / ms * 1e3), flush=True)
    env.close()
