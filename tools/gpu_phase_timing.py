"""Debug build with -DSF_PHASE_TIMING: cycles block 0 spends per phase (thread 0: barrier to barrier; all warps: busy vs waiting)."""
import os, subprocess, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
subprocess.run([sys.executable, os.path.join(ROOT, "spacefortress_b200", "build.py"), "--force"], env=dict(os.environ, SF_NVCC_DEFS="-DSF_PHASE_TIMING " + (sys.argv[2] if len(sys.argv) > 2 else "")), check=True)
import torch
from spacefortress_b200 import SFVecEnv, _lib
n, T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 64
env = SFVecEnv(os.environ.get("SF_GT", "autoturn"), num_envs=n, device=0); env.reset(to_numpy=False)
env.rollout(300, want=("reward",))
out = {"obs": torch.empty((T, n, 1, 84, 84), dtype=torch.uint8, device="cuda")}
env.rollout(T, out=out); torch.cuda.synchronize()
L = _lib.lib()
buf = (C.c_ulonglong * 96)()
L.sf_debug_cycles.restype = C.c_int; L.sf_debug_cycles.argtypes = [C.c_void_p, C.c_int]
L.sf_debug_cycles(buf, 1)
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
s.record(); env.rollout(T, out=out); e.record(); torch.cuda.synchronize()
L.sf_debug_cycles(buf, 1)
v = list(buf)
print("launch %.3f ms, %d ticks; block 0, cycles per tick:" % (s.elapsed_time(e), T))
names = ["step+scan (wait for warp 0)", "A env tasks", "B strokes", "C windows"]
for k in range(4): print("  thread0 phase %-28s %8.0f" % (names[k], v[k] / T))
print("  sum %8.0f cycles per tick = %.1f us" % (sum(v[:4]) / T, sum(v[:4]) / T / 1.965e3))
import re
m = re.search(r"SF_RENDER_WARPS=(\d+)", sys.argv[2] if len(sys.argv) > 2 else "")
W = int(m.group(1)) if m else 16
print("  all warps: waiting at barriers (+frame_end/step) %8.0f per warp-tick; busy A %6.0f  B %6.0f  C %6.0f" % (v[8] / T / W, v[9] / T / W, v[10] / T / W, v[11] / T / W))
sec = {16: "B geometry (xform, stroke quad, edges)", 17: "B open_regions", 18: "B publish_quads", 19: "B clear + phase 1 (spans)", 20: "B phase 2 (union + emit)", 21: "B arcs geometry", 22: "C patch init", 23: "C region test", 24: "C ship layer / explosion", 25: "C fortress layer", 26: "C projectile blends", 27: "C text + bar", 28: "C window_out", 29: "C task fetch", 31: "(other)", 64: "B stroke fetch", 65: "B bulk-copy wait", 66: "B base patch", 69: "B arcs / after last fetch", 70: "C run-ahead step (warp 0 only, /W)"}
print("  batches per tick %.1f, groups per batch %.1f" % (v[67] / T, v[68] / max(v[67], 1)))
print("  rounds per tick %.2f; slowest warp per round: B %.0f  C %.0f; tasks per round %.1f (env tasks %.1f), builds per round %.2f" % (v[73] / T, v[71] / max(v[73], 1), v[72] / max(v[73], 1), v[74] / max(v[73], 1), v[75] / max(v[73], 1), v[76] / max(v[73], 1)))
print("  phase B busy cycles per tick, by warp:", " ".join("%.0f" % (v[32 + k] / T) for k in range(min(W, 16))))
print("  phase C busy cycles per tick, by warp:", " ".join("%.0f" % (v[48 + k] / T) for k in range(min(W, 16))))
print("  sections, cycles per warp-tick (sum over the 16 warps / 16):")
for k in sorted(sec): print("    %-40s %8.0f" % (sec[k], v[k] / T / W))
subprocess.run([sys.executable, os.path.join(ROOT, "spacefortress_b200", "build.py"), "--force"], env=dict(os.environ, SF_NVCC_DEFS=""))
