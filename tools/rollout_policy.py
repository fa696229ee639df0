"""BASELINE.json configs[2]: youturn, N envs, frame-stack 4, fully on-device T-step rollout with the SF-GRU policy.
Prints env-only and env+policy throughput.  usage: python tools/rollout_policy.py [N] [T]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spacefortress_b200 import SFVecEnv
from spacefortress_b200.rollout import OnDeviceRollout, SFGRUPolicy
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 128
env = SFVecEnv("youturn", num_envs=n, device=0)
policy = SFGRUPolicy(env.num_actions).cuda().eval()
if os.environ.get('SF_POLICY_DTYPE', 'bf16') == 'bf16':
    policy = policy.bfloat16().to(memory_format=torch.channels_last)
ro = OnDeviceRollout(env, policy, num_steps=T)
ro.collect(); torch.cuda.synchronize()
t0 = time.perf_counter(); ro.collect(); torch.cuda.synchronize(); t1 = time.perf_counter()
both = n * T / (t1 - t0)
# env only: the same T steps through the fused rollout kernel with recorded actions
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
obs = ro.frames[4:4 + T].unsqueeze(2)
s.record(); env.rollout(T, actions=ro.actions, out={"obs": obs.contiguous()}); e.record(); torch.cuda.synchronize()
envonly = n * T / (s.elapsed_time(e) * 1e-3)
print(json.dumps({"config": "youturn, %d envs, stack 4, %d-step on-device rollout, SF-GRU" % (n, T), "env_plus_policy_steps_per_s": both,
                  "env_only_steps_per_s": envonly, "frames_buffer_GB": ro.frames.numel() / 1e9, "fort_kills": int(ro.num_destruction)}))
