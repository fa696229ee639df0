import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from spacefortress_b200 import SFVecEnv
from spacefortress_b200.rollout import OnDeviceRollout, SFGRUPolicy
n, T = 65536, 8
env = SFVecEnv("youturn", num_envs=n, device=0)
policy = SFGRUPolicy(env.num_actions).cuda().eval().bfloat16().to(memory_format=torch.channels_last)
ro = OnDeviceRollout(env, policy, num_steps=T)
ro.collect(); torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ro.collect(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
