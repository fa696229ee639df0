"""On-device rollout collection (BASELINE.json configs[2]): the env, the frame
history and the policy all stay on the GPU; nothing crosses PCIe inside the loop.

Replaces the host ping-pong of rl/train.py:73-98 (GPU->CPU actions :79, pipe per env :80, CPU->GPU obs :53-56,97).
Observations are kept as SINGLE frames in a time-major buffer [T+3+1, N, 84, 84] u8 — the reference stores the
whole 4-stack per step as float (rl/storage.py:11,37: 238 GB at 65 536 envs x 128 steps; single u8 frames are
59 GB). The 4-stack is a window over that buffer; frames older than the last `done` are zeroed exactly like
`current_obs *= mask` does (rl/train.py:92-95).

The policy is a consumer, kept in plain PyTorch (dense conv/GEMM work -> cuDNN/cuBLAS, not a hand-written
kernel): SFGRUPolicy has the layer shapes of the reference's ACNet (rl/networks.py:20-76).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class SFGRUPolicy(nn.Module):
    """conv(4->16,k8,s4) -> conv(16->32,k4,s2) -> fc 2592->256 -> GRUCell(256) -> action / value heads
    (rl/networks.py:23-33); orthogonal weights, zero biases (rl/networks.py:6-18)."""

    def __init__(self, num_actions, feedforward=False):
        super().__init__()
        self.feedforward = feedforward
        self.conv1 = nn.Conv2d(4, 16, kernel_size=8, stride=4)
        self.conv2 = nn.Conv2d(16, 32, kernel_size=4, stride=2)
        self.fc1 = nn.Linear(2592, 256)
        # module names as in the reference's ACNet (rl/networks.py:27-30), so that its checkpoints
        # (`*_ppo_actor.pth.tar`, rl/train.py:150-152) load with load_state_dict and vice versa
        if feedforward:
            self.fc2 = nn.Linear(256, 256)
        else:
            self.gru = nn.GRUCell(256, 256)
        self.action = nn.Linear(256, num_actions)
        self.value = nn.Linear(256, 1)
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.orthogonal_(m.weight); nn.init.zeros_(m.bias)
            elif isinstance(m, nn.GRUCell):
                nn.init.orthogonal_(m.weight_ih); nn.init.orthogonal_(m.weight_hh)
                nn.init.zeros_(m.bias_ih); nn.init.zeros_(m.bias_hh)

    def conv1_s2d_weight(self):
        """conv1 (4->16, k8, s4) as a 2x2 convolution over the 64 channels of the space-to-depth input that
        sf_policy_input builds (channel = frame*16 + (y%4)*4 + (x%4)): the same sums, 10x faster in cuDNN."""
        w = self.conv1.weight
        return w.view(16, 4, 2, 4, 2, 4).permute(0, 1, 3, 5, 2, 4).reshape(16, 64, 2, 2).contiguous(memory_format=torch.channels_last)

    def features(self, obs_u8, state, mask, s2d=False):
        """obs_u8: [N,4,84,84] u8 stack, or with s2d=True the [N,64,21,21] channels_last input of sf_policy_input
        (already scaled by 1/255 in the policy's dtype)."""
        dt = self.conv1.weight.dtype  # fp32 like the reference, or bf16 via policy.bfloat16()
        if s2d:
            # channels_last bf16 all the way: cuDNN's fused conv + bias + relu, and fc1 on the NHWC-flattened
            # activations (weight columns permuted accordingly) instead of a layout copy
            x = torch.cudnn_convolution_relu(obs_u8, self.conv1_s2d_weight(), self.conv1.bias, (1, 1), (0, 0), (1, 1), 1)
            x = torch.cudnn_convolution_relu(x, self.conv2.weight.contiguous(memory_format=torch.channels_last), self.conv2.bias, (2, 2), (0, 0), (1, 1), 1)
            x = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)   # a view for channels_last: [N, 9*9*32]
            w = self.fc1.weight.view(256, 32, 9, 9).permute(0, 2, 3, 1).reshape(256, 2592)
            x = F.relu(F.linear(x, w, self.fc1.bias))
        else:
            x = F.relu(self.conv1(obs_u8.to(dt) / 255.0))
            x = F.relu(self.conv2(x)).flatten(1)
            x = F.relu(self.fc1(x))
        if self.feedforward:
            return F.relu(self.fc2(x)), state
        h = self.gru(x, (state * mask).to(x.dtype))
        return h, h

    def load_autoturn_model(self, actor_state):
        """rl/networks.py:88-106 (`--transfer`): copy a reference-shaped state_dict (same keys) into this policy."""
        assert actor_state.keys() == self.state_dict().keys(), "Keys not same!"
        self.load_state_dict(actor_state)

    @torch.no_grad()
    def act(self, obs_u8, state, mask, deterministic=False, s2d=False):
        x, state = self.features(obs_u8, state, mask, s2d=s2d)
        logits = self.action(x)
        logp = F.log_softmax(logits, dim=1)
        action = logits.argmax(1, keepdim=True) if deterministic else torch.multinomial(logp.float().exp(), 1)
        return self.value(x), action, logp.gather(1, action), state


class OnDeviceRollout(object):
    """Collects T-step rollouts of `env` (SFVecEnv) under `policy` entirely on the device.

    Buffer contract (what `PPOLearner.update` replays, rl/storage.py:37,103 stores observations[step] instead):
    after collect() and until the NEXT collect(), frames[t .. t+S-1] are the frames the policy saw at step t and
    valid_hist[t] says how many of them belong to the running episode (older ones read 0, like
    `current_obs *= masks` in rl/train.py:92-95), so stack(t) / policy_input(t) reproduce the acted-on input for every
    0 <= t <= T. The carry-over of the last S frames into the history slots happens at the START of the next
    collect(), never behind the learner's back. state0 / mask0 are the recurrent state and mask before step 0.

    graph=True captures one step (policy input -> act -> sf_step -> bookkeeping, all on static buffers indexed by a
    device-side step counter) into a CUDA graph per step index and replays the T graphs: no Python or launch gaps
    inside the loop (rl/train.py:73-98 has a host round trip per step)."""

    def __init__(self, env, policy, num_steps=128, num_stack=4, graph=False, fused_input=None):
        self.env, self.policy, self.T, self.S = env, policy, int(num_steps), int(num_stack)
        n, dev = env.num_envs, env._device()
        pdt = next(policy.parameters()).dtype
        self.frames = torch.zeros((self.T + self.S, n, 84, 84), dtype=torch.uint8, device=dev)  # [S-1 history | T+1]
        self.valid_hist = torch.zeros((self.T + 1, n), dtype=torch.int32, device=dev)  # frames since the last reset, capped at S, per step
        self.state = torch.zeros(n, 256, device=dev, dtype=pdt)
        self.mask = torch.ones(n, 1, device=dev, dtype=pdt)
        self.state0 = torch.zeros_like(self.state)
        self.mask0 = torch.ones_like(self.mask)
        # hidden state BEFORE every step (rl/storage.py:12,38: the learner re-evaluates each sample from its stored state)
        self.states_hist = torch.zeros((self.T + 1, n, 256), device=dev, dtype=pdt) if not policy.feedforward else None
        self.actions = torch.zeros((self.T, n), dtype=torch.int32, device=dev)
        self.rewards = torch.zeros((self.T, n), dtype=torch.int32, device=dev)
        self.dones = torch.zeros((self.T, n), dtype=torch.bool, device=dev)
        self.values = torch.zeros((self.T, n), device=dev)
        self.logps = torch.zeros((self.T, n), device=dev)
        self.episode_return = torch.zeros(n, dtype=torch.int64, device=dev)
        self.final_return = torch.zeros(n, dtype=torch.int64, device=dev)
        self.num_destruction = torch.zeros((), dtype=torch.int64, device=dev)
        self._age_idx = torch.arange(self.S, device=dev, dtype=torch.int32).view(1, self.S, 1, 1)
        # the policy's first-layer input comes from the fused stack / mask / scale / space-to-depth kernel (bf16 or fp32)
        self.fused_input = pdt in (torch.bfloat16, torch.float32) and self.S == 4 and dev.type == "cuda"
        if fused_input is not None:   # False: hand the policy the plain [N,4,84,84] u8 stack (what the reference's ACNet takes)
            self.fused_input = self.fused_input and bool(fused_input)
        if self.fused_input:
            self.pin = torch.empty((n, 64, 21, 21), dtype=pdt, device=dev).contiguous(memory_format=torch.channels_last)
        self.use_graph = bool(graph)
        self._graphs = None
        self._collected = False
        first = env.reset(to_numpy=False)
        self.frames[self.S - 1].copy_(first[:, 0])
        self.valid_hist[0].fill_(1)

    @property
    def valid(self):
        """frames of the running episode in the newest stack (after collect(): the one ending at frame T)"""
        return self.valid_hist[self.T if self._collected else 0]

    def stack(self, t):
        """[N,S,84,84] u8 window the policy saw at step t (buffer frames t .. t+S-1); frames from before the last
        reset read 0."""
        w = self.frames[t:t + self.S].permute(1, 0, 2, 3)
        keep = self._age_idx >= (self.S - self.valid_hist[t]).view(-1, 1, 1, 1)
        return w * keep.to(torch.uint8)

    def policy_input(self, t):
        """The 4-frame stack of step t as the policy's space-to-depth bf16 input (one fused kernel)."""
        from . import _lib
        import ctypes as C
        fn = self.env.L.sf_policy_input if self.pin.dtype == torch.bfloat16 else self.env.L.sf_policy_input_f32
        _lib.check(fn(C.c_void_p(self.frames[t].data_ptr()), int(self.frames.stride(0)), self.env.num_envs,
                      C.c_void_p(self.valid_hist[t].data_ptr()), C.c_void_p(self.pin.data_ptr()), self.env._stream_ptr()))
        return self.pin

    def _step(self, t):
        env, S = self.env, self.S
        if self.states_hist is not None:
            self.states_hist[t].copy_(self.state)
        if self.fused_input:
            value, action, logp, state = self.policy.act(self.policy_input(t), self.state, self.mask, s2d=True)
        else:
            value, action, logp, state = self.policy.act(self.stack(t), self.state, self.mask)
        self.state.copy_(state)
        a = action.squeeze(1).to(torch.int32)
        _, reward, done, kill = env.step(a, out_obs=self.frames[t + S].unsqueeze(1))
        self.actions[t].copy_(a); self.rewards[t].copy_(reward); self.dones[t].copy_(done)
        self.values[t].copy_(value.squeeze(1).float()); self.logps[t].copy_(logp.squeeze(1).float())
        self.num_destruction.add_(kill.sum())
        self.episode_return.add_(reward)
        self.final_return.copy_(torch.where(done, self.episode_return, self.final_return))
        self.episode_return.masked_fill_(done, 0)
        self.mask.copy_((~done).to(self.state.dtype).unsqueeze(1))
        v = self.valid_hist[t]
        self.valid_hist[t + 1].copy_(torch.where(done, torch.ones_like(v), torch.clamp(v + 1, max=S)))

    def _begin(self):
        S, T = self.S, self.T
        if self._collected:  # carry the last S frames (and their validity) over as the history of this rollout
            self.frames[:S].copy_(self.frames[T:T + S].clone())
            self.valid_hist[0].copy_(self.valid_hist[T].clone())
        self.state0.copy_(self.state); self.mask0.copy_(self.mask)

    def collect(self):
        self._begin()
        if self.use_graph:
            self._collect_graphed()
        else:
            for t in range(self.T):
                self._step(t)
        if self.states_hist is not None:
            self.states_hist[self.T].copy_(self.state)
        self._collected = True
        self.env._device_work = True  # graph replays step the env on torch's stream without going through its methods
        return dict(frames=self.frames, actions=self.actions, rewards=self.rewards, dones=self.dones,
                    values=self.values, logps=self.logps, valid=self.valid_hist)

    def _collect_graphed(self):
        """One CUDA graph per step index (every step reads / writes a different slice of the static buffers), captured
        once into ONE shared memory pool (they replay in capture order, never concurrently), replayed back to back."""
        if self._graphs is None:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream(device=self.env._device())
            side.wait_stream(cur)
            with torch.cuda.stream(side):  # warm-up outside capture (cuDNN / cuBLAS handles, workspaces); no state is changed
                for _ in range(2):
                    if self.fused_input:
                        self.policy.act(self.policy_input(0), self.state, self.mask, s2d=True)
                    else:
                        self.policy.act(self.stack(0), self.state, self.mask)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            pool = torch.cuda.graph_pool_handle()
            graphs, t_env = [], self.env._t
            for t in range(self.T):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    self._step(t)
                graphs.append(g)
            self.env._t = t_env  # capture records, it does not run: the env has not moved
            self._graphs = graphs
        for g in self._graphs:
            g.replay()
        self.env._t += self.T
