"""Learner side of the reference's PPO loop, on the device: generalised advantage
estimation, advantage normalisation, the clipped-surrogate loss and the recurrent minibatch update.

Mirrors rl/storage.py:50-62 (`RolloutStorage.compute_returns`), rl/train.py:105-133 (`PPOAgent.train`, the update
part) and rl/networks.py:41-86 (`ACNet._fwd` / `evaluate_actions` over T x N sequences). Plain PyTorch: this is
dense, differentiable work for cuBLAS/cuDNN/autograd, not a hand-written kernel; what it consumes (frames, actions,
rewards, dones, values, log-probs of a whole rollout) is already on the device (`rollout.OnDeviceRollout`), so
nothing crosses PCIe between collecting a rollout and updating on it.
"""
import torch
import torch.nn.functional as F


def compute_returns(rewards, value_preds, masks, next_value, use_gae=True, gamma=0.99, tau=0.95):
    """rl/storage.py:50-62. rewards [T,N]; value_preds [T,N] (values of the states acted on); masks [T+1,N]
    (masks[t+1] == 0 where the step t ended an episode); next_value [N]. Returns `returns` [T,N].
    Same operations in the same order as the reference (a reversed scan over T of N-wide tensor ops)."""
    T = rewards.shape[0]
    returns = torch.empty_like(rewards)
    if use_gae:
        vp = torch.cat([value_preds, next_value.unsqueeze(0)], 0)
        gae = torch.zeros_like(next_value)
        for step in reversed(range(T)):
            delta = rewards[step] + gamma * vp[step + 1] * masks[step + 1] - vp[step]
            gae = delta + gamma * tau * masks[step + 1] * gae
            returns[step] = gae + vp[step]
    else:
        nxt = next_value
        for step in reversed(range(T)):
            nxt = nxt * gamma * masks[step + 1] + rewards[step]
            returns[step] = nxt
    return returns


def normalized_advantages(returns, value_preds):
    """rl/train.py:108-109: (A - mean) / (std + 1e-5), unbiased std over all T*N entries."""
    adv = returns - value_preds
    return (adv - adv.mean()) / (adv.std() + 1e-5)


def ppo_loss(values, action_log_probs, dist_entropy, returns, old_action_log_probs, adv_targ, clip_param=0.1, value_loss_coeff=0.5, entropy_coeff=0.05):
    """rl/train.py:122-129. Returns (total, action_loss, value_loss)."""
    ratio = torch.exp(action_log_probs - old_action_log_probs)
    surr1 = ratio * adv_targ
    surr2 = torch.clamp(ratio, 1.0 - clip_param, 1.0 + clip_param) * adv_targ
    action_loss = -torch.min(surr1, surr2).mean()
    value_loss = (values - returns).pow(2).mean()
    return action_loss + value_loss_coeff * value_loss - entropy_coeff * dist_entropy, action_loss, value_loss


def evaluate_actions(policy, obs_u8, states, masks, actions, bptt=False):
    """rl/networks.py:41-62,79-86 on a [T,N] minibatch: obs_u8 [T,N,4,84,84] u8 stacks, masks [T,N,1] (0 where the
    PREVIOUS step ended an episode), actions [T,N] int64. Returns values [T,N], action_log_probs [T,N], mean entropy.

    bptt=False (the reference): `states` [T,N,256] are the hidden states the policy had BEFORE each step, stored during
    the rollout. The reference's recurrent_generator (rl/storage.py:89-121) hands ACNet.evaluate_actions one stored state
    per sample, so ACNet._fwd takes its `obs.size(0) == state.size(0)` branch (rl/networks.py:47-48): ONE GRU step per
    sample from the stored state, no unrolling through time.
    bptt=True (extension): `states` [N,256] is the state before the first step and the GRU is unrolled over the T steps
    (the other branch of ACNet._fwd), which back-propagates through time."""
    T, N = actions.shape
    dt = policy.conv1.weight.dtype
    x = F.relu(policy.conv1(obs_u8.reshape(T * N, *obs_u8.shape[2:]).to(dt) / 255.0))
    x = F.relu(policy.conv2(x)).flatten(1)
    x = F.relu(policy.fc1(x))
    if policy.feedforward:
        h = F.relu(policy.fc2(x)).view(T, N, -1)
    elif not bptt:
        h = policy.gru(x, states.reshape(T * N, -1).to(dt) * masks.reshape(T * N, 1).to(dt)).view(T, N, -1)
    else:
        x = x.view(T, N, -1)
        hs, state = [], states.to(dt)
        for t in range(T):
            state = policy.gru(x[t], state * masks[t].to(dt))
            hs.append(state)
        h = torch.stack(hs, 0)
    logits = policy.action(h).float()
    logp = F.log_softmax(logits, dim=-1)
    values = policy.value(h).float().squeeze(-1)
    alp = logp.gather(-1, actions.unsqueeze(-1)).squeeze(-1)
    entropy = -(logp * logp.exp()).sum(-1).mean()
    return values, alp, entropy


class PPOLearner(object):
    """The update half of rl/train.py's PPOAgent on a rollout collected by OnDeviceRollout (Adam lr 1e-3, 4 epochs,
    4 env-chunk minibatches, clip 0.1, value 0.5, entropy 0.05, grad-norm 0.5: rl/arguments.py defaults)."""

    def __init__(self, policy, lr=1e-3, gamma=0.99, tau=0.95, ppo_epoch=4, num_mini_batch=4, clip_param=0.1,
                 value_loss_coeff=0.5, entropy_coeff=0.05, max_grad_norm=0.5, bptt=False):
        self.policy = policy
        self.bptt = bool(bptt)
        self.opt = torch.optim.Adam(policy.parameters(), lr=lr)
        self.gamma, self.tau, self.ppo_epoch, self.num_mini_batch = gamma, tau, ppo_epoch, num_mini_batch
        self.clip_param, self.value_loss_coeff, self.entropy_coeff, self.max_grad_norm = clip_param, value_loss_coeff, entropy_coeff, max_grad_norm

    def update(self, ro, state0=None, mask0=None, next_value=None, env_chunk=None):
        """ro: an OnDeviceRollout after collect() (its buffers describe step t until the next collect(): stack(t) is
        the observation acted on at step t); state0 / mask0: hidden state and mask before its first step (default: the
        ones the rollout recorded); next_value [N]: value of the state after its last step (default: evaluated here,
        rl/train.py:100-103). env_chunk: evaluate a minibatch in slices of this many envs with gradient accumulation
        (same loss: the minibatch means are weighted by slice size) so that 65 536 envs x 128 steps fit in memory."""
        T, N = ro.rewards.shape
        if state0 is None:
            state0, mask0 = ro.state0, ro.mask0
        if next_value is None:
            with torch.no_grad():
                x, _ = self.policy.features(ro.stack(T), ro.state, ro.mask)
                next_value = self.policy.value(x).float().squeeze(1)
        done = ro.dones.float()
        masks = torch.cat([mask0.float().view(1, N), 1.0 - done], 0)           # masks[t+1] = 0 where step t ended an episode
        returns = compute_returns(ro.rewards.float(), ro.values, masks, next_value.float(), True, self.gamma, self.tau)
        adv = normalized_advantages(returns, ro.values)
        per = max(N // self.num_mini_batch, 1)
        stats = []
        for _ in range(self.ppo_epoch):
            perm = torch.randperm(N, device=ro.rewards.device)
            for s in range(0, N, per):                                          # rl/storage.py:95-121: whole env sequences per minibatch
                idx = perm[s:s + per]
                self.opt.zero_grad()
                acc = [0.0, 0.0, 0.0, 0.0]
                ck = len(idx) if not env_chunk else int(env_chunk)
                for c0 in range(0, len(idx), ck):
                    sub = idx[c0:c0 + ck]
                    w = len(sub) / float(len(idx))
                    obs = torch.stack([ro.stack(t)[sub] for t in range(T)], 0)
                    states = state0[sub] if self.bptt or self.policy.feedforward else ro.states_hist[:T, sub]
                    v, alp, ent = evaluate_actions(self.policy, obs, states, masks[:-1, sub].unsqueeze(-1), ro.actions[:, sub].long(), bptt=self.bptt)
                    loss, al, vl = ppo_loss(v, alp, ent, returns[:, sub], ro.logps[:, sub], adv[:, sub], self.clip_param, self.value_loss_coeff, self.entropy_coeff)
                    (loss * w).backward()
                    for k, x in enumerate((loss, al, vl, ent)):
                        acc[k] += w * float(x.detach())
                torch.nn.utils.clip_grad_norm_(self.policy.parameters(), self.max_grad_norm)
                self.opt.step()
                stats.append(tuple(acc))
        return stats
