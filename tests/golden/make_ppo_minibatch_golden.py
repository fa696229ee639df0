"""Generates tests/golden/ppo_minibatch.npz from the REFERENCE's own code (imported from /root/reference in the build
container; both files need nothing but torch): rl/networks.py ACNet (SF-GRU) and rl/storage.py RolloutStorage are filled
with a seeded random rollout, then ONE minibatch of PPOAgent.train's update (rl/train.py:105-132: compute_returns with GAE,
advantage normalisation, recurrent_generator, ACNet.evaluate_actions, the clipped-surrogate loss, backward) is evaluated
in fp32 on the CPU. Saved: the weights, the rollout tensors, the minibatch env order, the three loss terms, the total
loss and the gradient of every parameter.  usage: python tests/golden/make_ppo_minibatch_golden.py"""
import importlib.util, os
import numpy as np, torch

def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m
storage = load("ref_storage", "/root/reference/rl/storage.py")
networks = load("ref_networks", "/root/reference/rl/networks.py")
class _Discrete:
    pass
_Discrete.__name__ = "Discrete"

torch.manual_seed(20261018)
T, N, A = 4, 6, 5
net = networks.ACNet(A, False)
def formula_weights(state_dict):
    """closed-form weights (no file, no RNG): the test rebuilds them with the same formula"""
    out = {}
    for i, (k, v) in enumerate(sorted(state_dict.items())):
        n = v.numel()
        w = torch.sin(torch.arange(n, dtype=torch.float64) * (0.37 + 0.11 * i) + i) * (1.5 / max(v.shape[-1] if v.dim() > 1 else 8, 8) ** 0.5)
        out[k] = w.reshape(v.shape).float()
    return out
net.load_state_dict(formula_weights(net.state_dict()))
st = storage.RolloutStorage(T, N, (4, 84, 84), _Discrete(), 256)
g = torch.Generator().manual_seed(7)
obs_u8 = torch.randint(0, 256, (T + 1, N, 4, 84, 84), generator=g, dtype=torch.uint8)
obs_u8 = obs_u8 * (torch.rand(T + 1, N, 4, 84, 84, generator=g) < 0.03).to(torch.uint8)   # frames are mostly black (small fixture)
st.observations.copy_(obs_u8.float())
st.states.copy_(torch.randn(T + 1, N, 256, generator=g) * 0.3)
st.masks.copy_((torch.rand(T + 1, N, 1, generator=g) > 0.15).float())
st.rewards.copy_(torch.randint(-1, 4, (T, N, 1), generator=g).float())
st.actions.copy_(torch.randint(0, A, (T, N, 1), generator=g))
with torch.no_grad():   # old values / log-probs come from the same net acting on the stored inputs (as in a real rollout)
    for t in range(T):
        v, lp, ent, _ = net.evaluate_actions(st.observations[t], st.states[t], st.masks[t], st.actions[t])
        st.value_preds[t].copy_(v); st.action_log_probs[t].copy_(lp + 0.05 * torch.randn(N, 1, generator=g))
    next_value = net.get_value(st.observations[-1], st.states[-1], st.masks[-1])
st.compute_returns(next_value, True, 0.99, 0.95)
adv = st.returns[:-1] - st.value_preds[:-1]
adv = (adv - adv.mean()) / (adv.std() + 1e-5)
torch.manual_seed(3)
perm_probe = torch.randperm(N)                           # what recurrent_generator will draw ...
torch.manual_seed(3)
sample = next(iter(st.recurrent_generator(adv, 2)))      # ... for its first minibatch of N / 2 envs
ob, sb, ab, rb, mb, olb, at = sample
values, alp, ent, _ = net.evaluate_actions(ob, sb, mb, ab)
ratio = torch.exp(alp - olb)
surr1, surr2 = ratio * at, torch.clamp(ratio, 1.0 - 0.1, 1.0 + 0.1) * at
action_loss = -torch.min(surr1, surr2).mean()
value_loss = (values - rb).pow(2).mean()
loss = action_loss + 0.5 * value_loss - 0.05 * ent
net.zero_grad()
loss.backward()
out = {"T": T, "N": N, "A": A, "env_order": perm_probe[:N // 2].numpy(),
       "obs_u8": obs_u8.numpy(), "states": st.states.numpy(), "masks": st.masks[..., 0].numpy(), "rewards": st.rewards[..., 0].numpy(),
       "actions": st.actions[..., 0].numpy(), "old_logp": st.action_log_probs[..., 0].numpy(), "values": st.value_preds[:-1, :, 0].numpy(),
       "next_value": next_value[:, 0].numpy(), "returns": st.returns[:-1, :, 0].numpy(), "adv": adv[..., 0].numpy(),
       "loss": loss.item(), "action_loss": action_loss.item(), "value_loss": value_loss.item(), "entropy": ent.item()}
for k, p in net.named_parameters():   # gradients: norm, sum and a strided sample of every parameter's
    gflat = p.grad.reshape(-1)
    out["gn_" + k] = gflat.double().norm().item(); out["gs_" + k] = gflat.double().sum().item()
    out["gx_" + k] = gflat[::max(1, gflat.numel() // 256)][:256].numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ppo_minibatch.npz"), **out)
print("wrote", len(out), "arrays; loss", loss.item(), action_loss.item(), value_loss.item(), ent.item())
