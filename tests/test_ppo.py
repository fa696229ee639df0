"""CPU: the learner-side PPO pieces (spacefortress_b200/ppo.py) against golden vectors generated from the reference's
own rl/storage.py (tests/golden/make_ppo_golden.py) and against the formulas of rl/train.py:108-129."""
import os
import types

import numpy as np
import torch

from spacefortress_b200 import ppo
from spacefortress_b200.rollout import SFGRUPolicy

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ppo_returns.npz"))


def _case(c):
    t = lambda k: torch.from_numpy(GOLD["c%d_%s" % (c, k)])
    return t("rewards"), t("values"), t("masks"), t("next")


def test_returns_match_the_reference_storage_bit_for_bit():
    for c in range(3):
        rewards, values, masks, nxt = _case(c)
        for gae, key in ((True, "gae"), (False, "mc")):
            got = ppo.compute_returns(rewards, values, masks, nxt, use_gae=gae, gamma=0.99, tau=0.95)
            assert np.array_equal(got.numpy(), GOLD["c%d_returns_%s" % (c, key)]), (c, key)


def test_advantage_normalisation_matches_the_reference():
    for c in range(3):
        rewards, values, masks, nxt = _case(c)
        ret = ppo.compute_returns(rewards, values, masks, nxt)
        adv = ppo.normalized_advantages(ret, values)
        assert np.allclose(adv.numpy(), GOLD["c%d_adv" % c], rtol=0, atol=1e-6), c


def test_ppo_loss_is_the_clipped_surrogate_of_the_reference():
    g = torch.Generator().manual_seed(1)
    v, ret, alp, old, adv = (torch.randn(64, generator=g) for _ in range(5))
    ent = torch.tensor(0.7)
    total, al, vl = ppo.ppo_loss(v, alp, ent, ret, old, adv, clip_param=0.1, value_loss_coeff=0.5, entropy_coeff=0.05)
    ratio = torch.exp(alp - old)                                   # rl/train.py:122-129, restated
    exp_al = -torch.min(ratio * adv, torch.clamp(ratio, 0.9, 1.1) * adv).mean()
    exp_vl = (v - ret).pow(2).mean()
    assert torch.equal(al, exp_al) and torch.equal(vl, exp_vl) and torch.equal(total, exp_al + 0.5 * exp_vl - 0.05 * ent)


def test_sequence_evaluation_equals_step_by_step_acting():
    torch.manual_seed(0)
    T, N = 5, 3
    policy = SFGRUPolicy(5).eval()
    obs = torch.randint(0, 256, (T, N, 4, 84, 84), dtype=torch.uint8)
    masks = (torch.rand(T, N, 1) > 0.3).float()
    state = torch.randn(N, 256)
    acts, vals, lps, s = [], [], [], state
    with torch.no_grad():
        for t in range(T):
            v, a, lp, s = policy.act(obs[t], s, masks[t], deterministic=True)
            acts.append(a.squeeze(1)); vals.append(v.squeeze(1)); lps.append(lp.squeeze(1))
        v2, lp2, ent = ppo.evaluate_actions(policy, obs, state, masks, torch.stack(acts, 0), bptt=True)
    assert torch.allclose(v2, torch.stack(vals, 0), atol=1e-5) and torch.allclose(lp2, torch.stack(lps, 0), atol=1e-5)
    assert 0.0 <= float(ent) <= float(np.log(5)) + 1e-5


def test_learner_update_runs_on_a_rollout_and_changes_the_policy():
    torch.manual_seed(0)
    T, N = 4, 8
    policy = SFGRUPolicy(3)
    frames = torch.randint(0, 256, (T + 4, N, 84, 84), dtype=torch.uint8)
    ro = types.SimpleNamespace(rewards=torch.randint(-1, 2, (T, N), dtype=torch.int32), dones=torch.rand(T, N) > 0.8, values=torch.randn(T, N),
                               logps=-torch.rand(T, N), actions=torch.randint(0, 3, (T, N), dtype=torch.int32),
                               stack=lambda t: frames[t:t + 4].permute(1, 0, 2, 3), states_hist=torch.randn(T + 1, N, 256) * 0.1)
    before = [p.detach().clone() for p in policy.parameters()]
    learner = ppo.PPOLearner(policy, ppo_epoch=1, num_mini_batch=2)
    stats = learner.update(ro, torch.zeros(N, 256), torch.ones(N, 1), torch.randn(N))
    assert len(stats) == 2 and all(np.isfinite(s).all() for s in stats)
    assert any(not torch.equal(a, b) for a, b in zip(before, policy.parameters()))


# ---- one PPO minibatch against the reference's own code (tests/golden/make_ppo_minibatch_golden.py) -------------------
MB = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ppo_minibatch.npz"))


def formula_weights(state_dict):
    """the closed-form weights the golden generator loaded into the reference's ACNet"""
    out = {}
    for i, (k, v) in enumerate(sorted(state_dict.items())):
        n = v.numel()
        w = torch.sin(torch.arange(n, dtype=torch.float64) * (0.37 + 0.11 * i) + i) * (1.5 / max(v.shape[-1] if v.dim() > 1 else 8, 8) ** 0.5)
        out[k] = w.reshape(v.shape).float()
    return out


def minibatch_loss_and_grads(device):
    """The reference's first minibatch (rl/train.py:110-130 with rl/storage.py:89-121's env order) through ppo.py."""
    T, N, A = int(MB["T"]), int(MB["N"]), int(MB["A"])
    policy = SFGRUPolicy(A)
    assert sorted(policy.state_dict().keys()) == sorted(["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "fc1.weight", "fc1.bias",
                                                         "gru.weight_ih", "gru.weight_hh", "gru.bias_ih", "gru.bias_hh", "action.weight", "action.bias",
                                                         "value.weight", "value.bias"])  # the reference's checkpoint keys (rl/networks.py:23-33)
    policy.load_autoturn_model(formula_weights(policy.state_dict()))
    policy = policy.to(device)
    t = lambda k, dt=None: torch.from_numpy(MB[k]).to(device) if dt is None else torch.from_numpy(MB[k]).to(device=device, dtype=dt)
    rewards, values, masks, nxt = t("rewards"), t("values"), t("masks"), t("next_value")
    returns = ppo.compute_returns(rewards, values, masks, nxt, True, 0.99, 0.95)
    adv = ppo.normalized_advantages(returns, values)
    idx = torch.from_numpy(MB["env_order"]).to(device)
    obs, states = t("obs_u8")[:T][:, idx], t("states")[:T][:, idx]
    v, alp, ent = ppo.evaluate_actions(policy, obs, states, masks[:T, idx].unsqueeze(-1), t("actions")[:, idx].long())
    loss, al, vl = ppo.ppo_loss(v, alp, ent, returns[:, idx], t("old_logp")[:, idx], adv[:, idx], 0.1, 0.5, 0.05)
    policy.zero_grad()
    loss.backward()
    return returns, adv, (loss, al, vl, ent), {k: p.grad.detach().cpu() for k, p in policy.named_parameters()}


def check_minibatch(device, rtol, gtol):
    returns, adv, (loss, al, vl, ent), grads = minibatch_loss_and_grads(device)
    assert np.allclose(returns.cpu().numpy(), MB["returns"], rtol=0, atol=1e-5) and np.allclose(adv.cpu().numpy(), MB["adv"], rtol=0, atol=1e-4)
    for got, key in ((loss, "loss"), (al, "action_loss"), (vl, "value_loss"), (ent, "entropy")):
        assert abs(float(got) - float(MB[key])) <= rtol * max(1.0, abs(float(MB[key]))), (key, float(got), float(MB[key]))
    for k, g in grads.items():
        flat = g.reshape(-1)
        ref_n = float(MB["gn_" + k])
        assert abs(float(flat.double().norm()) - ref_n) <= gtol * max(ref_n, 1e-3), (k, float(flat.double().norm()), ref_n)
        sample = flat[::max(1, flat.numel() // 256)][:256].numpy()
        assert np.allclose(sample, MB["gx_" + k], rtol=0, atol=gtol * max(ref_n, 1e-3)), k


def test_one_ppo_minibatch_matches_the_reference_code_on_cpu():
    """loss terms and every parameter's gradient of the reference's first minibatch, fp32 on the CPU (same kernels as the
    generator: tight tolerances)."""
    check_minibatch(torch.device("cpu"), 1e-5, 1e-4)
