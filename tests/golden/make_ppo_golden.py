"""Generates tests/golden/ppo_returns.npz from the REFERENCE's own rl/storage.py (imported from /root/reference in
the build container; it needs nothing but torch): RolloutStorage.compute_returns with and without GAE on seeded
random rollouts, and the advantage normalisation of rl/train.py:108-109.  usage: python tests/golden/make_ppo_golden.py"""
import importlib.util, os, sys
import numpy as np, torch
spec = importlib.util.spec_from_file_location("ref_storage", "/root/reference/rl/storage.py")
ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
class _Discrete:  # RolloutStorage only looks at the class name
    pass
_Discrete.__name__ = "Discrete"
out = {}
g = torch.Generator().manual_seed(20261018)
for case, (T, N) in enumerate(((128, 16), (7, 5), (1, 3))):
    st = ref.RolloutStorage(T, N, (1, 2, 2), _Discrete(), 4)
    st.rewards.copy_(torch.randint(-1, 4, (T, N, 1), generator=g).float())
    st.value_preds[:-1].copy_(torch.randn(T, N, 1, generator=g))
    st.masks.copy_((torch.rand(T + 1, N, 1, generator=g) > 0.1).float())
    nv = torch.randn(N, 1, generator=g)
    out["c%d_rewards" % case] = st.rewards[..., 0].numpy().copy(); out["c%d_values" % case] = st.value_preds[:-1, :, 0].numpy().copy()
    out["c%d_masks" % case] = st.masks[..., 0].numpy().copy(); out["c%d_next" % case] = nv[:, 0].numpy().copy()
    for gae in (True, False):
        st.compute_returns(nv, gae, 0.99, 0.95)
        out["c%d_returns_%s" % (case, "gae" if gae else "mc")] = st.returns[:-1, :, 0].numpy().copy()
        if gae:
            adv = st.returns[:-1] - st.value_preds[:-1]
            adv = (adv - adv.mean()) / (adv.std() + 1e-5)
            out["c%d_adv" % case] = adv[..., 0].numpy().copy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ppo_returns.npz"), **out)
print("wrote", len(out), "arrays")
