"""spacefortress_b200 — B200-native batched Space Fortress simulator (env step + 84x84 render + auto-reset)
behind the reference's gym / gym_vecenv interface. Hot path: hand-written sm_100a kernels in
csrc/, reached only through the C-ABI of include/sf_b200.h (libsf_b200.so). No CPU fallback."""
from . import _lib  # noqa: F401
from .vec_env import SFVecEnv, SubprocVecEnv, DummyVecEnv, GAMETYPE_OF_ENV_ID  # noqa: F401
from .rl_envs import make_env, WrapPyTorch  # noqa: F401

__all__ = ["SFVecEnv", "SubprocVecEnv", "DummyVecEnv", "make_env", "WrapPyTorch", "GAMETYPE_OF_ENV_ID"]
