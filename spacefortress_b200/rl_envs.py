"""Counterpart of the reference's rl/envs.py: make_env(env_id, seed, rank) and WrapPyTorch.

`make_env` returns a thunk exactly like rl/envs.py:10-16, so `SubprocVecEnv([make_env(id, seed, i) for i
in range(N)])` (rl/train.py:30-34) runs unchanged with `from spacefortress_b200 import make_env,
SubprocVecEnv`. Calling the thunk builds a single-env SSF_Env wrapped by WrapPyTorch."""
import numpy as np

from .spaces import Box
from .vec_env import EnvThunk


def make_env(env_id, seed, rank):
    return EnvThunk(env_id, seed, rank)


class WrapPyTorch(object):
    """rl/envs.py:19-30: observation -> cv2.resize(obs,(84,84),INTER_AREA)[None]. The resize runs on the GPU
    inside the same kernel that draws the frame; this wrapper only asks the env for the 84x84 output."""

    def __init__(self, env=None):
        self.env = env
        self.observation_space = Box(0, 255, [1, 84, 84], dtype=np.uint8)
        self.action_space = env.action_space
        self.metadata = getattr(env, "metadata", {})

    def observation(self, observation=None):
        return self.env.obs84()[None]

    def reset(self):
        self.env.reset()
        return self.observation()

    def step(self, action):
        _, reward, done, info = self.env.step(action)
        return self.observation(), reward, done, info

    def seed(self, seed=None):
        return self.env.seed(seed)

    def render(self, mode="human", close=False):
        return self.env.render(mode, close)

    def close(self):
        return self.env.close()

    def __getattr__(self, name):
        return getattr(self.env, name)
