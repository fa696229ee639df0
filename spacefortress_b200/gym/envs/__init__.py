from .ssf_env import SSF_Env  # noqa: F401
