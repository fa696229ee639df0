"""Registry of the reference's four env ids (python/spacefortress.gym/spacefortress/gym/__init__.py:3-29).
`gym` itself is not installed offline; `make(id)` gives the single-env facade, and if the real gym package
is importable the ids are registered there as well."""
from .envs import SSF_Env

REGISTRY = {
    "SpaceFortress-youturn-image-v0": dict(gametype="youturn", obs_type="image"),
    "SpaceFortress-autoturn-image-v0": dict(gametype="autoturn", obs_type="image"),
    "SpaceFortress-testyouturn-image-v0": dict(gametype="test-youturn", obs_type="image"),
    "SpaceFortress-testautoturn-image-v0": dict(gametype="test-autoturn", obs_type="image"),
}


def make(env_id, **kwargs):
    if env_id not in REGISTRY:
        raise KeyError("No registered env with id: %s" % env_id)
    kw = dict(REGISTRY[env_id])
    kw.update(kwargs)
    return SSF_Env(**kw)


try:  # pragma: no cover
    from gym.envs.registration import register as _register
    for _id, _kw in REGISTRY.items():
        try:
            _register(id=_id, entry_point="spacefortress_b200.gym.envs:SSF_Env", kwargs=_kw, nondeterministic=False)
        except Exception:
            pass
except Exception:
    pass
