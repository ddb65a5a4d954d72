"""shems_b200 — B200-native batched shems_LU1 environment + DDPG update behind a C ABI.

The importable name is `shems_b200` (see /shems_b200.py at the repo root, which registers this
directory — whose name is not a valid Python identifier — under that name).
"""
from . import _lib
from ._lib import (POLICY_RANDOM, POLICY_RULE, POLICY_TAPE, RESET_DETERMINISTIC, RESET_DEVICE_PHILOX, RESET_HOST_DRAWS,  # noqa: F401
                   ShemsBoundsError, ShemsError, ShemsKeyError, default_ddpg_params, params_for_charger, params_for_env, ENV_LU1, ENV_LU7,
                   ENV_LU1_INPUT0607)
from . import series  # noqa: F401
from . import tracker  # noqa: F401
from . import dataprep  # noqa: F401


def __getattr__(name):  # torch-dependent parts are imported lazily (the ABI/symbol tests run without CUDA)
    if name in ("Shems", "ShemsAction", "reset_", "step_", "action", "finished", "state", "actions"):
        from . import env
        return getattr(env, name)
    if name in ("Replay", "Learner", "Driver", "PopulationDriver"):
        from . import ddpg
        return getattr(ddpg, name)
    raise AttributeError(name)
