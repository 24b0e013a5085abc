"""gym_craftingworld_b200 -- B200-native batched CraftingWorld (the env hot path of lauradarcy/gym-craftingworld).

Importing the package builds (if stale) and loads ``libcw_b200.so``; there is no CPU fallback -- if the CUDA
library cannot be built or loaded the import raises.
"""
from . import _lib

_lib.load()   # fail loudly at import time

from .env import (ACTION_NAMES, COLORS_N, MAX_STEPS, OBJECTS, PICKUPABLE, TASK_LIST,  # noqa: E402
                  BatchedCraftingWorldEnv, BatchedCraftingWorldEnvFlat, BatchedCraftingWorldEnvOneHot,
                  BatchedCraftingWorldEnvAltObs, make_config)
from .host_env import HostCraftingWorldEnv  # noqa: E402
from .dist import StatsReducer, bind_to_gpu_numa_node, shard_range  # noqa: E402
from .vector import CraftingWorldVectorEnv, GifRecorder, register_envs  # noqa: E402

__all__ = ["BatchedCraftingWorldEnv", "BatchedCraftingWorldEnvFlat", "BatchedCraftingWorldEnvOneHot", "BatchedCraftingWorldEnvAltObs", "HostCraftingWorldEnv", "CraftingWorldVectorEnv", "GifRecorder", "register_envs", "StatsReducer", "shard_range", "bind_to_gpu_numa_node", "make_config", "TASK_LIST",
           "OBJECTS", "PICKUPABLE", "ACTION_NAMES", "COLORS_N", "MAX_STEPS"]
__version__ = "0.1.0"
