"""gym-2048_b200 — B200-native batched 2048 environment (drop-in for rgal/gym-2048's env).

Public surface:
  BatchedGame2048   N boards on one GPU, one fused CUDA kernel per step (batched.py); step_n / capture /
                    StepSchedule issue many steps per interpreter trip (small batches are launch-bound)
  HostSteppedEnv    host-buffer handle of the C ABI (actions/results in pinned host memory)
  Game2048Env       single-env class with the reference's gymnasium + game API (env.py)
  Game2048VecEnv    Stable-Baselines3-style VecEnv adapter over BatchedGame2048 (vec_env.py)
  Transitions / TransitionRecorder  the reference's transition table + CSV schema on the GPU (transitions.py)
  RolloutCollector  on-device PPO rollout loop + compact rollout buffer + GAE kernel (rollout.py)
  evaluate_model    the reference's train.evaluate_model with all episodes at once (evaluate.py)
  stack, IllegalMove  as in the reference module
The CUDA extension is built in-tree by `_lib.build()` (nvcc, sm_100a); nothing here has a
CPU fallback.
"""
from . import _lib
from ._lib import G2048Error, build
from .batched import ALL_OUTPUTS, BatchedGame2048, HostSteppedEnv, StepResult, StepSchedule, shard_range, tile_to_exp
from .stats import EpisodeStats
from .env import Game2048Env, IllegalMove, register, stack
from .vec_env import Game2048VecEnv
from .transitions import TransitionRecorder, Transitions
from .rollout import RolloutCollector, gae_reference
from .evaluate import evaluate_model, report_evaluation_results
from .policy import ResNetActorCritic

__all__ = ["BatchedGame2048", "HostSteppedEnv", "StepResult", "StepSchedule", "Game2048Env", "Game2048VecEnv", "IllegalMove",
           "stack", "register", "shard_range", "tile_to_exp", "EpisodeStats", "build", "G2048Error", "ALL_OUTPUTS",
           "Transitions", "TransitionRecorder", "RolloutCollector", "gae_reference", "evaluate_model",
           "report_evaluation_results", "ResNetActorCritic"]
__version__ = "0.2.0"

import os as _os
if _os.environ.get("G2048_NO_REGISTER", "") != "1":      # like the reference's env/__init__.py:1-6: id '2048-v0' on import
    register()
