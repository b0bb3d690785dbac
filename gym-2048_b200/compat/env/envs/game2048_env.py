"""`env.envs.game2048_env` of the reference, served by the GPU implementation (module-level names the
reference module exports: Game2048Env, IllegalMove, stack — game2048_env.py:14-34)."""
from gym_2048_b200.env import Game2048Env, IllegalMove, stack  # noqa: F401
