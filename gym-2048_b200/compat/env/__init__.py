"""Drop-in `env` package: the reference's `env/__init__.py:1-6` with the GPU class behind it.

Put `gym-2048_b200/compat` on PYTHONPATH ahead of the reference checkout and the reference's callers
(`ppo_train.py:18` `import env`, `train.py`, `gather_training_data.py:232` `gym.make('2048-v0')`,
`env/envs/test_game2048_env.py:6` `import env.envs.game2048_env`) resolve to gym_2048_b200 unchanged."""
from gymnasium.envs.registration import register

register(
    id='2048-v0',
    entry_point='env.envs:Game2048Env',
)
