from env.envs.game2048_env import Game2048Env  # noqa: F401  (reference env/envs/__init__.py:1)
