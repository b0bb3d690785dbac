import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_rollouts():
    z = load_golden("rollouts.npz")
    out = {}
    for name in z["names"]:
        name = str(name)
        out[name] = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}
    return out


def rollout_cfg(rec):
    seed, n_envs, n_steps, base, legal, max_exp, auto_reset = (int(x) for x in rec["meta"])
    return dict(seed=seed, n=n_envs, T=n_steps, env_id_base=base, max_tile_exp=max_exp,
                auto_reset=bool(auto_reset), illegal_move_reward=float(rec["illegal_reward"]))


STEP_KEYS = ("boards", "rewards", "dones", "illegal", "highest_exp", "legal_mask", "ep_score", "ep_len")
DONE_KEYS = ("terminal_boards", "final_score", "final_len")
