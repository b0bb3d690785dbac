"""Drop-in conformance of the single-env surface.

* registration: `import gym_2048_b200` registers '2048-v0' like the reference's `env/__init__.py:1-6`, and
  `gymnasium.make('2048-v0')` hands back the GPU class (through oracle/shim's stand-in when the real
  gymnasium is absent — the shim is test infrastructure, put on the path by this file only);
* the compat `env` package (`gym-2048_b200/compat`) makes the reference's own import lines resolve to the
  GPU class;
* `-m gpu`: the reference's test file `env/envs/test_game2048_env.py`, UNMODIFIED (the copy pip installed
  into baseline/_ref), is run by pytest against gym_2048_b200.Game2048Env.

Each check runs in a fresh interpreter: module aliasing (`env`, `gymnasium`) must not leak into the other
tests of this process, which import the real reference under the same names."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "gym-2048_b200", "compat")
SHIM = os.path.join(ROOT, "oracle", "shim")
REF_TEST = os.path.join(ROOT, "baseline", "_ref", "env", "envs", "test_game2048_env.py")


def _have_gymnasium():
    try:
        import gymnasium  # noqa: F401
        return not getattr(gymnasium, "__version__", "").endswith("shim")
    except ImportError:
        return False


def _env(*front):
    e = dict(os.environ)
    paths = list(front) + ([] if _have_gymnasium() else [SHIM]) + [ROOT]
    e["PYTHONPATH"] = os.pathsep.join(paths + [e.get("PYTHONPATH", "")])
    e.pop("G2048_NO_REGISTER", None)
    return e


def _run(code, *front):
    return subprocess.run([sys.executable, "-c", code], env=_env(*front), capture_output=True, text=True, timeout=600)


def test_import_registers_2048_v0_and_make_returns_the_gpu_class():
    r = _run("""
import gymnasium
import gym_2048_b200
from gym_2048_b200.env import Game2048Env
e = gymnasium.make('2048-v0')
u = e.unwrapped
assert type(u) is Game2048Env, type(u)
assert u.action_space.n == 4 and tuple(u.observation_space.shape) == (16, 4, 4)
assert u.metadata['render_modes'] == ['ansi', 'human', 'rgb_array']
e2 = gymnasium.make('2048-v0', render_mode='ansi')
assert e2.unwrapped.render_mode == 'ansi'
print('ok')
""")
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr


def test_registration_can_be_disabled():
    code = """
import os
os.environ['G2048_NO_REGISTER'] = '1'
import gymnasium
from gymnasium.envs.registration import registry
import gym_2048_b200
assert '2048-v0' not in registry
assert gym_2048_b200.register() and '2048-v0' in registry
print('ok')
"""
    r = _run(code)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr


def test_compat_env_package_serves_the_reference_import_lines():
    r = _run("""
import env                                    # ppo_train.py:18, registers '2048-v0' -> env.envs:Game2048Env
import env.envs.game2048_env as game2048_env  # test_game2048_env.py:6
from env.envs import Game2048Env              # env/envs/__init__.py:1
import gymnasium
import gym_2048_b200.env as ours
assert game2048_env.Game2048Env is ours.Game2048Env is Game2048Env
assert game2048_env.IllegalMove is ours.IllegalMove and game2048_env.stack is ours.stack
assert type(gymnasium.make('2048-v0').unwrapped) is ours.Game2048Env
print('ok')
""", COMPAT)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr


_RUNNER = """
import sys, pytest
import env.envs.game2048_env as m                      # the compat package: cached in sys.modules before pytest imports the file
import gym_2048_b200.env as ours
assert m.Game2048Env is ours.Game2048Env, m.__file__


class Check:                                           # the file's own `game2048_env` must be that module
    def pytest_collection_modifyitems(self, items):
        assert len(items) == 10, [i.nodeid for i in items]
        for it in items:
            assert it.module.game2048_env.Game2048Env is ours.Game2048Env, it.module.game2048_env.__file__


sys.exit(pytest.main([%r, '-q', '--import-mode=importlib', '-p', 'no:cacheprovider', '-c', %r, '--rootdir', %r],
                     plugins=[Check()]))
"""


REF_TEST_SHA256 = "03776d72f254e764c4045fb647d96e064f756b9d25d57d9ca4ea3d4113f53247"   # of /root/reference/env/envs/test_game2048_env.py


def _run_reference_test_file():
    import hashlib
    with open(REF_TEST, "rb") as f:                        # unmodified: the bytes of the reference's own file
        assert hashlib.sha256(f.read()).hexdigest() == REF_TEST_SHA256
    return _run(_RUNNER % (REF_TEST, os.devnull, os.path.dirname(REF_TEST)), COMPAT)


@pytest.mark.skipif(not os.path.isfile(REF_TEST), reason="baseline/_ref not installed (run __graft_entry__.build())")
def test_reference_test_file_resolves_to_the_gpu_class_without_a_gpu():
    """Plumbing check that runs anywhere: the unmodified file is collected (10 tests) and its Game2048Env is
    ours — without a CUDA device every test that touches the device fails LOUDLY with G2048Error (there is no
    CPU fallback), with one it passes (the -m gpu test below)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("covered by the -m gpu run of the file")
    r = _run_reference_test_file()
    assert "10 failed" in r.stdout or "failed" in r.stdout, r.stdout + r.stderr
    assert "G2048Error" in r.stdout and "no CPU fallback" in r.stdout, r.stdout[-3000:]


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isfile(REF_TEST), reason="baseline/_ref not installed (run __graft_entry__.build())")
def test_reference_test_file_unmodified_passes_against_the_gpu_class():
    """/root/reference/env/envs/test_game2048_env.py:10-231, byte for byte as pip installed it, run by pytest in a
    fresh interpreter whose `env.envs.game2048_env` is gym_2048_b200 (compat package first on the path)."""
    r = _run_reference_test_file()
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "10 passed" in r.stdout, r.stdout[-2000:]
