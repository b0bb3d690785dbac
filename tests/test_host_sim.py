"""The byte-SIMD device code (gym-2048_b200/csrc/g2048_device.cuh), compiled for the host
with prmt/umulhi/popc emulated, against the golden vectors and the oracle.  This is the
GPU-less check of the exact logic the CUDA kernels execute; tests/test_gpu_parity.py runs
the same checks on the device through the C ABI."""
import parity_checks as pc
from backends import SimOps


def test_philox():
    pc.check_philox(SimOps)


def test_shift_table_all_directions():
    pc.check_shift_table_all_directions(SimOps)


def test_csv_transitions():
    pc.check_csv_transitions(SimOps)


def test_special_boards():
    pc.check_special_boards(SimOps)


def test_golden_rollouts(golden_rollouts):
    pc.check_rollouts(SimOps, golden_rollouts)


def test_status_and_move_random_boards():
    pc.check_status_random_boards(SimOps)


def test_add_tile():
    pc.check_add_tile(SimOps)


def test_rollout_vs_oracle_random():
    pc.check_against_oracle(SimOps, n=4096, steps=48, seed=0, policy="random")


def test_rollout_vs_oracle_legal_midgame():
    pc.check_against_oracle(SimOps, n=2048, steps=400, seed=42, policy="legal", illegal_move_reward=-1.0,
                            env_id_base=(1 << 40) + 3)


def test_rollout_vs_oracle_max_tile_noreset():
    pc.check_against_oracle(SimOps, n=1024, steps=300, seed=456, policy="legal", max_tile_exp=6,
                            auto_reset=False)
