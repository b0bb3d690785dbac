"""CPU-only checks of the boundary: the C-ABI library builds, loads and exports every symbol
include/g2048.h declares (no compute calls without a GPU), struct layouts match the header,
host helpers behave, and the product never imports the oracle."""
import ctypes as C
import os
import re
import subprocess

import pytest

import gym_2048_b200 as g
from conftest import ROOT
from oracle import oracle


def header_functions():
    src = open(os.path.join(ROOT, "include", "g2048.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(g2048_\w+)\s*\(", src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    g.build()
    L = g._lib.lib()
    names = header_functions()
    assert len(names) >= 17
    for name in names:
        assert hasattr(L, name), name
    assert sorted(g._lib.EXPORTS) == names
    assert L.g2048_abi_version() == g._lib.ABI_VERSION == 5


def test_struct_layouts_match_header():
    """sizeof/offsetof of the ctypes mirrors == what gcc computes from include/g2048.h."""
    prog = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "g2048.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu\n", sizeof(G2048StepArgs), offsetof(G2048StepArgs, n),
             offsetof(G2048StepArgs, illegal_move_reward), offsetof(G2048StepArgs, flags),
             sizeof(G2048EnvConfig), sizeof(G2048HostStepOut));
      printf("%zu %zu %zu %zu\n", sizeof(G2048StepManyArgs), offsetof(G2048StepManyArgs, n),
             offsetof(G2048StepManyArgs, n_steps), offsetof(G2048StepManyArgs, flags));
      printf("%zu %zu %zu %zu %zu\n", offsetof(G2048StepArgs, boards_out), offsetof(G2048StepArgs, ep_return),
             offsetof(G2048StepArgs, final_return), sizeof(G2048OneIO), offsetof(G2048OneIO, reward));
      printf("%zu %zu\n", offsetof(G2048OneIO, done), offsetof(G2048OneIO, bad_cells));
      printf("%zu %zu %zu %zu %u\n", offsetof(G2048StepArgs, boards_nibble), offsetof(G2048StepArgs, nibble_overflow),
             offsetof(G2048EnvConfig, board_format), offsetof(G2048StepArgs, chain), G2048_CHAIN_WORDS);
      return 0;
    }'''
    d = os.path.join(ROOT, "tests", "host_sim")
    src, exe = os.path.join(d, "_layout.c"), os.path.join(d, "_layout.bin")
    open(src, "w").write(prog)
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src])
    out = [int(x) for x in subprocess.check_output([exe]).split()]
    os.remove(src), os.remove(exe)
    for S in (g._lib.StepArgs, oracle.StepArgs):
        assert [C.sizeof(S), S.n.offset, S.illegal_move_reward.offset, S.flags.offset] == out[:4]
    assert C.sizeof(g._lib.EnvConfig) == out[4] and C.sizeof(g._lib.HostStepOut) == out[5]
    M = g._lib.StepManyArgs
    assert [C.sizeof(M), M.n.offset, M.n_steps.offset, M.flags.offset] == out[6:10]
    O = g._lib.OneIO
    for S in (g._lib.StepArgs, oracle.StepArgs):
        assert [S.boards_out.offset, S.ep_return.offset, S.final_return.offset] == out[10:13]
    assert [C.sizeof(O), O.reward.offset, O.done.offset, O.bad_cells.offset] == [out[13], out[14], out[15], out[16]]
    for S in (g._lib.StepArgs, oracle.StepArgs):
        assert [S.boards_nibble.offset, S.nibble_overflow.offset] == out[17:19]
    assert g._lib.EnvConfig.board_format.offset == out[19]
    for S in (g._lib.StepArgs, oracle.StepArgs):
        assert S.chain.offset == out[20]
    assert g._lib.CHAIN_WORDS == out[21]


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(g.G2048Error):
        g.BatchedGame2048(8)
    with pytest.raises(g.G2048Error):
        g.Game2048Env().reset(seed=0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gym-2048_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no CPU", ""), os.path.join(dirpath, f)


def test_tile_to_exp_and_shard_range():
    assert g.tile_to_exp(None) == 0 and g.tile_to_exp(2048) == 11 and g.tile_to_exp(2) == 1
    assert g.tile_to_exp(3000) == 63 and g.tile_to_exp(0) == 63 and g.tile_to_exp(1) == 63
    with pytest.raises(AssertionError):
        g.tile_to_exp(2048.0)                      # reference :72 asserts int
    for total, world in ((1 << 20, 8), (10, 3), (7, 8)):
        spans = [g.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (b0, c0), (b1, _) in zip(spans, spans[1:]):
            assert b0 + c0 == b1


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py --impl reference (the reference's own CPU step(), or the oracle port when the reference package
    is absent) needs no GPU; the driver parses its stdout, which must be ONE JSON line with the contract's keys."""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0


def test_argument_checks_that_need_no_gpu():
    """Arguments are validated before anything touches the device: bad flags, missing pointers and alignment are
    reported through the return code and g2048_last_error() on a box without a GPU too."""
    L = g._lib.lib()
    fake = 0x10000                      # never dereferenced: every call below fails validation first
    a = g._lib.StepArgs()
    a.boards, a.actions, a.rewards, a.dones, a.n = fake, fake, fake, fake, 8
    a.flags = g._lib.FLAG_AUTO_RESET | 64
    assert L.g2048_step(C.byref(a), None) == -1 and b"unknown flags" in L.g2048_last_error()
    a.flags = g._lib.FLAG_AUTO_RESET | g._lib.FLAG_POLICY_LEGAL            # the random-legal policy reads legal_mask
    assert L.g2048_step(C.byref(a), None) == -1 and b"legal_mask" in L.g2048_last_error()
    a.flags = g._lib.FLAG_POLICY_UNIFORM | g._lib.FLAG_POLICY_LEGAL
    assert L.g2048_step(C.byref(a), None) == -1 and b"choose one" in L.g2048_last_error()
    a.flags = g._lib.FLAG_AUTO_RESET | g._lib.FLAG_CHAINED                 # a chained launch needs the chain buffer
    assert L.g2048_step(C.byref(a), None) == -1 and b"chain buffer" in L.g2048_last_error()
    a.flags, a.chain = g._lib.FLAG_AUTO_RESET, fake + 4
    assert L.g2048_step(C.byref(a), None) == -2 and b"chain must be 8-byte aligned" in L.g2048_last_error()
    a.chain, a.step_counter = fake, fake
    assert L.g2048_step(C.byref(a), None) == -1 and b"cannot be combined" in L.g2048_last_error()
    a.chain, a.step_counter = None, None
    a.flags = g._lib.FLAG_AUTO_RESET
    assert L.g2048_step_n(C.byref(a), 4, 3, None) == -1 and b"row_stride" in L.g2048_last_error()
    a.step_counter = fake
    assert L.g2048_step_n(C.byref(a), 4, 8, None) == -1 and b"step_counter" in L.g2048_last_error()
    a.step_counter = None
    assert L.g2048_step_list(None, 3, None) == -1 and L.g2048_step_list(None, 0, None) == 0
    assert L.g2048_one(None, 0, 0, 0, 0, 0, 0.0, 0, None, 1) == -1 and b"io is NULL" in L.g2048_last_error()
    assert L.g2048_one(fake, 9, 0, 0, 0, 0, 0.0, 0, None, 1) == -1 and b"unknown op" in L.g2048_last_error()
    a.flags, a.max_tile_exp = g._lib.FLAG_AUTO_RESET, 64
    assert L.g2048_step(C.byref(a), None) == -1 and b"max_tile_exp" in L.g2048_last_error()
    a.max_tile_exp, a.boards = 0, fake + 4
    assert L.g2048_step(C.byref(a), None) == -2 and b"16-byte aligned" in L.g2048_last_error()
    a.boards, a.rewards = fake, None
    assert L.g2048_step(C.byref(a), None) == -1 and b"required" in L.g2048_last_error()
    m = g._lib.StepManyArgs()
    m.boards, m.rewards, m.dones, m.n, m.n_steps = fake, fake, fake, 8, 4
    assert L.g2048_step_many(C.byref(m), None) == -1 and b"actions are required" in L.g2048_last_error()
    m.flags = g._lib.FLAG_POLICY_UNIFORM | g._lib.FLAG_POLICY_LEGAL
    assert L.g2048_step_many(C.byref(m), None) == -1 and b"choose one" in L.g2048_last_error()
    m.flags, m.actions, m.actions_out = 0, fake, fake
    assert L.g2048_step_many(C.byref(m), None) == -1 and b"actions_out" in L.g2048_last_error()
    m.actions_out, m.flags = None, 64
    assert L.g2048_step_many(C.byref(m), None) == -1 and b"unknown flags" in L.g2048_last_error()
    m.n_steps = 0                       # nothing to do is not an error
    assert L.g2048_step_many(C.byref(m), None) == 0


def test_bench_rank_cpus_are_disjoint_whole_cores():
    """bench.py pins every rank of an N-GPU run to its own physical cores: the sets are disjoint and lie in the allowed set."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    allowed = sorted(os.sched_getaffinity(0))
    for world in (1, 2, 4, 8):
        sets = [bench.rank_cpus(allowed, r, world) for r in range(world)]
        flat = [c for s_ in sets for c in s_]
        assert len(flat) == len(set(flat)) and set(flat) <= set(allowed)
        assert world > len(allowed) or all(sets)


def test_bench_issue_plan_is_the_single_stream_schedule_per_env_set():
    """bench.py with T issuing threads: every env set is stepped exactly as often, with exactly the action rows, and in the
    order the single-stream schedule (launch j -> set j % S, row j % P) steps it — the state checksum cannot depend on T."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for total, S, P, T, R, K in ((16384 + 5 + 500, 32, 16, 2, 25, 20), (2048 + 5 + 500, 32, 16, 4, 25, 20), (64 + 3 + 12, 4, 16, 2, 3, 4),
                                 (4096 + 25 * 50, 16, 16, 2, 25, 50), (300, 6, 5, 3, 2, 6)):
        single = {}
        for j in range(total):
            single.setdefault(j % S, []).append(j % P)
        plan = bench.issue_plan(total, S, P, T, total - R * K, R, K // T)
        got = {}
        for t, (items, start) in enumerate(plan):
            assert all(s_ % T == t for s_, _ in items)
            assert 0 <= start and start + R * (K // T) <= len(items)
            for s_, row in items:
                got.setdefault(s_, []).append(row)
        assert got == single
        assert sum(len(items) for items, _ in plan) == total


def test_unpack_boards_host_is_the_nibble_layout():
    """g2048_unpack_boards_host (what the library's threads run for G2048_BOARDS_BYTES_PACKED_WIRE): cell c of a board is
    nibble c of its 8 packed bytes — odd counts (scalar tail), unaligned destinations, every nibble value."""
    import numpy as np
    L = g._lib.lib()
    rng = np.random.default_rng(3)
    for n, off in ((0, 0), (1, 0), (2, 0), (3, 1), (255, 0), (1000, 5), (65537, 16)):
        boards = rng.integers(0, 16, (n, 16)).astype(np.uint8)
        packed = (boards[:, 0::2] | (boards[:, 1::2] << 4)).astype(np.uint8)          # byte b = cells 2b (low), 2b+1 (high)
        packed = np.ascontiguousarray(packed)
        raw = np.zeros(n * 16 + off + 16, np.uint8)
        out = raw[off:off + n * 16]
        assert L.g2048_unpack_boards_host(packed.ctypes.data if n else None, out.ctypes.data if n else None, n) == 0
        assert np.array_equal(out.reshape(n, 16), boards) and not raw[off + n * 16:].any() and not raw[:off].any()
    assert L.g2048_unpack_boards_host(None, None, 4) == -1


def test_chain_state_machine_of_the_batched_class():
    """BatchedGame2048._chain_args without a GPU: which launches carry the chain buffer, G2048_FLAG_CHAINED and
    G2048_FLAG_CHAIN_INTERLEAVED — a chained step only chains to a step of the same chain, same mode, with nothing
    else in between; plain steps are launched without the buffer."""
    import torch
    L = g._lib
    game = g.BatchedGame2048.__new__(g.BatchedGame2048)
    game.device, game._chain, game._chain_broken, game._chain_mode, game._step_counter = torch.device("cpu"), None, True, "direct", None

    def flags(chained):
        a = L.StepArgs()
        a.flags = L.FLAG_AUTO_RESET
        game._chain_args(a, chained)
        return bool(a.chain), bool(a.flags & L.FLAG_CHAINED), bool(a.flags & L.FLAG_CHAIN_INTERLEAVED)
    assert flags(False) == (False, False, False) and game._chain is None          # plain: no buffer is ever made
    assert flags(True) == (True, False, False)                                    # first chained step: the head
    assert game._chain.numel() == L.CHAIN_WORDS and int(game._chain.abs().sum()) == 0
    assert flags(True) == (True, True, False) and flags(True) == (True, True, False)
    assert flags("interleaved") == (True, False, True)                            # another launch shape: a head again
    assert flags("interleaved") == (True, True, True)
    assert flags(False) == (False, False, False)                                  # a plain step in between ...
    assert flags("interleaved") == (True, False, True)                            # ... breaks the chain
    assert flags("interleaved") == (True, True, True)
    game._chain_broken = True                                                     # what reset / set_boards / step_many do
    assert flags("interleaved") == (True, False, True)
    game._step_counter = torch.zeros(2, dtype=torch.int64)                        # device-side step index: no chaining
    assert flags(False) == (False, False, False)
    with pytest.raises(g.G2048Error):
        flags(True)
    with pytest.raises(ValueError):
        flags("sideways")
