#!/usr/bin/env python
"""Every kernel of libg2048.so other than the step, timed alone on one B200 through the C ABI and set
against the measured HBM roofline (MEASURED_PEAKS.json): algorithmic bytes per item / time per launch.

    python scripts/bench_kernels.py [n_boards]        # under gpurun; prints a markdown table

Inputs are real boards (reset + 40 random-legal steps); buffers rotate over a working set larger than
the 126 MB L2; CUDA-event timing after warm-up.  Measured for the record (profiles/), the headline
number is bench.py's."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gym_2048_b200 as g  # noqa: E402
from gym_2048_b200._lib import check, lib  # noqa: E402

DEV = torch.device("cuda", 0)


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def timed(fn, iters=200, warm=20):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / iters


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    torch.cuda.set_device(0)
    L = lib()
    s = torch.cuda.current_stream().cuda_stream
    P = lambda t: C.c_void_p(t.data_ptr())
    game = g.BatchedGame2048(n, seed=3, device=DEV, outputs=("legal_mask",))
    game.reset()
    for _ in range(40):
        game.step(game.sample_actions(legal=True))
    R = 8                                   # rotating copies: the working set exceeds the L2
    boards = [game.boards.clone() for _ in range(R)]
    nxt = [game.boards.clone() for _ in range(R)]
    masks = [game.legal_mask.clone() for _ in range(R)]
    acts = [torch.randint(0, 4, (n,), dtype=torch.uint8, device=DEV) for _ in range(R)]
    rew = [torch.rand(n, device=DEV) for _ in range(R)]
    dones = [(torch.rand(n, device=DEV) < 0.02).to(torch.uint8) for _ in range(R)]
    u8 = [torch.empty(n, dtype=torch.uint8, device=DEV) for _ in range(4)]
    u32 = torch.empty(n, dtype=torch.int32, device=DEV)
    rows = []

    def add(name, bytes_per_item, seconds, items=n, note=""):
        gbs = bytes_per_item * items / seconds / 1e9
        rows.append((name, bytes_per_item, seconds * 1e6, gbs, note))

    add("g2048_reset", 16, timed(lambda i: check(L.g2048_reset(P(boards[i % R]), None, n, 0, 1, i, s))))
    add("g2048_add_tile", 32, timed(lambda i: check(L.g2048_add_tile(P(nxt[i % R]), n, 0, 1, i, s))))
    add("g2048_move (boards, scores, changed out)", 38,
        timed(lambda i: check(L.g2048_move(P(boards[i % R]), P(nxt[i % R]), P(acts[i % R]), P(u32), P(u8[0]), n, s))))
    add("g2048_status (mask, highest, empties, isend)", 20,
        timed(lambda i: check(L.g2048_status(P(boards[i % R]), P(u8[0]), P(u8[1]), P(u8[2]), P(u8[3]), 0, n, s))))
    add("g2048_sample_actions (uniform)", 1, timed(lambda i: check(L.g2048_sample_actions(None, P(u8[0]), n, 0, 1, i, s))),
        note="compute-bound: one Philox2x32 block per byte written")
    add("g2048_sample_actions (legal)", 2,
        timed(lambda i: check(L.g2048_sample_actions(P(masks[i % R]), P(u8[0]), n, 0, 1, i, s))), note="same")
    for name, code, dt, per in (("u8", g._lib.OBS_U8, torch.uint8, 1), ("bf16", g._lib.OBS_BF16, torch.bfloat16, 2),
                                ("f32", g._lib.OBS_F32, torch.float32, 4), ("i64", g._lib.OBS_I64, torch.int64, 8)):
        m = n if per <= 2 else n // 4
        obs = [torch.empty((m, 16, 4, 4), dtype=dt, device=DEV) for _ in range(2 if per >= 4 else 3)]
        add("g2048_encode_obs %s" % name, 16 + 256 * per,
            timed(lambda i: check(L.g2048_encode_obs(P(boards[i % R]), P(obs[i % len(obs)]), code, m, s)), iters=100), items=m)
        del obs
    vals = torch.empty((n, 16), dtype=torch.int64, device=DEV)
    add("g2048_values_from_exp", 16 * 9, timed(lambda i: check(L.g2048_values_from_exp(P(boards[i % R]), P(vals), n * 16, s)), iters=100))
    add("g2048_exp_from_values", 16 * 9, timed(lambda i: check(L.g2048_exp_from_values(P(vals), P(nxt[i % R]), n * 16, None, s)), iters=100))
    del vals
    add("g2048_symmetry (boards + next + actions, hflip, k=1)", 66,
        timed(lambda i: check(L.g2048_symmetry(P(boards[i % R]), P(boards[(i + 1) % R]), P(nxt[i % R]), P(nxt[(i + 1) % R]),
                                               P(acts[i % R]), P(acts[(i + 1) % R]), n, 1, 1, s))))
    m = n // 4
    ab = torch.empty((8 * m, 16), dtype=torch.uint8, device=DEV)
    an = torch.empty((8 * m, 16), dtype=torch.uint8, device=DEV)
    aa = torch.empty(8 * m, dtype=torch.uint8, device=DEV)
    ar = torch.empty(8 * m, dtype=torch.float32, device=DEV)
    ad = torch.empty(8 * m, dtype=torch.uint8, device=DEV)
    add("g2048_augment (1 row in, 8 out)", 38 * 9,
        timed(lambda i: check(L.g2048_augment(P(boards[i % R]), P(nxt[i % R]), P(acts[i % R]), P(rew[i % R]), P(dones[i % R]), m,
                                              P(ab), P(an), P(aa), P(ar), P(ad), s)), iters=100), items=m)
    del ab, an, aa, ar, ad
    ret = torch.empty(n, dtype=torch.float64, device=DEV)
    add("g2048_discounted_return (f64 out, 2 % done rows)", 13,
        timed(lambda i: check(L.g2048_discounted_return(P(rew[i % R]), P(dones[i % R]), P(ret), n, 0.99, s)), iters=50),
        note="one thread per episode segment walks it backwards: latency-bound")
    T, ne = 256, 65536
    rw = torch.rand((T, ne), device=DEV); va = torch.rand((T, ne), device=DEV)
    es = (torch.rand((T, ne), device=DEV) < 0.02).to(torch.uint8)
    lv = torch.rand(ne, device=DEV); ld = torch.zeros(ne, dtype=torch.uint8, device=DEV)
    adv = torch.empty((T, ne), device=DEV); rt = torch.empty((T, ne), device=DEV)
    add("g2048_gae (T=256 x 65,536 envs)", 17,
        timed(lambda i: check(L.g2048_gae(P(rw), P(va), P(es), P(lv), P(ld), P(adv), P(rt), T, ne, 0.99, 0.95, s)), iters=30),
        items=T * ne, note="65,536 threads walk 256 dependent steps")
    # references on the same box: a pure write stream and a copy (what the roofline denominator measures)
    big = torch.empty(1 << 30, dtype=torch.uint8, device=DEV)
    big2 = torch.empty(1 << 30, dtype=torch.uint8, device=DEV)
    add("(reference) cudaMemset, 1 GiB", 1, timed(lambda i: big.zero_(), iters=20, warm=3), items=1 << 30, note="write-only stream")
    add("(reference) device copy, 1 GiB", 2, timed(lambda i: big2.copy_(big), iters=20, warm=3), items=1 << 30, note="read + write")
    del big, big2
    pk, src = peak()
    print("| kernel | algorithmic B/item | µs per launch | GB/s | of %s HBM roofline (%.0f GB/s) | note |" % (src, pk))
    print("|---|---|---|---|---|---|")
    for name, b, us, gbs, note in rows:
        print("| `%s` | %d | %.1f | %.0f | %.2f | %s |" % (name, b, us, gbs, gbs / pk, note))


if __name__ == "__main__":
    main()
