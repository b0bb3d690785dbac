#!/usr/bin/env python
"""Measure the BASELINE.json configs that are not the bench.py line (2, 4, 5) on one B200.

    python scripts/bench_configs.py [c2] [c4] [c5] [--steps K]

Prints one JSON line per config (device-resident, CUDA-event timed, boards produced by the env
itself).  These are parity-test configurations measured for the record (profiles/), not the
headline number — bench.py reports config 3."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import gym_2048_b200 as g  # noqa: E402

DEV = torch.device("cuda", 0)


def timed(fn, iters, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / iters


def c2(args):
    """65,536 envs, random actions, lean outputs: per-step launches, and 64 steps per CUDA graph."""
    n = 65536
    sets = 64                                                     # 64 x 2.4 MB > L2? no: 156 MB > 126 MB L2
    games = [g.BatchedGame2048(n, seed=s, device=DEV, env_id_base=s * n, outputs=()) for s in range(sets)]
    gen = torch.Generator(device=DEV).manual_seed(1)
    pool = torch.randint(0, 4, (16, n), generator=gen, device=DEV, dtype=torch.uint8)
    for gm in games:
        gm.reset()
        for t in range(6):
            gm.step(pool[t])
    state = {"i": 0}

    def one():
        i = state["i"]
        games[i % sets].step(pool[i % 16])
        state["i"] = i + 1
    dt = timed(one, args.steps)
    out = {"config": "c2", "workload": "65,536 envs, uniform-random actions, lean outputs, 64 env sets round-robin",
           "us_per_step": dt * 1e6, "env_steps_per_s": n / dt, "launch": "one kernel launch per step (PDL)"}
    # same work as one CUDA graph of 64 steps (one per set), device-side step counters
    for gm in games:
        gm.use_device_step_counter(True)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for gm in games:
            gm.step(pool[0])
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for s, gm in enumerate(games):
                gm.step(pool[s % 16])
    dtg = timed(graph.replay, max(args.steps // sets, 20), warm=5) / sets
    out.update({"graph_us_per_step": dtg * 1e6, "graph_env_steps_per_s": n / dtg,
                "graph": "%d steps (+ counter bumps) per CUDA graph replay" % sets})
    return out


def c4(args):
    """262,144 envs, legal-mask output + auto-reset, actions uniform among legal moves (device policy)."""
    n = 262144
    sets = 8
    games = [g.BatchedGame2048(n, seed=42, device=DEV, env_id_base=s * n, outputs=("legal_mask",)) for s in range(sets)]
    acts = torch.empty(n, dtype=torch.uint8, device=DEV)
    for gm in games:
        gm.reset()
        for _ in range(300):                                      # reach the mid-game regime
            gm.step(gm.sample_actions(legal=True, out=acts))
    state = {"i": 0}

    def one():
        gm = games[state["i"] % sets]
        gm.step(gm.sample_actions(legal=True, out=acts))
        state["i"] += 1

    def policy_only():
        gm = games[state["i"] % sets]
        gm.sample_actions(legal=True, out=acts)
        state["i"] += 1
    dt = timed(one, args.steps)
    empties = float((games[0].boards == 0).float().sum(1).mean())
    dtp = timed(policy_only, args.steps)
    # the same play as one launch per 64 steps: the policy runs inside g2048_step_many (mask and action of
    # every step still written out)
    K = 64
    lean = [g.BatchedGame2048(n, seed=42, device=DEV, env_id_base=s * n, outputs=("legal_mask",)) for s in range(2)]
    for gm, src in zip(lean, games):
        gm.set_boards(src.boards)
    rew = torch.empty((K, n), dtype=torch.float32, device=DEV)
    dn = torch.empty((K, n), dtype=torch.uint8, device=DEV)
    ac = torch.empty((K, n), dtype=torch.uint8, device=DEV)
    lm = torch.empty((K, n), dtype=torch.uint8, device=DEV)

    def fused():
        gm = lean[state["i"] % 2]
        gm.step_many(policy="legal", n_steps=K, rewards=rew, dones=dn, actions_out=ac, legal_mask_out=lm)
        state["i"] += 1
    dtf = timed(fused, 40, warm=3) / K
    return {"config": "c4", "fused_us_per_step": dtf * 1e6, "fused_env_steps_per_s": n / dtf,
            "fused": "g2048_step_many with G2048_FLAG_POLICY_LEGAL, 64 steps per launch, actions + legal masks written every step", "workload": "262,144 envs, legal mask + auto-reset, random-legal policy on device, "
                                          "8 env sets round-robin, mid-game boards",
            "us_per_step_policy_plus_step": dt * 1e6, "env_steps_per_s": n / dt,
            "us_policy_kernel_only": dtp * 1e6, "us_step_kernel_by_difference": (dt - dtp) * 1e6,
            "mean_empty_cells": empties}


def c5(args):
    """PPO rollout: 65,536 envs x 256-step horizon, ppo_train.py's ResNet trunk (random init)."""
    n, T = 65536, 256 if not args.quick else 16
    torch.manual_seed(42)
    game = g.BatchedGame2048(n, seed=42, device=DEV, outputs=())
    game.reset()
    policy = g.ResNetActorCritic().to(DEV).eval().to(memory_format=torch.channels_last)
    out = {"config": "c5", "workload": "PPO rollout 65,536 envs x %d steps, ResNet 64f x 4 blocks (%.2f MFLOP/obs), "
                                       "random init" % (T, policy.flops_per_obs() / 1e6)}
    for name, dtype in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        pol = policy.to(dtype)                                    # in place: fp32 first, then bf16
        col = g.RolloutCollector(game, pol, T, obs_dtype=dtype, channels_last=True, seed=1)
        small = g.RolloutCollector(game, pol, 2, obs_dtype=dtype, channels_last=True, seed=1)
        small.collect()                                           # warm-up (cuDNN autotune)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        col.collect()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        # the env's share: the same T observe+step calls without the policy
        acts = col.actions

        def env_only():
            for t in range(T):
                game.observe(dtype, out=col._obs)
                game.step(acts[t], boards_out=col.boards[t + 1])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        game.boards = col.boards[0]
        e0.record()
        env_only()
        e1.record()
        torch.cuda.synchronize()
        env_s = e0.elapsed_time(e1) * 1e-3
        out[name] = {"rollout_s": wall, "env_steps_per_s": n * T / wall, "env_share_of_rollout": env_s / wall,
                     "env_only_s": env_s, "policy_TFLOPs": n * (T + 1) * policy.flops_per_obs() / wall / 1e12,
                     "rollout_buffer_MB": sum(x.numel() * x.element_size() for x in
                                              (col.boards, col.actions, col.rewards, col.values, col.log_probs,
                                               col.advantages, col.returns, col.episode_starts)) / 1e6}
        del col, small
    return out


def ev(args):
    """train.py's evaluate_model, batched: 16,384 episodes to the end (or the 2001-move cap), greedy on a fixed
    preference order with illegal moves masked (a corner strategy; ~180 moves per episode), and on the random-init
    ResNet policy in bf16."""
    n = 16384 if not args.quick else 1024
    pref = torch.tensor([[4.0, 1.0, 2.0, 3.0]], device=DEV)
    out = {"config": "eval", "workload": "%d episodes, evaluate_model semantics (illegal reward -1, 2001-move cap)" % n}
    net = g.ResNetActorCritic().to(DEV).eval().to(torch.bfloat16).to(memory_format=torch.channels_last)

    def resnet(obs):
        return net(obs.to(memory_format=torch.channels_last))[0]
    for name, model, dt in (("fixed_preference", lambda obs: pref.expand(obs.shape[0], 4), torch.uint8),
                            ("resnet_bf16", resnet, torch.bfloat16)):
        g.evaluate_model(model, 256, mask_illegal=True, obs_dtype=dt, device=DEV)          # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = g.evaluate_model(model, n, mask_illegal=True, obs_dtype=dt, device=DEV)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        moves = sum(e["moves"] for e in res["Episodes"])
        out[name] = {"wall_s": wall, "episodes_per_s": n / wall, "env_steps_per_s": moves / wall,
                     "mean_moves": moves / n, "average_score": res["Average score"], "highest_tile": res["Highest tile"]}
    return out


def rec(args):
    """gather_training_data.py's recording loop, batched: 4,096 envs x 256 steps of random-legal play recorded out of
    place, then transitions() (game order, illegal moves dropped), augment() (8 symmetries) and the discounted
    return, all on the device."""
    n, T = 4096, 256
    game = g.BatchedGame2048(n, seed=0, device=DEV, illegal_move_reward=-1.0, outputs=("illegal", "legal_mask"))
    game.reset()
    recorder = g.TransitionRecorder(game, horizon=T)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(T):
        recorder.step(game.sample_actions(legal=True))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    data = recorder.transitions()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    rows = data.size()
    data.augment()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    data.get_discounted_return(0.99)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    return {"config": "recorder", "workload": "4,096 envs x 256 random-legal steps recorded, then transitions / augment / returns",
            "record_s": t1 - t0, "record_env_steps_per_s": n * T / (t1 - t0), "rows": rows,
            "transitions_s": t2 - t1, "augment_s": t3 - t2, "augmented_rows": data.size(), "returns_s": t4 - t3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["c2", "c4", "c5"])
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    for c in args.configs:
        res = {"c2": c2, "c4": c4, "c5": c5, "eval": ev, "rec": rec}[c](args)
        res["gpu"] = torch.cuda.get_device_name(0)
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
