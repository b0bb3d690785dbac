"""Batched model evaluation — `evaluate_model` of `/root/reference/train.py:122-229` with all
episodes running at once on the GPU.

Reference semantics kept: illegal-move reward -1 (:183-184), an illegal move ends the episode
(the env terminates on it), epsilon-greedy over the model's action probabilities (:100-119),
the 2000-move cap checked after the increment so an episode plays at most 2001 moves
(:157-160), per-episode `total_reward / highest / moves / illegal_moves`, the summary keys and
the `scores_<label>.csv` report (:217-229).  Episode i is env id i of the draw stream (the
reference seeds numpy's PCG64 with 456+i, which no test pins — SURVEY §8c)."""
import csv

import torch

from .batched import BatchedGame2048


@torch.no_grad()
def evaluate_model(model, episodes, epsilon=0.0, seed=456, agent_seed=123, device=None, max_moves=2000,
                   mask_illegal=False, obs_dtype=torch.float32, verbose=False, env_id_base=0):
    """model: callable obs [B,16,4,4] -> action probabilities (or scores) [B,4].
    mask_illegal=True restricts the greedy choice to legal moves (not in the reference)."""
    game = BatchedGame2048(episodes, seed=seed, device=device, env_id_base=env_id_base, illegal_move_reward=-1.0,
                           auto_reset=False, outputs=("illegal", "highest", "legal_mask"))
    dev = game.device
    game.reset()
    gen = torch.Generator(device=dev).manual_seed(int(agent_seed))
    n = episodes
    active = torch.ones(n, dtype=torch.bool, device=dev)
    total_reward = torch.zeros(n, dtype=torch.float64, device=dev)
    moves = torch.zeros(n, dtype=torch.int64, device=dev)
    illegals = torch.zeros(n, dtype=torch.int64, device=dev)
    highest = torch.zeros(n, dtype=torch.int64, device=dev)
    bits = torch.arange(4, device=dev, dtype=torch.uint8)[None, :]
    steps = 0
    while steps <= max_moves:                                   # the cap fires after move 2001 (:157-160)
        obs = game.observe(obs_dtype)
        scores = model(obs).float()
        if mask_illegal:
            legal = ((game.legal_mask[:, None] >> bits) & 1).bool()
            legal = legal | ~legal.any(dim=1, keepdim=True)
            scores = scores.masked_fill(~legal, float("-inf"))
        action = torch.argmax(scores, dim=1)
        if epsilon > 0:
            explore = torch.rand(n, generator=gen, device=dev) <= epsilon     # `uniform(0,1) > epsilon` is greedy
            rnd = torch.randint(0, 4, (n,), generator=gen, device=dev)
            action = torch.where(explore, rnd, action)
        r = game.step(action.to(torch.uint8))
        total_reward += torch.where(active, r.rewards.double(), 0.0)
        illegals += (active & r.illegal).long()
        moves += active.long()
        highest = torch.where(active, r.highest_exp.long(), highest)
        active = active & ~r.dones
        steps += 1
        if steps % 16 == 0 and not bool(active.any()):
            break
    hi = torch.where(highest > 0, torch.ones_like(highest) << highest, highest)
    tr, mv, il, hv = total_reward.tolist(), moves.tolist(), illegals.tolist(), hi.tolist()
    scores_list = [{"total_reward": tr[i], "highest": int(hv[i]), "moves": int(mv[i]), "illegal_moves": int(il[i])}
                   for i in range(n)]
    if verbose:
        for i, s in enumerate(scores_list):
            print(f"Episode {i}, epsilon {epsilon}, highest {s['highest']}, reward {s['total_reward']:.1f}, "
                  f"moves {s['moves']}, illegals {s['illegal_moves']}")
    return {
        "Average score": sum(tr) / n,
        "Max score": max(tr),
        "Highest tile": int(max(hv)),
        "Episodes": scores_list,
    }


def report_evaluation_results(results, label="eval"):
    """Write `scores_<label>.csv` exactly as the reference does (:217-229)."""
    with open(f"scores_{label}.csv", "w") as f:
        fieldnames = ["total_reward", "highest", "moves", "illegal_moves"]
        writer = csv.DictWriter(f, fieldnames=fieldnames, lineterminator="\n")
        writer.writeheader()
        for s in results["Episodes"]:
            writer.writerow(s)
