"""Episode statistics over a sharded batch.

The step path has no collective; the only cross-GPU exchange this package offers is an
optional all-reduce (NCCL on GPUs, gloo in the CPU tests) of a ~40-element vector of episode
statistics — what `HighestTileCallback` (`/root/reference/ppo_train.py:69-82`) and SB3's
`Monitor` aggregate per process.  Pure torch plumbing, off the timed path.
"""
import torch
import torch.distributed as dist

N_EXP = 32          # histogram bins for the highest tile exponent


class EpisodeStats:
    """Running sums over finished episodes: count, return, length, illegal endings, and a
    histogram of the highest tile (as exponent).  `update()` takes the tensors a step returned."""

    def __init__(self, device="cpu"):
        self.vec = torch.zeros(4 + N_EXP, dtype=torch.float64, device=device)
        self.steps = 0

    def update(self, dones, final_score, final_len, highest_exp, illegal):
        d = dones.to(torch.bool)
        w = d.to(torch.float64)
        self.vec[0] += w.sum()
        self.vec[1] += (final_score.to(torch.float64) * w).sum()
        self.vec[2] += (final_len.to(torch.float64) * w).sum()
        self.vec[3] += (illegal.to(torch.float64) * w).sum()
        self.vec[4:] += torch.bincount(highest_exp[d].to(torch.int64), minlength=N_EXP)[:N_EXP].to(torch.float64)
        self.steps += int(dones.numel())

    def all_reduce(self, group=None):
        """Sum over all ranks (no-op without an initialised process group)."""
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(self.vec, op=dist.ReduceOp.SUM, group=group)
        return self

    def summary(self):
        v = self.vec.detach().cpu()
        n = max(float(v[0]), 1.0)
        hist = v[4:]
        return {"episodes": int(v[0]), "mean_return": float(v[1]) / n, "mean_length": float(v[2]) / n,
                "illegal_endings": int(v[3]),
                "highest_tile_hist": {int(1 << e): int(c) for e, c in enumerate(hist.tolist()) if c > 0},
                "mean_highest_tile": float(sum((1 << e) * c for e, c in enumerate(hist.tolist()))) / n}
