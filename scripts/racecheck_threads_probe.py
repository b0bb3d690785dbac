"""compute-sanitizer probe: a schedule of step launches over S env sets issued by one or two host threads.
   python scripts/racecheck_threads_probe.py <plain|chained> [n] [K] [threads] [episode 0|1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import gym_2048_b200 as g  # noqa: E402

mode = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 64
T = int(sys.argv[4]) if len(sys.argv) > 4 else 2
ep = int(sys.argv[5]) if len(sys.argv) > 5 else 0
S = 4
games = [g.BatchedGame2048(n, seed=1, env_id_base=s * n, outputs=("episode",) if (ep and s % 2) else ()) for s in range(S)]
for x in games:
    x.reset()
acts = torch.randint(0, 4, (8, n), device="cuda", dtype=torch.uint8)
sched = g.StepSchedule()
for j in range(K):
    sched.add(games[j % S], acts[j % 8], chained=("interleaved" if mode == "chained" else False))
if T > 1:
    sched.run_threads(T)
else:
    sched.run()
torch.cuda.synchronize()
print("ok", mode, n, K, T, ep, flush=True)
