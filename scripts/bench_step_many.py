import sys, torch
sys.path.insert(0, '.')
import gym_2048_b200 as g
dev = torch.device('cuda', 0)
for n in (65536, 131072, 262144, 1 << 20):
    K = 64
    game = g.BatchedGame2048(n, seed=1, device=dev, outputs=())
    game.reset()
    acts = torch.randint(0, 4, (K, n), dtype=torch.uint8, device=dev)
    rew = torch.empty((K, n), dtype=torch.float32, device=dev); dn = torch.empty((K, n), dtype=torch.uint8, device=dev)
    for _ in range(3): game.step_many(acts, rewards=rew, dones=dn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L = 50
    e0.record()
    for _ in range(L): game.step_many(acts, rewards=rew, dones=dn)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (L * K)
    print("step_many n %8d  %.3f us/step  %.3e steps/s" % (n, us, n / us * 1e6))
