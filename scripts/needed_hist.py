"""Distribution of controlling nodes / needed rows / conv2 edges per graph-round at the bench state (GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from melissa_b200.batched_env import BatchedGraphEnv, ResetTuplesDevice
from melissa_b200.networks import NETWORKS
from melissa_b200.rollout import Rollout

N, B, G = 50, 8192, 1024
dev = torch.device("cuda")
pool = bench.load_pool(N, G)
gi, src, inter, scr = bench.load_tuples(N, G, 65536)
env = BatchedGraphEnv(B, N, pool, device=dev)
torch.manual_seed(9)
net = NETWORKS["l_dgn"](5, 128, 2, 4, N, dueling_param=({"hidden_sizes": [128, 128]}, {"hidden_sizes": [128, 128]}), device="cuda").to(dev)
net.set_precision("bf16")
ro = Rollout(env, net, eps=0.05, seed=9)
ro.start(ResetTuplesDevice(gi, src, inter, scr, N, dev))
adj = torch.as_tensor(pool.adj, device=dev)
for r in range(40):
    ro.round()
    if r >= 36:
        act = env.active.bool()
        gid = env.episode[:, 3].long()
        a = adj[gid]                                            # [B, i, j]
        need = act | (a & act[:, :, None]).any(1)
        cnt, nc = act.sum(1), need.sum(1)
        E = (a & act[:, :, None]).sum((1, 2)) + cnt
        q = lambda t: [float(torch.quantile(t.float(), p)) for p in (0.5, 0.75, 0.9, 0.99, 1.0)]
        small = ((cnt <= 32) & (nc <= 32) & (E <= 256)).float().mean().item()
        mid = ((cnt <= 32) & (nc <= 48) & (E <= 320)).float().mean().item()
        print(f"round {r}: cnt mean {cnt.float().mean():.1f} q {q(cnt)}  nc mean {nc.float().mean():.1f} q {q(nc)}  E mean {E.float().mean():.1f} q {q(E)} "
              f"small {small:.3f} mid {mid:.3f} empty {(cnt == 0).float().mean():.3f}")
