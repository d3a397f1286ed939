"""Where does one DQN update go?  (GPU, torch profiler summary of DQNPolicy.update at the bench's training configuration)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from melissa_b200.batched_env import BatchedGraphEnv, ResetTuplesDevice
from melissa_b200.data_parallel import FlatParameters, FusedAdam
from melissa_b200.networks import NETWORKS
from melissa_b200.policy import BatchedCollector, DQNPolicy, Batch
from melissa_b200.replay import DeviceReplay
from melissa_b200.networks.autograd import q_values

N, B, G, M = 50, 8192, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda")
pool = bench.load_pool(N, G)
gi, src, inter, scr = bench.load_tuples(N, G, 65536)
torch.manual_seed(9)
net = NETWORKS["l_dgn"](5, 128, 2, 4, N, dueling_param=({"hidden_sizes": [128, 128]}, {"hidden_sizes": [128, 128]}), device="cuda").to(dev)
net.set_precision("bf16")
optim = FusedAdam(FlatParameters(net), lr=1e-3)
pol = DQNPolicy(net, optim, 0.99, 4, 500, eps=0.05)
env = BatchedGraphEnv(B, N, pool, device=dev, want_info=True)
replay = DeviceReplay(B, N, 8, device=dev)
col = BatchedCollector(agents_num=N, policy=pol, env=env, buffer=replay, exploration_noise=True, tuples=ResetTuplesDevice(gi, src, inter, scr, N, dev, pool_size=G))
for _ in range(8):
    col.iterate(0.05)
for _ in range(3):
    pol.update(M, replay)
torch.cuda.synchronize()
ev = lambda: torch.cuda.Event(enable_timing=True)
t = [ev() for _ in range(6)]
t[0].record()
rho, ep, ag = replay.sample_indices(M, 4)
t[1].record()
b = Batch(replay.gather(rho, ep, ag, 4, 0.99))
t[2].record()
q = q_values(net, b.obs)
t[3].record()
loss = (b.returns - q.gather(1, b.act.view(-1, 1)).squeeze(1)).pow(2).mean()
optim.zero_grad()
loss.backward()
t[4].record()
optim.step()
t[5].record()
torch.cuda.synchronize()
names = ["sample", "gather", "forward", "backward", "adam"]
print({n: round(t[i].elapsed_time(t[i + 1]), 3) for i, n in enumerate(names)})
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    pol.update(M, replay)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
