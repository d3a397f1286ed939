"""TEST INFRASTRUCTURE / REPORTED CPU BASELINE ONLY.

CPU rollout of the same hot path with the oracle port: numpy environment round
(oracle/env_oracle.py) + pure-torch network forward per agent observation
(oracle/net_oracle.py, the reference's cost structure: one GNN pass per agent-obs row,
reference shared_policy.py:135-154) + greedy / epsilon-greedy action selection.

Used by ``bench.py`` for the ``cpu_baseline`` object and for ``--impl reference``.  The
unmodified reference itself cannot run on the GPU box (/root/reference is not there and
pettingzoo / tianshou / torch_geometric are not installable), hence ``kind: "port"``.
"""
from __future__ import annotations

import os
import time

import numpy as np


def _setup(args):
    import torch
    torch.set_num_threads(1)
    from melissa_b200 import reset_chain
    from melissa_b200.topology import GraphPool
    from oracle import net_oracle
    from oracle.env_oracle import BatchedEnvOracle
    kind, N, B, n_graphs, seed, eps = args["kind"], args["N"], args["B"], args["n_graphs"], args["seed"], args["eps"]
    pool = GraphPool.synthetic(N, n_graphs, first_seed=0)
    gi, src, inter, scr, _ = reset_chain.episode_pool(seed, B, N, n_graphs)
    env = BatchedEnvOracle(B, N)
    env.reset(np.arange(B), pool.adj[gi], pool.pos[gi], src, inter, scr)
    sd = net_oracle.init_state_dict(kind, seed=9) if kind else None
    return dict(env=env, sd=sd, pool=pool, rng=np.random.default_rng(seed), kind=kind, N=N, B=B, eps=eps,
                tuples=(gi, src, inter, scr), cursor=0)


def _round(st):
    import torch
    from oracle import net_oracle
    env, N, B = st["env"], st["N"], st["B"]
    n_acted = int(env.active.sum())
    if st["kind"]:
        kw = dict(aggregator="max") if st["kind"] == "hl_dgn" else {}
        with torch.no_grad():
            q = net_oracle.forward_graphs(st["kind"], st["sd"], torch.as_tensor(env.obs()),
                                          torch.as_tensor(env.active), N, **kw).numpy()
        act = (q[..., 1] > q[..., 0]).astype(np.int8)
        if st["eps"] > 0:
            u = st["rng"].random((B, N, 3))
            rnd = (u[..., 2] + 1.0 > u[..., 1] + 1.0).astype(np.int8)
            act = np.where(u[..., 0] < st["eps"], rnd, act)
    else:
        act = st["rng"].integers(0, 2, size=(B, N)).astype(np.int8)
    acted = env.active.copy()
    ids = np.arange(B)
    env.steps_taken += acted
    env._world_step(ids, np.where(acted, act, -1).astype(np.int8))
    rew = env.reward_vectorised(acted)
    env.episode_rewards_sum += rew.sum(1)
    env.num_moves += 1
    env.truncated |= acted & (env.steps_taken >= 4)
    env.active = env.has_message & ~env.truncated & ~env.scripted
    done = np.flatnonzero(~env.active.any(axis=1))
    if len(done):                                   # restart finished episodes (vector-env behaviour)
        gi, src, inter, scr = st["tuples"]
        t = (done + st["cursor"]) % B
        st["cursor"] += 1
        env.reset(done, st["pool"].adj[gi[t]], st["pool"].pos[gi[t]], src[t], inter[t], scr[t])
    return n_acted


def worker(args):
    """Run ``warmup`` + ``rounds`` rounds on this process's shard; -> (transitions, seconds)."""
    st = _setup(args)
    for _ in range(args["warmup"]):
        _round(st)
    t0 = time.perf_counter()
    n = 0
    for _ in range(args["rounds"]):
        n += _round(st)
    return n, time.perf_counter() - t0


def run_parallel(kind, N, episodes_per_proc, rounds, warmup, n_procs=None, eps=0.05, n_graphs=32):
    """All host cores, one shard per process (the reference's own parallelism is one env
    per subprocess, l_dgn.py:137).  -> dict(value transitions/s, cores, transitions, seconds)."""
    import multiprocessing as mp
    n_procs = n_procs or os.cpu_count() or 1
    jobs = [dict(kind=kind, N=N, B=episodes_per_proc, n_graphs=n_graphs, seed=9 + 1000 * p, eps=eps, rounds=rounds,
                 warmup=warmup) for p in range(n_procs)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(n_procs) as pool:
        res = pool.map(worker, jobs)
    total = sum(r[0] for r in res)
    secs = max(r[1] for r in res)
    return dict(value=total / secs if secs > 0 else 0.0, cores=n_procs, transitions=total, seconds=secs)
