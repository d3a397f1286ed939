"""GPU: policy / collector surfaces (melissa_b200/policy.py) -- tianshou DQNPolicy semantics and the
MultiAgentCollector statistics, checked against the oracle driven in lockstep."""
import numpy as np
import pytest
import torch

from melissa_b200 import reset_chain
from melissa_b200.topology import GraphPool
from oracle import net_oracle as no
from oracle.env_oracle import BatchedEnvOracle

pytestmark = pytest.mark.gpu
DUELING = lambda: ({"hidden_sizes": [128, 128]}, {"hidden_sizes": [128, 128]})


def _setup(N=20, B=48, P=256, seed=3):
    from melissa_b200.batched_env import BatchedGraphEnv, ResetTuplesDevice
    from melissa_b200.networks import LDGNNetwork
    from melissa_b200.policy import BatchedCollector, DQNPolicy, MultiAgentSharedPolicy
    pool = GraphPool.synthetic(N, 8, first_seed=seed)
    tup = reset_chain.episode_pool(seed, P, N, 8)
    sd = no.init_state_dict("l_dgn", seed=seed)
    net = LDGNNetwork(5, 128, 2, 4, N, dueling_param=DUELING(), device="cuda")
    net.load_state_dict(sd)
    net = net.cuda()
    env = BatchedGraphEnv(B, N, pool, want_info=True)
    pol = MultiAgentSharedPolicy(DQNPolicy(net, eps=0.0, seed=9), [str(i) for i in range(N)])
    col = BatchedCollector(pol, env, ResetTuplesDevice(*tup[:4], N, "cuda"))
    return pool, tup, sd, net, env, pol, col


def test_dqn_policy_forward_and_checkpoint_prefix():
    from melissa_b200.policy import DQNPolicy
    pool, tup, sd, net, env, pol, col = _setup()
    N = 20
    rows = np.concatenate([env.obs.cpu().numpy().reshape(env.B, -1), np.arange(env.B)[:, None] % N], axis=1).astype(np.float32)
    out = pol(rows, mask=np.ones((env.B, 2)))
    want_q = no.l_dgn_forward(sd, torch.as_tensor(rows), N).numpy()
    assert np.abs(out["logits"].cpu().numpy() - want_q).max() <= 1e-5 * max(1.0, np.abs(want_q).max())
    sure = np.abs(want_q[:, 1] - want_q[:, 0]) > 1e-4
    np.testing.assert_array_equal(out["act"][sure], no.dqn_act(want_q, np.ones((env.B, 2)))[sure])
    ck = pol.policy.state_dict()
    assert all(k.startswith("model.") for k in ck) and "model.conv1.lin_l.weight" in ck
    p2 = DQNPolicy(type(net)(5, 128, 2, 4, N, dueling_param=DUELING()).cuda())
    p2.load_state_dict(ck)
    assert torch.equal(p2.model.conv2.att, net.conv2.att)
    # host-side exploration noise follows tianshou: rand < eps -> argmax(rand(2) + mask)
    pol.policy.set_eps(0.5)
    act = np.zeros(4000, dtype=np.int64)
    noisy = pol.exploration_noise(act, np.ones((4000, 2)), rng=np.random.default_rng(0))
    assert 0.2 < noisy.mean() < 0.3
    pol.policy.set_eps(0.0)
    assert pol.exploration_noise(act) is act


def test_collector_statistics_match_oracle_in_lockstep():
    """Greedy policy; the oracle environment is driven with the actions the device chose.  Transitions,
    episode returns (fp64 bits), lengths and logger stats must agree."""
    N, B, P = 20, 48, 256
    pool, tup, sd, net, env, pol, col = _setup(N, B, P)
    gi, src, inter, scr = tup[:4]
    o = BatchedEnvOracle(B, N)
    o.reset(np.arange(B), pool.adj[gi[:B]], pool.pos[gi[:B]], src[:B], inter[:B], scr[:B])
    n_resets = np.ones(B, dtype=np.int64)
    want_steps, want_returns, want_lens, want_cov = 0, [], [], []
    got_returns, got_lens, got_cov = [], [], []
    for r in range(30):
        want_steps += int(o.active.sum())
        col.iterate()
        acts = col.act.cpu().numpy()
        o.step(acts)
        done = ~o.active.any(axis=1)
        np.testing.assert_array_equal(env.done.cpu().numpy().astype(bool), done)
        ids = np.flatnonzero(done)
        if len(ids):
            inf_o = o.info()
            want_returns += inf_o["episode_rewards_sum"][ids].tolist()
            want_lens += inf_o["num_moves"][ids].tolist()
            want_cov += (inf_o["covered"][ids] / N).tolist()
            inf = env.last_info()
            got_returns += inf["episode_rewards_sum"][ids].tolist()
            got_lens += inf["num_moves"][ids].tolist()
            got_cov += inf["coverage"][ids].tolist()
            t = (ids + n_resets[ids] * B) % P
            o.reset(ids, pool.adj[gi[t]], pool.pos[gi[t]], src[t], inter[t], scr[t])
            n_resets[ids] += 1
        np.testing.assert_array_equal(env.obs.cpu().numpy().view(np.uint32), o.obs().view(np.uint32))
    assert int(env.transitions.item()) == want_steps
    assert len(want_returns) > B
    np.testing.assert_array_equal(np.array(got_returns).view(np.uint64), np.array(want_returns).view(np.uint64))
    assert got_lens == want_lens and got_cov == want_cov


def test_collect_n_episode_and_n_step():
    pool, tup, sd, net, env, pol, col = _setup()
    st = col.collect(n_episode=60)
    assert st.n_collected_episodes >= 60 and len(st.returns) == st.n_collected_episodes == len(st.lens)
    assert st.n_collected_steps == int(env.transitions.item()) and st.collect_speed > 0
    assert st.returns_stat.max >= st.returns_stat.mean >= st.returns_stat.min
    assert 0.0 < st.info["coverage"].mean <= 1.0 and st.info["messages_sent"].min >= 1
    before = col.collect_step
    st2 = col.collect(n_step=500)
    assert st2.n_collected_steps >= 500 and col.collect_step == before + st2.n_collected_steps
    with pytest.raises(TypeError):
        col.collect()
    st3 = col.collect(n_step=200, random=True)
    assert st3.n_collected_steps >= 200


def test_evaluation_collect_in_testing_mode():
    """C5 "eval": is_testing=True environment, reset tuples from the reference's test stream
    (RandomState(17) seeds, density ladder), greedy policy (eps_test ~ 0), n_episode collection."""
    from melissa_b200.batched_env import BatchedGraphEnv, ResetTuplesDevice
    from melissa_b200.networks import HLDGNNetwork
    from melissa_b200.policy import BatchedCollector, DQNPolicy
    N, B = 20, 16
    pool = GraphPool.synthetic(N, 6, first_seed=900)
    gi, src, inter, scr, mv, dens = reset_chain.testing_episode_pool(64, N, 6, num_test_episodes=10, seed=1)
    net = HLDGNNetwork(5, 128, 2, 4, N, aggregator="max", dueling_param=DUELING(), device="cuda").cuda()
    env = BatchedGraphEnv(B, N, pool, is_testing=True, want_info=True)
    col = BatchedCollector(DQNPolicy(net, eps=0.001), env, ResetTuplesDevice(gi, src, inter, scr, N, "cuda"))
    st = col.collect(n_episode=40)
    assert st.n_collected_episodes >= 40 and (st.lens >= 1).all()
    assert 0.0 < st.info["coverage"].mean <= 1.0
    assert set(np.unique(st.info and np.round(dens, 1))) <= {0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0}


@pytest.mark.parametrize("precision,dynamic", [("bf16", False), ("fp32", True)])
def test_pipelined_host_round_equals_single_shot_round(precision, dynamic):
    """Rollout.round_host(sub_batches=4): H2D / compute / D2H of episode slices overlap on three streams.
    The host buffers after every round equal those of the single-shot host round, WITH exploration (eps = 0.3): the
    sliced environment step keys recycling and the device movement stream by the batch-wide episode index
    (include/melissa_b200.h MlsEnvDesc.episode_offset) and the exploration draws are keyed by the batch-wide row
    (MlsForwardArgs.philox_row0).  Host observations travel packed (12 bytes per node); unpacked they are bit-equal
    to the environment's obs rows."""
    from melissa_b200.batched_env import BatchedGraphEnv, ResetTuplesDevice
    from melissa_b200.networks import LDGNNetwork
    from melissa_b200.rollout import Rollout
    N, B, P, seed = 20, 50, 160, 5
    pool = GraphPool.synthetic(N, 8, first_seed=seed)
    tup = reset_chain.episode_pool(seed, P, N, 8)
    sd = no.init_state_dict("l_dgn", seed=seed)
    outs = []
    for sub in (1, 4, -4, -3):                   # negative: slices replayed from CUDA graphs, rounds issued without waiting
        net = LDGNNetwork(5, 128, 2, 4, N, dueling_param=DUELING(), device="cuda")
        net.load_state_dict(sd)
        net = net.cuda().set_precision(precision)
        env = BatchedGraphEnv(B, N, pool, dynamic_graph=dynamic)
        ro = Rollout(env, net, eps=0.3, seed=9)
        nowait = sub < 0
        if sub < 0:
            sub = -sub
            ro.start(ResetTuplesDevice(*tup[:4], N, "cuda"))
            ro.capture_host(sub)
        ro.start(ResetTuplesDevice(*tup[:4], N, "cuda"))
        ro.sync_host()                           # pinned mirrors <- reset state
        ro.round_host(sub)
        trace = []
        for _ in range(14):                      # several episode lifetimes: restarts from the recycle pool included
            if nowait:
                ro.round_host(sub, wait=False)
                ro.host_drain()
            else:
                ro.round_host(sub)
            torch.cuda.synchronize()
            trace.append({k: v.clone() for k, v in ro._host.items()})
            assert np.array_equal(Rollout.unpack_obs_host(ro._host["obs"]).view(np.uint32), env.obs.cpu().numpy().view(np.uint32))
        outs.append((trace, ro.transitions()))
        assert ro.feature_violations() == 0
    (t1, n1) = outs[0]
    assert n1 > 0
    for t4, n4 in outs[1:]:
        assert n1 == n4
        for r, (a, b) in enumerate(zip(t1, t4)):
            for k in a:
                assert torch.equal(a[k], b[k]), (r, k)
