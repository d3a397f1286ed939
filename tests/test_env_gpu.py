"""GPU parity of the CUDA environment round (mls_env_reset / mls_env_step / mls_env_info,
through the C ABI) against (a) the golden traces recorded from the unmodified reference
and (b) the numpy oracle on larger seeded batches.  Bit-exact: obs (fp32 bits), rewards
(fp64 bits), active / terminated / done, counters."""
import numpy as np
import pytest
import torch

from golden_util import GoldenEnv, check_reset, check_round, golden_env_files
from melissa_b200 import reset_chain
from melissa_b200.topology import GraphPool
from oracle.env_oracle import BatchedEnvOracle

pytestmark = pytest.mark.gpu

STAT_KEYS = ["total_messages_transmitted", "messages_sent", "messages_received", "n_neighbours",
             "interested_agents", "coverage_interested_count", "uninterested_with_message"]


def _env(*a, **k):
    from melissa_b200.batched_env import BatchedGraphEnv
    return BatchedGraphEnv(*a, **k)


def _tuples(gi, src, inter, scr, N):
    from melissa_b200.batched_env import ResetTuplesDevice
    return ResetTuplesDevice(gi, src, inter, scr, N, "cuda")


@pytest.mark.parametrize("path", golden_env_files(), ids=lambda p: p.split("env_")[-1][:-4])
def test_cuda_env_replays_reference_golden(path):
    g = GoldenEnv(path)
    kw = g.reset_args()
    pool = GraphPool(kw["adj"], kw["pos"])
    env = _env(g.E, g.N, pool, dynamic_graph=g.dynamic, is_testing=g.is_testing, heuristic=g.heuristic, want_info=True)
    tup = _tuples(np.arange(g.E), kw["source"], kw["interested"], kw["scripted"], g.N)
    obs, active = env.reset(tup, move_offsets=kw["move_offsets"])
    fl = env.flags()
    check_reset(g, obs.cpu().numpy(), active.cpu().numpy(), fl["has_message"], fl["msgs"])
    np.testing.assert_array_equal(env.recv_count.cpu().numpy(), g.z["reset_recv_count"])
    for r in range(g.max_rounds):
        alive, actions, mo = g.round_inputs(r)
        obs, rew, active, term, done = env.step(actions, move_offsets=mo)
        torch.cuda.synchronize()
        check_round(g, r, alive, obs.cpu().numpy(), rew.cpu().numpy(), active.cpu().numpy(), term.cpu().numpy(),
                    done.cpu().numpy(), env.rewards_sum.cpu().numpy(), env.episode.cpu().numpy()[:, 1])
        np.testing.assert_array_equal(env.recv_count.cpu().numpy()[alive], g.round_expected(r, "recv_count")[alive])
        inf = env.last_info()
        inf2 = env.info()
        st = g.round_expected(r, "stats")[alive]
        for k, key in enumerate(STAT_KEYS):
            np.testing.assert_array_equal(inf[key][alive], st[:, k], err_msg=key)
            np.testing.assert_array_equal(inf2[key][alive], st[:, k], err_msg=key + " (mls_env_info)")
        np.testing.assert_array_equal(inf["coverage"][alive], g.round_expected(r, "stats_coverage")[alive])
        np.testing.assert_array_equal(inf["coverage_interested_fraction"][alive],
                                      g.round_expected(r, "stats_cov_int_frac")[alive])
        np.testing.assert_array_equal(inf["n_acted"][alive], (actions >= 0).sum(1)[alive])


def _rollout_vs_oracle(N, B, n_graphs, rounds, *, dynamic=False, heuristic=None, ratio=0.0, is_testing=False, seed=0):
    side = None if N in (20, 50) else min(1.0, (N / 50.0) ** 0.5)
    pool = GraphPool.synthetic(N, n_graphs, first_seed=1000 + seed, side=side)
    gi, src, inter, scr, _ = reset_chain.episode_pool(9 + seed, B, N, n_graphs, scripted_agents_ratio=ratio)
    rng = np.random.default_rng(seed)
    mo0 = 0.06 * rng.uniform(-1, 1, size=(B, 2, N)) if dynamic else None
    env = _env(B, N, pool, dynamic_graph=dynamic, heuristic=heuristic, is_testing=is_testing, want_info=True)
    obs, active = env.reset(_tuples(gi, src, inter, scr, N), move_offsets=mo0)
    o = BatchedEnvOracle(B, N, dynamic=dynamic, heuristic=heuristic, is_testing=is_testing)
    o.reset(np.arange(B), pool.adj[gi], pool.pos[gi], src, inter, scr, move_offsets=mo0)
    np.testing.assert_array_equal(obs.cpu().numpy().view(np.uint32), o.obs().view(np.uint32))
    np.testing.assert_array_equal(active.cpu().numpy().astype(bool), o.active)
    total = 0
    for r in range(rounds):
        actions = rng.integers(0, 2, size=(B, N)).astype(np.int8)
        mo = 0.06 * rng.uniform(-1, 1, size=(B, 2, N)) if dynamic else None
        total += int(o.active.sum())
        acted = o.active.copy()
        eo, er, ea, et, ed = o.step(actions, move_offsets=mo)
        er = o.reward_vectorised(acted) if B > 256 else er
        obs, rew, active, term, done = env.step(actions, move_offsets=mo)
        ctx = f"round {r}"
        np.testing.assert_array_equal(obs.cpu().numpy().view(np.uint32), eo.view(np.uint32), err_msg=ctx)
        np.testing.assert_array_equal(rew.cpu().numpy().view(np.uint64), er.view(np.uint64), err_msg=ctx)
        np.testing.assert_array_equal(active.cpu().numpy().astype(bool), ea, err_msg=ctx)
        np.testing.assert_array_equal(term.cpu().numpy().astype(bool), et, err_msg=ctx)
        np.testing.assert_array_equal(done.cpu().numpy().astype(bool), ed, err_msg=ctx)
        np.testing.assert_array_equal(env.rewards_sum.cpu().numpy().view(np.uint64),
                                      o.episode_rewards_sum.view(np.uint64), err_msg=ctx)
        fl = env.flags()
        np.testing.assert_array_equal(fl["has_message"], o.has_message, err_msg=ctx)
        np.testing.assert_array_equal(fl["msgs"], o.msgs, err_msg=ctx)
        np.testing.assert_array_equal(fl["steps_taken"], o.steps_taken, err_msg=ctx)
        np.testing.assert_array_equal(env.recv_count.cpu().numpy(), o.received_from.sum(2), err_msg=ctx)
    assert int(env.transitions.item()) == total
    return total


@pytest.mark.parametrize("N,B", [(20, 512), (50, 256), (12, 64), (33, 64), (64, 32), (100, 16), (200, 8)])
def test_static_random_actions_vs_oracle(N, B):
    assert _rollout_vs_oracle(N, B, min(B, 32), 14) > 0


@pytest.mark.parametrize("N,B", [(20, 256), (50, 64), (200, 4)])
def test_dynamic_vs_oracle(N, B):
    _rollout_vs_oracle(N, B, min(B, 16), 10, dynamic=True, seed=3)


@pytest.mark.parametrize("heuristic", ["mpr", "simple_broadcast", "silent", "broadcast_if_any_interested"])
@pytest.mark.parametrize("dynamic", [False, True])
@pytest.mark.parametrize("is_testing", [False, True])
def test_scripted_vs_oracle(heuristic, dynamic, is_testing):
    _rollout_vs_oracle(20, 64, 8, 10, dynamic=dynamic, heuristic=heuristic, ratio=0.4, is_testing=is_testing, seed=5)
    if heuristic == "mpr" and not dynamic:
        _rollout_vs_oracle(50, 32, 8, 12, heuristic=heuristic, ratio=0.5, is_testing=is_testing, seed=6)


def test_config2_mpr_all_scripted_4096_episodes():
    """BASELINE config 2: MPR-heuristic env-only rollout, 20-node graphs, 4096 episodes.
    All agents scripted -> driven in testing mode so that every informed agent is stepped
    and the heuristic overrides it (graph.py:244,340)."""
    N, B = 20, 4096
    _rollout_vs_oracle(N, B, 64, 8, heuristic="mpr", ratio=1.0, is_testing=True, seed=7)


def test_probabilistic_heuristics_host_fed_bits():
    N, B = 20, 32
    for heuristic in ("probabilistic_gossip", "probabilistic_relay"):
        pool = GraphPool.synthetic(N, 8, first_seed=77)
        gi, src, inter, scr, _ = reset_chain.episode_pool(3, B, N, 8, scripted_agents_ratio=0.5)
        rng = np.random.default_rng(1)
        bits = lambda: dict(gossip_bits=rng.integers(0, 2, size=(B, N)).astype(np.uint8)) if heuristic.endswith("gossip") \
            else dict(relay_bits=rng.integers(0, 2, size=(B, N, N)).astype(np.uint8))
        env = _env(B, N, pool, heuristic=heuristic)
        o = BatchedEnvOracle(B, N, heuristic=heuristic)
        kw = bits()
        env.reset(_tuples(gi, src, inter, scr, N), **kw)
        o.reset(np.arange(B), pool.adj[gi], pool.pos[gi], src, inter, scr, **kw)
        for r in range(8):
            actions = rng.integers(0, 2, size=(B, N)).astype(np.int8)
            kw = bits()
            eo, er, ea, et, ed = o.step(actions, **kw)
            obs, rew, active, term, done = env.step(actions, **kw)
            np.testing.assert_array_equal(obs.cpu().numpy().view(np.uint32), eo.view(np.uint32))
            np.testing.assert_array_equal(rew.cpu().numpy().view(np.uint64), er.view(np.uint64))
            np.testing.assert_array_equal(active.cpu().numpy().astype(bool), ea)


def test_recycling_restarts_finished_episodes():
    """Episodes that end are restarted in the same launch from the recycle pool; the fresh
    episode must equal an explicit reset from the same tuple."""
    N, B, P = 20, 128, 512
    pool = GraphPool.synthetic(N, 16, first_seed=5)
    gi, src, inter, scr, _ = reset_chain.episode_pool(100, P, N, 16)
    env = _env(B, N, pool)
    tup_all = _tuples(gi, src, inter, scr, N)
    env.set_recycling(tup_all)
    env.reset(_tuples(gi[:B], src[:B], inter[:B], scr[:B], N))
    o = BatchedEnvOracle(B, N)
    o.reset(np.arange(B), pool.adj[gi[:B]], pool.pos[gi[:B]], src[:B], inter[:B], scr[:B])
    n_resets = np.ones(B, dtype=np.int64)
    rng = np.random.default_rng(0)
    restarted = 0
    for r in range(40):
        actions = rng.integers(0, 2, size=(B, N)).astype(np.int8)
        eo, er, ea, et, ed = o.step(actions)
        obs, rew, active, term, done = env.step(actions)
        np.testing.assert_array_equal(rew.cpu().numpy().view(np.uint64), er.view(np.uint64))
        np.testing.assert_array_equal(done.cpu().numpy().astype(bool), ed)
        ids = np.flatnonzero(ed)
        if len(ids):
            t = (ids + n_resets[ids] * B) % P
            o.reset(ids, pool.adj[gi[t]], pool.pos[gi[t]], src[t], inter[t], scr[t])
            n_resets[ids] += 1
            restarted += len(ids)
        np.testing.assert_array_equal(obs.cpu().numpy().view(np.uint32), o.obs().view(np.uint32), err_msg=f"round {r}")
        np.testing.assert_array_equal(active.cpu().numpy().astype(bool), o.active)
    assert restarted > B


def test_invalid_arguments_raise_value_error():
    pool = GraphPool.synthetic(20, 2)
    with pytest.raises(ValueError):
        _env(4, 20, pool, heuristic="nope")
    env = _env(4, 20, pool)
    with pytest.raises(ValueError):
        env.step(np.zeros((4, 19), dtype=np.int8))


def test_dynamic_philox_movement_is_consistent_and_deterministic():
    """dynamic_graph=True without host-fed offsets: the device draws 0.06*U(-1,1) from Philox4x32-10.
    Every round each coordinate moves by at most 0.06, the adjacency is exactly the d^2 <= r^2 graph of
    the moved positions (fp64), the same seed reproduces the trajectory, another seed does not."""
    N, B = 50, 64
    pool = GraphPool.synthetic(N, 8, first_seed=11)
    gi, src, inter, scr, _ = reset_chain.episode_pool(21, B, N, 8)
    from melissa_b200.topology import unpack_adjacency

    def run(seed):
        env = _env(B, N, pool, dynamic_graph=True)
        env.philox_seed = seed
        env.reset(_tuples(gi, src, inter, scr, N))
        traj = []
        prev = env.pos.cpu().numpy().copy()
        rng = np.random.default_rng(0)
        for r in range(6):
            env.step(rng.integers(0, 2, size=(B, N)).astype(np.int8))
            pos = env.pos.cpu().numpy()
            adj = unpack_adjacency(env.adj.cpu().numpy().view(np.uint32), N)
            assert np.abs(pos - prev).max() <= 0.06 + 1e-12 and np.abs(pos - prev).max() > 0.03
            d = pos[:, :, None, :] - pos[:, None, :, :]
            want = ((d * d).sum(-1) <= 0.2 * 0.2) & ~np.eye(N, dtype=bool)
            np.testing.assert_array_equal(adj, want)
            obs = env.obs.cpu().numpy()
            np.testing.assert_array_equal(obs[:, :, 0], pos[:, :, 0].astype(np.float32))
            np.testing.assert_array_equal(obs[:, :, 2], want.sum(2).astype(np.float32))
            traj.append(pos.copy())
            prev = pos.copy()
        return np.stack(traj)

    a, b, c = run(5), run(5), run(6)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    moves = np.diff(np.concatenate([a[:1] * 0 + pool.pos[gi][None], a]), axis=0)[1:]
    assert abs(moves.mean()) < 2e-3 and 0.03 < moves.std() < 0.04      # U(-0.06, 0.06): std = 0.0346


@pytest.mark.parametrize("N,B", [(200, 8), (72, 16), (150, 8)])
def test_phantom_nodes_from_stale_shared_memory(N, B):
    """N = 65..96 and 129..224: the episode's bitmask rows have more words (W = 4 / 8) than the episode has warps.
    The unused words must read as zero whatever an earlier kernel left in shared memory (they used to be
    uninitialised: phantom acting nodes, wrong reward sums, episodes that never finish).  A tcgen05 GEMM with random
    operands dirties ~200 KB of every SM's shared memory first."""
    import ctypes as C
    from melissa_b200 import _lib
    L = _lib.lib()
    L.mls_test_gemm_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    M, K, Nn = 148 * 128 * 2, 512, 256
    A = (torch.randn(M, K, device="cuda") * 1e4).to(torch.bfloat16)
    Bm = (torch.randn(Nn, K, device="cuda") * 1e4).to(torch.bfloat16)
    out = torch.empty(M, Nn, dtype=torch.bfloat16, device="cuda")
    for _ in range(2):
        _lib.check(L.mls_test_gemm_bf16(A.data_ptr(), Bm.data_ptr(), None, None, 0, 1, out.data_ptr(), M, Nn, K, 0, None,
                                        _lib.current_stream_ptr()))
    torch.cuda.synchronize()
    assert _rollout_vs_oracle(N, B, min(B, 8), 10) > 0
