"""Pins oracle/env_oracle.py against the UNMODIFIED reference (builder container only).

The reference is driven through its AEC API (oracle/ref_driver.py); the oracle gets the
same reset tuple, the same per-round actions (and movement offsets / scripted random
bits) and must reproduce obs, rewards, active sets and world counters bit-exactly.
"""
import os
import tempfile

import networkx as nx
import numpy as np
import pytest

from melissa_b200 import reset_chain, topology
from oracle.env_oracle import BatchedEnvOracle, mpr_select, two_hop
from oracle.ref_loader import in_dir, load_reference, reference_available

pytestmark = pytest.mark.reference


def _mk_env(ref, graph, N, **kw):
    return ref.GraphEnv(graph=graph, number_of_agents=N, radius=0.2, **kw)


def _compare_state(o: BatchedEnvOracle, rec, b=0, ctx=""):
    np.testing.assert_array_equal(o.has_message[b], rec["has_message"], err_msg=ctx)
    np.testing.assert_array_equal(o.msgs[b], rec["msgs"], err_msg=ctx)
    np.testing.assert_array_equal(o.received_from[b], rec["received_from"], err_msg=ctx)
    np.testing.assert_array_equal(o.transmitted_to[b], rec["transmitted_to"], err_msg=ctx)
    np.testing.assert_array_equal(o.steps_taken[b], rec["steps_taken"], err_msg=ctx)
    np.testing.assert_array_equal(o.has_taken_action[b], rec["has_taken_action"], err_msg=ctx)
    np.testing.assert_array_equal(o.adj[b], rec["adj"], err_msg=ctx)
    np.testing.assert_array_equal(two_hop(o.adj[b]), rec["two_hop"], err_msg=ctx)
    assert o.world_msgs[b] == rec["world_msgs"], ctx
    np.testing.assert_array_equal(o.active[b], rec["active"], err_msg=ctx)
    got = o.obs()[b]
    assert got.dtype == np.float32
    np.testing.assert_array_equal(got.view(np.uint32), rec["obs"].view(np.uint32), err_msg=ctx)


def _play_and_compare(ref, graph, N, seeds, *, dynamic=False, is_testing=False, heuristic=None,
                      ratio=0.0, heuristic_params=None, rng_seed=0):
    from oracle.ref_driver import MovementTap, run_reference_episode
    env = _mk_env(ref, graph, N, dynamic_graph=dynamic, is_testing=is_testing, heuristic=heuristic,
                  scripted_agents_ratio=ratio, heuristic_params=heuristic_params,
                  num_test_episodes=10 if is_testing else None)
    tap = MovementTap(env.world) if dynamic else None
    n_rounds = 0
    for seed in seeds:
        arng = np.random.default_rng(rng_seed + seed)
        table = arng.integers(0, 2, size=(64, N))
        # with a fixed graph= the reference resets on the graph object as it currently is
        # (mutated in place by earlier dynamic episodes)
        adj_pre, pos_pre = topology.graph_to_arrays(env.world.graph, N)
        snap0, rounds = run_reference_episode(env, seed, lambda r, i: table[r, i], tap=tap)
        o = BatchedEnvOracle(1, N, dynamic=dynamic, is_testing=is_testing, heuristic=heuristic)
        o.reset([0], adj_pre[None], pos_pre[None],
                np.array([snap0["source"]]), snap0["interested"][None], snap0["scripted"][None],
                move_offsets=snap0["move_offsets"][None] if dynamic else None)
        if dynamic:
            np.testing.assert_array_equal(o.pos[0], snap0["pos"])
        _compare_state(o, snap0, ctx=f"reset seed={seed}")
        for r, rec in enumerate(rounds):
            obs, rew, active, term, done = o.step(
                rec["actions"][None], move_offsets=rec["move_offsets"][None] if dynamic else None)
            ctx = f"seed={seed} round={r}"
            _compare_state(o, rec, ctx=ctx)
            np.testing.assert_array_equal(rew[0].view(np.uint64), rec["reward"].view(np.uint64), err_msg=ctx)
            np.testing.assert_array_equal(term[0], rec["terminated"], err_msg=ctx)
            assert o.episode_rewards_sum[0] == rec["episode_rewards_sum"], ctx
            np.testing.assert_array_equal(o.reward_vectorised(rec["actions"][None] >= 0)[0].view(np.uint64),
                                          rec["reward"].view(np.uint64), err_msg=ctx)
            ls = rec["logger_stats"]
            inf = o.info()
            assert inf["total_messages_transmitted"][0] == ls["total_messages_transmitted"]
            assert inf["covered"][0] / N == ls["coverage"]
            assert inf["messages_sent"][0] == ls["messages_sent"]
            assert inf["messages_received"][0] == ls["messages_received"]
            assert inf["n_neighbours"][0] == ls["n_neighbours"]
            assert inf["interested_agents"][0] == ls["interested_agents"]
            assert inf["coverage_interested_count"][0] == ls["coverage_interested_count"]
            assert inf["uninterested_with_message"][0] == ls["uninterested_with_message"]
            assert inf["episode_rewards_sum"][0] == ls["episode_rewards_sum"]
            n_rounds += 1
        if rounds:
            assert bool(done[0]) == (not rounds[-1]["active"].any())
    return n_rounds


@pytest.mark.parametrize("N,side,n_graphs,n_seeds", [(20, 0.6, 4, 6), (50, 1.0, 2, 3)])
def test_static_random_actions(N, side, n_graphs, n_seeds):
    ref = load_reference()
    total = 0
    for gs in range(n_graphs):
        g = topology.make_connected_graph(N, gs, side)
        total += _play_and_compare(ref, g, N, range(gs * 10, gs * 10 + n_seeds))
    assert total > 20


def test_dynamic_random_actions():
    ref = load_reference()
    total = 0
    for gs in range(3):
        g = topology.make_connected_graph(20, 100 + gs, 0.6)
        total += _play_and_compare(ref, g, 20, range(5), dynamic=True)
    assert total > 10


@pytest.mark.parametrize("heuristic", ["simple_broadcast", "silent", "broadcast_if_any_interested", "mpr"])
@pytest.mark.parametrize("dynamic", [False, True])
def test_mixed_scripted(heuristic, dynamic):
    ref = load_reference()
    for gs in range(2):
        g = topology.make_connected_graph(20, 200 + gs, 0.6)
        _play_and_compare(ref, g, 20, range(4), heuristic=heuristic, ratio=0.4, dynamic=dynamic)


@pytest.mark.parametrize("heuristic", ["simple_broadcast", "mpr", "silent"])
def test_testing_mode_scripted_are_policy_stepped(heuristic, tmp_path):
    """is_testing=True: scripted agents are in env.agents and the heuristic overrides them
    (graph.py:244,340).  Testing mode loads graphs from graph_topologies/testing_N/."""
    ref = load_reference()
    topology.write_topology_dir(str(tmp_path), 20, 3, split="testing", first_seed=300, side=0.6)
    with in_dir(str(tmp_path)):
        env = ref.GraphEnv(number_of_agents=20, radius=0.2, is_testing=True, heuristic=heuristic,
                           scripted_agents_ratio=0.5, num_test_episodes=4)
        from oracle.ref_driver import run_reference_episode
        for seed in range(4):
            table = np.random.default_rng(seed).integers(0, 2, size=(64, 20))
            snap0, rounds = run_reference_episode(env, seed, lambda r, i: table[r, i])
            o = BatchedEnvOracle(1, 20, is_testing=True, heuristic=heuristic)
            o.reset([0], snap0["adj"][None], snap0["pos"][None], np.array([snap0["source"]]),
                    snap0["interested"][None], snap0["scripted"][None])
            _compare_state(o, snap0, ctx=f"reset {seed}")
            for r, rec in enumerate(rounds):
                obs, rew, *_ = o.step(rec["actions"][None])
                _compare_state(o, rec, ctx=f"seed={seed} round={r}")
                np.testing.assert_array_equal(rew[0].view(np.uint64), rec["reward"].view(np.uint64))


@pytest.mark.parametrize("heuristic", ["mpr", "simple_broadcast", "silent", "broadcast_if_any_interested"])
@pytest.mark.parametrize("N,side", [(20, 0.6), (50, 1.0)])
def test_all_scripted_worlds(heuristic, N, side):
    """scripted_agents_ratio=1.0 worlds stepped directly with World.step."""
    ref = load_reference()
    for gs in range(6 if N == 20 else 2):
        g = topology.make_connected_graph(N, 400 + gs, side)
        rng, _ = ref.np_random(gs)
        w = ref.World(number_of_agents=N, radius=0.2, np_random=rng, graph=g,
                      scripted_agents_ratio=1.0, heuristic=heuristic)
        adj, pos = topology.graph_to_arrays(g)
        inter = np.array([a.is_interested for a in w.agents])
        scr = np.array([a.is_scripted for a in w.agents])
        assert scr.all()
        o = BatchedEnvOracle(1, N, heuristic=heuristic)
        o.reset([0], adj[None], pos[None], np.array([w.origin_agent]), inter[None], scr[None])
        for r in range(12):
            hm = np.array([bool(a.state.has_message) for a in w.agents])
            np.testing.assert_array_equal(o.has_message[0], hm, err_msg=f"g={gs} r={r}")
            np.testing.assert_array_equal(o.msgs[0], [a.messages_transmitted for a in w.agents])
            np.testing.assert_array_equal(
                o.received_from[0], np.stack([a.state.received_from for a in w.agents]).astype(np.int32))
            for a in w.agents:
                a.action = None
            w.step()
            o._world_step(np.array([0]), np.full((1, N), -1, dtype=np.int8))


def test_mpr_select_matches_reference_function():
    ref = load_reference()
    from graph_env.env.utils.heuristics.mpr import mpr_heuristic
    for N, side, gs in [(20, 0.6, 500), (20, 0.6, 501), (50, 1.0, 502)]:
        g = topology.make_connected_graph(N, gs, side)
        rng, _ = ref.np_random(1)
        w = ref.World(number_of_agents=N, radius=0.2, np_random=rng, graph=g)
        adj, _ = topology.graph_to_arrays(g)
        for a in w.agents:
            want = mpr_heuristic(a).astype(bool)
            np.testing.assert_array_equal(mpr_select(adj, a.id), want)


def test_probabilistic_heuristics_host_fed_bits():
    """probabilistic_* draw from the global numpy RNG (heuristics/core.py:27,40); parity is
    defined on host-fed bits: replay the global stream in the reference's call order."""
    ref = load_reference()
    N = 20
    for heuristic, prob in [("probabilistic_gossip", 0.5), ("probabilistic_relay", 0.6)]:
        g = topology.make_connected_graph(N, 600, 0.6)
        rng, _ = ref.np_random(5)
        np.random.seed(1234)
        w = ref.World(number_of_agents=N, radius=0.2, np_random=rng, graph=g,
                      scripted_agents_ratio=1.0, heuristic=heuristic, heuristic_params={"prob": prob})
        adj, pos = topology.graph_to_arrays(g)
        inter = np.array([a.is_interested for a in w.agents])
        scr = np.array([a.is_scripted for a in w.agents])
        replay = np.random.RandomState(1234)

        def bits_for_round(o):
            if heuristic == "probabilistic_gossip":
                gb = np.zeros((1, N), dtype=np.int8)
                for i in range(N):
                    if not o.has_taken_action[0, i]:
                        gb[0, i] = replay.binomial(1, prob)
                return dict(gossip_bits=gb)
            rb = np.zeros((1, N, N), dtype=np.int8)
            for i in range(N):
                rb[0, i] = replay.binomial(1, prob, size=(N,))
            return dict(relay_bits=rb)

        o = BatchedEnvOracle(1, N, heuristic=heuristic)
        # the reset's forced step consumes draws too
        o.has_taken_action[:] = False
        kw = bits_for_round(o)
        o.reset([0], adj[None], pos[None], np.array([w.origin_agent]), inter[None], scr[None], **kw)
        for r in range(8):
            np.testing.assert_array_equal(o.has_message[0], [bool(a.state.has_message) for a in w.agents])
            np.testing.assert_array_equal(o.msgs[0], [a.messages_transmitted for a in w.agents])
            for a in w.agents:
                a.action = None
            w.step()
            o._world_step(np.array([0]), np.full((1, N), -1, dtype=np.int8), **bits_for_round(o))


def test_reference_unit_test_known_answers():
    """The reference's own known answers (tests/unit/.../test_core.py:97-169) on the oracle's masks."""
    g = nx.Graph()
    g.add_edges_from([(0, 1), (0, 2), (0, 3), (0, 4), (3, 4), (2, 5), (2, 6), (3, 7), (7, 8), (7, 9),
                      (8, 9), (4, 11), (3, 10)])
    for n in g.nodes:
        g.nodes[n]["pos"] = (0, 0)
    adj, _ = topology.graph_to_arrays(g, 12)
    th = two_hop(adj)
    assert th[0].astype(int).tolist() == [0, 1, 1, 1, 1, 1, 1, 1, 0, 0, 1, 1]
    assert th[3].astype(int).tolist() == [1, 1, 1, 0, 1, 0, 0, 1, 1, 1, 1, 1]
    assert th[11].astype(int).tolist() == [1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0]
    g2 = nx.Graph()
    g2.add_nodes_from(range(12))
    g2.add_edges_from([(0, 1), (0, 3), (1, 5), (2, 3), (2, 5), (2, 6), (5, 6), (3, 4), (3, 7), (7, 8),
                       (4, 11), (3, 10), (10, 11)])
    adj2, _ = topology.graph_to_arrays(g2, 12)
    th2 = two_hop(adj2)
    assert th2[0].astype(int).tolist() == [0, 1, 1, 1, 1, 1, 0, 1, 0, 0, 1, 0]
    assert th2[9].astype(int).tolist() == [0] * 12
    assert th2[10].astype(int).tolist() == [1, 0, 1, 1, 1, 0, 0, 1, 0, 0, 0, 1]


def test_reset_chain_matches_reference(tmp_path):
    ref = load_reference()
    N = 20
    g = topology.make_connected_graph(N, 0, 0.6)
    for ratio, heuristic in [(0.0, None), (0.3, None), (0.5, "simple_broadcast")]:
        env = _mk_env(ref, g, N, scripted_agents_ratio=ratio, heuristic=heuristic)
        for seed in (0, 9, 42, 123, 999):
            env.reset(seed=seed)
            rng, _ = reset_chain.make_np_random(seed)
            t = reset_chain.training_reset(rng, N, n_graphs=0, scripted_agents_ratio=ratio)
            assert t.source == env.world.origin_agent
            np.testing.assert_array_equal(t.interested, [a.is_interested for a in env.world.agents])
            np.testing.assert_array_equal(t.scripted, [a.is_scripted for a in env.world.agents])
            want = np.random.RandomState(t.movement_seed).uniform(-1, 1)
            assert env.world.movement_np_random.uniform(-1, 1) == want
            # a second, unseeded reset continues the same generator
            env.reset()
            t2 = reset_chain.training_reset(rng, N, n_graphs=0, scripted_agents_ratio=ratio)
            assert t2.source == env.world.origin_agent
            np.testing.assert_array_equal(t2.interested, [a.is_interested for a in env.world.agents])
            np.testing.assert_array_equal(t2.scripted, [a.is_scripted for a in env.world.agents])


def test_reset_chain_graph_draw_and_testing_mode(tmp_path):
    ref = load_reference()
    N = 20
    topology.write_topology_dir(str(tmp_path), N, 5, split="training", first_seed=0, side=0.6)
    topology.write_topology_dir(str(tmp_path), N, 4, split="testing", first_seed=50, side=0.6)
    with in_dir(str(tmp_path)):
        env = ref.GraphEnv(number_of_agents=N, radius=0.2)
        train_paths = list(env.world.train_graphs)
        for seed in (1, 2, 3):
            env.reset(seed=seed)
            rng, _ = reset_chain.make_np_random(seed)
            t = reset_chain.training_reset(rng, N, n_graphs=len(train_paths))
            assert train_paths[t.graph_index] == env.world.selected_graph
            assert t.source == env.world.origin_agent
        envt = ref.GraphEnv(number_of_agents=N, radius=0.2, is_testing=True, num_test_episodes=3)
        stream = reset_chain.TestingResetStream(N, 3, 4)
        test_paths = topology.list_topology_dir(".", N, "testing")
        # the reference constructor already consumed two resets (core.py:190, graph.py:117)
        rng, _ = reset_chain.make_np_random(0)
        stream.next(rng); stream.next(rng)
        for k in range(5):
            envt.reset(seed=k)
            t = stream.next(rng)
            adj, _ = topology.graph_to_arrays(topology.load_graph(test_paths[t.graph_index]))
            np.testing.assert_array_equal(adj, np.stack([a.one_hop_neighbours_ids for a in envt.world.agents]).astype(bool))
            assert t.source == envt.world.origin_agent
            np.testing.assert_array_equal(t.interested, [a.is_interested for a in envt.world.agents])
