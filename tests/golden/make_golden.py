"""Generates tests/golden/env_*.npz by running the UNMODIFIED reference
(/root/reference, via oracle/ref_loader.py) in the builder container.

    python tests/golden/make_golden.py

Each file holds E episodes of one configuration, flattened over rounds:
  per episode : adj0 bool[E,N,N], pos0 f64[E,N,2] (graph as it was when reset() was called),
                source, interested, scripted, reset_* (state right after reset),
                round_ptr i32[E+1] (episode e owns rounds round_ptr[e]:round_ptr[e+1])
  per round   : actions i8[R,N] (-1 = agent did not act), move_offsets f64[R,2,N] (dynamic),
                obs f32[R,N,8], reward f64[R,N], active/terminated/has_message bool[R,N],
                msgs i32[R,N], recv_count i32[R,N], world_msgs i32[R], rewards_sum f64[R],
                stats i64[R,8] + stats_f f64[R] (get_info counters)
The GPU box has no /root/reference: the -m gpu parity tests replay these files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from melissa_b200 import topology  # noqa: E402
from oracle.ref_driver import MovementTap, run_reference_episode  # noqa: E402
from oracle.ref_loader import in_dir, load_reference  # noqa: E402

STAT_KEYS = ["total_messages_transmitted", "messages_sent", "messages_received", "n_neighbours",
             "interested_agents", "coverage_interested_count", "uninterested_with_message"]


def record(name, N, side, graph_seeds, seeds_per_graph, *, dynamic=False, heuristic=None, ratio=0.0,
           is_testing=False, tmpdir=None):
    ref = load_reference()
    eps = []
    for gs in graph_seeds:
        g = topology.make_connected_graph(N, gs, side)
        if is_testing:
            os.makedirs(os.path.join(tmpdir, "graph_topologies", f"testing_{N}"), exist_ok=True)
            for f in os.listdir(os.path.join(tmpdir, "graph_topologies", f"testing_{N}")):
                os.remove(os.path.join(tmpdir, "graph_topologies", f"testing_{N}", f))
            topology.save_graph(g, os.path.join(tmpdir, "graph_topologies", f"testing_{N}", "g.gpickle"))
            with in_dir(tmpdir):
                env = ref.GraphEnv(number_of_agents=N, radius=0.2, is_testing=True, heuristic=heuristic,
                                   scripted_agents_ratio=ratio, num_test_episodes=4, dynamic_graph=dynamic)
        else:
            env = ref.GraphEnv(graph=g, number_of_agents=N, radius=0.2, dynamic_graph=dynamic,
                               heuristic=heuristic, scripted_agents_ratio=ratio)
        tap = MovementTap(env.world) if dynamic else None
        for s in range(seeds_per_graph):
            seed = 1000 * gs + s
            table = np.random.default_rng(seed).integers(0, 2, size=(64, N))
            if is_testing:
                with in_dir(tmpdir):
                    # testing mode reloads the graph file on every reset (core.py:357-359)
                    adj0, pos0 = topology.graph_to_arrays(g, N)
                    snap0, rounds = run_reference_episode(env, seed, lambda r, i: table[r, i], tap=tap)
            else:
                adj0, pos0 = topology.graph_to_arrays(env.world.graph, N)
                snap0, rounds = run_reference_episode(env, seed, lambda r, i: table[r, i], tap=tap)
            eps.append((adj0, pos0, snap0, rounds))
    E = len(eps)
    R = sum(len(r) for *_, r in eps)
    out = dict(
        n_nodes=np.int32(N), dynamic=np.bool_(dynamic), is_testing=np.bool_(is_testing),
        heuristic=np.str_(heuristic or ""), scripted_ratio=np.float64(ratio),
        adj0=np.stack([e[0] for e in eps]), pos0=np.stack([e[1] for e in eps]),
        source=np.array([e[2]["source"] for e in eps], dtype=np.int32),
        interested=np.stack([e[2]["interested"] for e in eps]),
        scripted=np.stack([e[2]["scripted"] for e in eps]),
        reset_obs=np.stack([e[2]["obs"] for e in eps]),
        reset_active=np.stack([e[2]["active"] for e in eps]),
        reset_has_message=np.stack([e[2]["has_message"] for e in eps]),
        reset_msgs=np.stack([e[2]["msgs"] for e in eps]).astype(np.int32),
        reset_recv_count=np.stack([e[2]["received_from"].sum(1) for e in eps]).astype(np.int32),
        reset_adj=np.stack([e[2]["adj"] for e in eps]),
        round_ptr=np.concatenate([[0], np.cumsum([len(e[3]) for e in eps])]).astype(np.int32),
    )
    if dynamic:
        out["reset_move_offsets"] = np.stack([e[2]["move_offsets"] for e in eps])
    rr = [r for *_, rs in eps for r in rs]
    out.update(
        actions=np.stack([r["actions"] for r in rr]),
        obs=np.stack([r["obs"] for r in rr]),
        reward=np.stack([r["reward"] for r in rr]),
        active=np.stack([r["active"] for r in rr]),
        terminated=np.stack([r["terminated"] for r in rr]),
        has_message=np.stack([r["has_message"] for r in rr]),
        msgs=np.stack([r["msgs"] for r in rr]).astype(np.int32),
        recv_count=np.stack([r["received_from"].sum(1) for r in rr]).astype(np.int32),
        adj=np.packbits(np.stack([r["adj"] for r in rr]), axis=-1),
        world_msgs=np.array([r["world_msgs"] for r in rr], dtype=np.int32),
        rewards_sum=np.array([r["episode_rewards_sum"] for r in rr], dtype=np.float64),
        stats=np.array([[r["logger_stats"][k] for k in STAT_KEYS] for r in rr], dtype=np.int64),
        stats_coverage=np.array([r["logger_stats"]["coverage"] for r in rr], dtype=np.float64),
        stats_cov_int_frac=np.array([r["logger_stats"]["coverage_interested_fraction"] for r in rr], dtype=np.float64),
    )
    if dynamic:
        out["move_offsets"] = np.stack([r["move_offsets"] for r in rr])
    path = os.path.join(HERE, f"env_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {E} episodes, {R} rounds -> {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        record("static_n20", 20, 0.6, range(6), 4)
        record("static_n50", 50, 1.0, range(3), 2)
        record("dynamic_n20", 20, 0.6, range(10, 14), 3, dynamic=True)
        record("dynamic_n50", 50, 1.0, range(10, 11), 2, dynamic=True)
        record("mixed_mpr_n20", 20, 0.6, range(20, 24), 3, heuristic="mpr", ratio=0.4)
        record("mixed_mpr_dynamic_n20", 20, 0.6, range(24, 26), 3, heuristic="mpr", ratio=0.4, dynamic=True)
        record("mixed_bcast_n20", 20, 0.6, range(30, 33), 3, heuristic="simple_broadcast", ratio=0.4)
        record("mixed_bcast_dynamic_n20", 20, 0.6, range(33, 35), 3, heuristic="simple_broadcast", ratio=0.4, dynamic=True)
        record("mixed_bii_dynamic_n20", 20, 0.6, range(35, 37), 3, heuristic="broadcast_if_any_interested", ratio=0.4, dynamic=True)
        record("mixed_silent_n20", 20, 0.6, range(37, 39), 2, heuristic="silent", ratio=0.3)
        record("testing_mpr_n20", 20, 0.6, range(40, 43), 3, heuristic="mpr", ratio=0.5, is_testing=True, tmpdir=tmp)
        record("testing_bcast_n20", 20, 0.6, range(43, 45), 3, heuristic="simple_broadcast", ratio=0.5, is_testing=True, tmpdir=tmp)


if __name__ == "__main__":
    main()
