"""CPU: host-side logic and the C-ABI surface (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import networkx as nx
import numpy as np
import pytest
import torch

from melissa_b200 import _lib, reset_chain, topology

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "melissa_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mls_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.lib()
    assert L.mls_version() >= 100
    names = _declared_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/melissa_b200.h but not exported"
    assert set(_lib.EXPORTS) <= set(names)
    assert L.mls_words_per_row(20) == 1 and L.mls_words_per_row(50) == 2 and L.mls_words_per_row(200) == 8
    assert _lib.words_per_row(100) == L.mls_words_per_row(100) == 4


def test_ctypes_structs_match_the_header_layout():
    structs = ["MlsEnvDesc", "MlsEnvState", "MlsResetTuples", "MlsInfo", "MlsRoundInputs", "MlsRoundOutputs",
               "MlsNetDesc", "MlsNetWeights", "MlsForwardArgs"]
    prog = '#include <stdio.h>\n#include "melissa_b200.h"\nint main(){' + "".join(
        f'printf("{s} %zu\\n", sizeof({s}));' for s in structs) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")], text=True)
    sizes = dict(line.split() for line in out.strip().splitlines())
    for s in structs:
        assert int(sizes[s]) == C.sizeof(getattr(_lib, s)), s


def test_no_cpu_fallback():
    from melissa_b200.batched_env import BatchedGraphEnv
    pool = topology.GraphPool.synthetic(20, 2)
    with pytest.raises(_lib.MelissaLibraryError):
        BatchedGraphEnv(4, 20, pool, device="cpu")
    with pytest.raises(ValueError):
        BatchedGraphEnv(4, 20, pool, heuristic="does_not_exist", device="cpu")
    from melissa_b200.networks import HLDGNNetwork
    net = HLDGNNetwork(5, 128, 2, 4, 20, aggregator="max",
                       dueling_param=({"hidden_sizes": [128, 128]}, {"hidden_sizes": [128, 128]}))
    with pytest.raises(ValueError):
        net(np.zeros((2, 160), dtype=np.float32))           # wrong width: reference error (networks/common.py:27-29)
    with pytest.raises(ValueError):
        net(np.zeros(161, dtype=np.float32))
    with pytest.raises(_lib.MelissaLibraryError):
        net(np.zeros((2, 161), dtype=np.float32))           # CPU parameters: refuse, never fall back
    with pytest.raises(KeyError):
        HLDGNNetwork(5, 128, 2, 4, 20, aggregator="median", dueling_param=({"hidden_sizes": [128, 128]},) * 2)


def test_numpy_graph_generator_equals_the_networkx_recipe():
    for n, seeds in ((20, range(12)), (50, range(6)), (200, range(2))):
        for s in seeds:
            adj, pos = topology.connected_geometric_arrays(n, s)
            adj2, pos2 = topology.graph_to_arrays(topology.make_connected_graph(n, s))
            assert np.array_equal(adj, adj2) and np.array_equal(pos, pos2)
            assert nx.is_connected(nx.from_numpy_array(adj))


def test_bitmask_packing_and_io_roundtrip(tmp_path):
    from melissa_b200.batched_env import pack_bits
    rng = np.random.default_rng(0)
    for n in (12, 32, 33, 50, 64, 200):
        adj = rng.random((3, n, n)) < 0.2
        packed = topology.pack_adjacency(adj)
        assert packed.shape == (3, n, (n + 31) // 32)
        assert np.array_equal(topology.unpack_adjacency(packed, n), adj)
        W = _lib.words_per_row(n)
        pb = pack_bits(adj, W)
        assert pb.shape == (3, n, W)
        assert np.array_equal(pb[..., : packed.shape[-1]], packed) and not pb[..., packed.shape[-1]:].any()
    paths = topology.write_topology_dir(str(tmp_path), 20, 3, split="training", first_seed=7)
    assert sorted(os.path.basename(p) for p in paths) == ["g00007.gpickle", "g00008.gpickle", "g00009.gpickle"]
    g = topology.load_graph(paths[0])                      # the reference's own format: pickled networkx graph with pos
    assert isinstance(g, nx.Graph) and len(g.nodes[0]["pos"]) == 2
    pool = topology.GraphPool.from_files(sorted(paths))
    pool2 = topology.GraphPool.synthetic(20, 3, first_seed=7)
    assert np.array_equal(pool.adj, pool2.adj) and np.array_equal(pool.pos, pool2.pos)
    pool.save_npz(str(tmp_path / "p.npz"))
    pool3 = topology.GraphPool.load_npz(str(tmp_path / "p.npz"))
    assert np.array_equal(pool3.adj, pool.adj) and np.array_equal(pool3.pos, pool.pos)


def test_reset_chain_is_deterministic_and_seed_sensitive():
    a = reset_chain.episode_pool(9, 16, 20, 8, 0.3)
    b = reset_chain.episode_pool(9, 16, 20, 8, 0.3)
    c = reset_chain.episode_pool(10, 16, 20, 8, 0.3)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert not np.array_equal(a[1], c[1]) or not np.array_equal(a[2], c[2])
    assert np.array_equal(a[1][1:], c[1][:-1])               # env k of seed s == env k-1 of seed s+1
    gi, src, inter, scr, mv = a
    assert gi.tolist() == [k % 8 for k in range(16)]
    assert (scr.sum(1) <= 6).all() and not scr[np.arange(16), src].any()      # the source is never scripted (core.py:213-215)
    k = inter.sum(1)
    assert ((k >= 2) & (k <= 20)).all()                       # int(U(0.1,1)*20) interested nodes
    off = reset_chain.movement_offsets(np.random.RandomState(int(mv[0])), 20)
    assert off.shape == (2, 20) and np.abs(off).max() <= 0.06


def test_golden_fixtures_are_self_consistent():
    from golden_util import GoldenEnv, golden_env_files
    for p in golden_env_files():
        g = GoldenEnv(p)
        assert g.ptr[-1] == g.z["actions"].shape[0]
        act = g.z["actions"]
        assert set(np.unique(act)) <= {-1, 0, 1}
        # an agent acted (action >= 0) exactly when it was in the previous active set
        for e in range(g.E):
            prev = g.z["reset_active"][e]
            for r in range(g.ptr[e], g.ptr[e + 1]):
                assert np.array_equal(act[r] >= 0, prev)
                prev = g.z["active"][r]


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from melissa_b200.sharding import reduce_job, shard_tuples, whole_job_throughput
    pool = (np.arange(64), np.arange(64) * 2)
    mine = shard_tuples(pool, rank, 32, world)
    assert len(mine[0]) == 32
    ms, units = (10.0, 1000.0) if rank == 0 else (20.0, 3000.0)
    ms_max, total = reduce_job(ms, units)
    thr = whole_job_throughput(ms, units)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([mine[0][0], mine[1][0], ms_max, total, thr]))
    dist.destroy_process_group()


def test_world_size_2_gloo_sharding_and_reduction(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert r0[0] == 0 and r1[0] == 32 and r1[1] == 64        # disjoint slices of the shared pool
    for r in (r0, r1):
        assert r[2] == 20.0 and r[3] == 4000.0 and r[4] == 4000.0 / 0.020   # max over ranks, sum over ranks


def test_testing_episode_pool_follows_the_test_stream():
    gi, src, inter, scr, mv, dens = reset_chain.testing_episode_pool(25, 20, 7, num_test_episodes=10, seed=3)
    # the seeds are used cyclically: episode k and k+10 draw the same graph / source / movement seed ...
    assert np.array_equal(gi[:15], gi[10:25]) and np.array_equal(src[:15], src[10:25]) and np.array_equal(mv[:15], mv[10:25])
    # ... and, the ladder having period 10 as well, the same interest density
    assert np.allclose(dens[:15], dens[10:25])
    assert set(np.round(dens, 1)) == {0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0}
    assert np.array_equal(inter.sum(1), (dens * 20).astype(int))
    assert ((gi >= 0) & (gi < 7)).all() and not scr.any()
    s = reset_chain.TestingResetStream(20, 10, 7)
    rng, _ = reset_chain.make_np_random(3)
    s.next(rng); s.next(rng)
    t = s.next(rng)
    assert t.graph_index == gi[0] and t.source == src[0] and np.array_equal(t.interested, inter[0])
