"""Graph topologies: synthetic generation, packing into the device layout, and IO in the
reference's own on-disk format.

* The reference stores one pickled ``networkx.Graph`` per file, node attribute
  ``pos=[x, y]``, under ``graph_topologies/{training,testing}_{N}/`` and globs them
  (reference graph_env/env/utils/core.py:165-175, 450-457).
* The reference's own generator is ``nx.random_geometric_graph`` rejected until
  connected (core.py:440-447); :func:`make_connected_graph` is the same recipe with a
  configurable square side (L=1.0 for N=50/200, L=0.6 for N=20).
* Device layout: adjacency as per-node bitmask rows ``uint32[N][W]``, ``W = ceil(N/32)``;
  bit ``j`` of row ``i`` set iff ``{i, j}`` is an edge.  Positions ``float64[N][2]``.
"""
from __future__ import annotations

import glob
import os
import pickle

import networkx as nx
import numpy as np

RADIUS_OF_INFLUENCE = 0.20      # reference constants.py:1


def words_per_row(n_nodes: int) -> int:
    return (n_nodes + 31) // 32


def default_square_side(n_nodes: int) -> float:
    return 0.6 if n_nodes <= 20 else 1.0


def make_connected_graph(n_nodes: int, graph_seed: int, side: float | None = None,
                         radius: float = RADIUS_OF_INFLUENCE) -> nx.Graph:
    """Connected random geometric graph, nodes 0..N-1 with ``pos=[x, y]``."""
    side = default_square_side(n_nodes) if side is None else side
    rng = np.random.default_rng(graph_seed)
    while True:
        p = rng.uniform(0.0, side, size=(n_nodes, 2))
        g = nx.random_geometric_graph(
            n_nodes, radius, pos={i: [float(p[i, 0]), float(p[i, 1])] for i in range(n_nodes)})
        if nx.is_connected(g):
            return g


def _is_connected(adj: np.ndarray) -> bool:
    n = adj.shape[0]
    seen = np.zeros(n, dtype=bool)
    seen[0] = True
    frontier = seen.copy()
    while frontier.any():
        nxt = adj[frontier].any(axis=0) & ~seen
        seen |= nxt
        frontier = nxt
    return bool(seen.all())


def connected_geometric_arrays(n_nodes: int, graph_seed: int, side: float | None = None,
                               radius: float = RADIUS_OF_INFLUENCE):
    """numpy twin of :func:`make_connected_graph`: -> (adj bool [N,N], pos float64 [N,2])."""
    side = default_square_side(n_nodes) if side is None else side
    rng = np.random.default_rng(graph_seed)
    eye = np.eye(n_nodes, dtype=bool)
    while True:
        p = rng.uniform(0.0, side, size=(n_nodes, 2))
        d = p[:, None, :] - p[None, :, :]
        adj = ((d * d).sum(-1) <= radius * radius) & ~eye
        if _is_connected(adj):
            return adj, p


def graph_to_arrays(g: nx.Graph, n_nodes: int | None = None):
    """-> (adj bool [N,N], pos float64 [N,2]).  Node labels must be 0..N-1."""
    n = g.number_of_nodes() if n_nodes is None else n_nodes
    adj = np.zeros((n, n), dtype=bool)
    for u, v in g.edges():
        if u != v:
            adj[u, v] = adj[v, u] = True
    pos = np.zeros((n, 2), dtype=np.float64)
    for i in range(n):
        p = g.nodes[i].get("pos", (0.0, 0.0))
        pos[i, 0], pos[i, 1] = float(p[0]), float(p[1])
    return adj, pos


def arrays_to_graph(adj: np.ndarray, pos: np.ndarray) -> nx.Graph:
    """Inverse of :func:`graph_to_arrays`: networkx graph with nodes 0..N-1 and ``pos=[x, y]``."""
    n = adj.shape[0]
    g = nx.Graph()
    for i in range(n):
        g.add_node(i, pos=[float(pos[i, 0]), float(pos[i, 1])])
    iu, ju = np.nonzero(np.triu(adj, 1))
    g.add_edges_from(zip(iu.tolist(), ju.tolist()))
    return g


def pack_adjacency(adj: np.ndarray) -> np.ndarray:
    """bool [..., N, N] -> uint32 [..., N, W] bitmask rows (bit j of word j//32)."""
    n = adj.shape[-1]
    w = words_per_row(n)
    pad = w * 32 - n
    a = np.concatenate([adj, np.zeros(adj.shape[:-1] + (pad,), dtype=bool)], axis=-1) if pad else adj
    a = a.reshape(adj.shape[:-1] + (w, 32)).astype(np.uint32)
    weights = (np.uint32(1) << np.arange(32, dtype=np.uint32))
    return (a * weights).sum(axis=-1, dtype=np.uint64).astype(np.uint32)


def unpack_adjacency(packed: np.ndarray, n_nodes: int) -> np.ndarray:
    """uint32 [..., N, W] -> bool [..., N, N]."""
    bits = (packed[..., None] >> np.arange(32, dtype=np.uint32)) & np.uint32(1)
    bits = bits.reshape(packed.shape[:-1] + (packed.shape[-1] * 32,))
    return bits[..., :n_nodes].astype(bool)


def pack_mask(mask: np.ndarray) -> np.ndarray:
    """bool [..., N] -> uint32 [..., W]."""
    return pack_adjacency(mask[..., None, :])[..., 0, :]


def save_graph(graph: nx.Graph, path: str) -> None:
    """Same format as reference core.py:455-457."""
    with open(path, "wb") as f:
        pickle.dump(graph, f)


def load_graph(path: str) -> nx.Graph:
    """Same format as reference core.py:450-452."""
    with open(path, "rb") as f:
        return pickle.load(f)


def write_topology_dir(root: str, n_nodes: int, count: int, split: str = "training",
                       first_seed: int = 0, side: float | None = None) -> list[str]:
    """Write ``count`` graphs to ``root/graph_topologies/{split}_{N}/g{seed:05d}.gpickle``."""
    d = os.path.join(root, "graph_topologies", f"{split}_{n_nodes}")
    os.makedirs(d, exist_ok=True)
    paths = []
    for s in range(first_seed, first_seed + count):
        p = os.path.join(d, f"g{s:05d}.gpickle")
        save_graph(make_connected_graph(n_nodes, s, side), p)
        paths.append(p)
    return paths


def list_topology_dir(root: str, n_nodes: int, split: str = "training") -> list[str]:
    """The reference globs unsorted for training and sorted for testing (core.py:165-175)."""
    paths = glob.glob(os.path.join(root, "graph_topologies", f"{split}_{n_nodes}", "*"))
    return sorted(paths) if split == "testing" else paths


class GraphPool:
    """A set of G topologies of N nodes packed for the device: ``adj_bits uint32[G,N,W]``,
    ``pos float64[G,N,2]`` (numpy, host side; the engine uploads them once)."""

    def __init__(self, adj: np.ndarray, pos: np.ndarray):
        assert adj.ndim == 3 and adj.shape[1] == adj.shape[2] and pos.shape == adj.shape[:2] + (2,)
        self.n_nodes = adj.shape[1]
        self.adj = adj.astype(bool)
        self.pos = np.ascontiguousarray(pos, dtype=np.float64)
        self.adj_bits = np.ascontiguousarray(pack_adjacency(self.adj))

    def __len__(self):
        return self.adj.shape[0]

    @classmethod
    def synthetic(cls, n_nodes: int, count: int, first_seed: int = 0, side: float | None = None):
        """``count`` connected random geometric graphs, graph k from ``graph_seed = first_seed + k``.
        Same draws and same graphs as :func:`make_connected_graph` (checked in
        tests/test_host_logic.py), without building networkx objects."""
        arrs = [connected_geometric_arrays(n_nodes, s, side) for s in range(first_seed, first_seed + count)]
        return cls(np.stack([a for a, _ in arrs]), np.stack([p for _, p in arrs]))

    @classmethod
    def from_graphs(cls, graphs):
        arrs = [graph_to_arrays(g) for g in graphs]
        return cls(np.stack([a for a, _ in arrs]), np.stack([p for _, p in arrs]))

    @classmethod
    def from_files(cls, paths):
        return cls.from_graphs([load_graph(p) for p in paths])

    def save_npz(self, path: str) -> None:
        np.savez_compressed(path, adj_bits=self.adj_bits, pos=self.pos, n_nodes=self.n_nodes)

    @classmethod
    def load_npz(cls, path: str):
        z = np.load(path)
        n = int(z["n_nodes"])
        return cls(unpack_adjacency(z["adj_bits"], n), z["pos"])
