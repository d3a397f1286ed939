"""Host-side episode-reset sampling: which graph, source node, interested set and
scripted set an episode gets.

This is the reference's RNG chain (graph_env/env/graph.py:222-225 +
graph_env/env/utils/core.py:343-395, 200-223), reproduced draw for draw so that the
same seeds give the same episodes:

training mode
    np_random = Generator(PCG64(SeedSequence(seed)))        (gymnasium seeding)
    episode_seed  = np_random.integers(0, 1e9)
    ep            = RandomState(episode_seed)
    [graph        = np_random.choice(train_graphs)]          only if no fixed graph
    movement_seed = ep.randint(0, 1e9)
    source        = ep.randint(0, N)
    density       = ep.uniform(0.1, 1.0)
    interested    = ep.choice(N, size=int(density*N), replace=False)
    scripted      = set(np_random.choice(N, size=round(ratio*N), replace=False)) - {source if ratio<1}
testing mode
    episode seeds = RandomState(17).randint(0, 1e9) x num_test_episodes, used cyclically
    graph         = ep.choice(sorted(test_graphs)); movement_seed; source as above
    density       = [0.1 .. 1.0][(index_after_increment) % 10]

The draws are host work (a few numpy calls per episode); the device only ever sees the
resulting tuples.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

TEST_SEED_STREAM = 17           # reference core.py:184


def make_np_random(seed):
    """gymnasium.utils.seeding.np_random restated: Generator(PCG64(SeedSequence(seed)))."""
    ss = np.random.SeedSequence(seed)
    return np.random.Generator(np.random.PCG64(ss)), ss.entropy


@dataclass
class ResetTuple:
    graph_index: int            # index into the graph list (-1: fixed graph)
    source: int
    interested: np.ndarray      # bool [N]
    scripted: np.ndarray        # bool [N]
    movement_seed: int
    episode_seed: int
    interest_density: float


def _scripted_mask(np_random, n_nodes, ratio, source):
    n_scripted = int(round(ratio * n_nodes))
    chosen = set(int(i) for i in np_random.choice(n_nodes, size=n_scripted, replace=False))
    if ratio < 1.0:
        chosen.discard(int(source))
    m = np.zeros(n_nodes, dtype=bool)
    if chosen:
        m[list(chosen)] = True
    return m


def training_reset(np_random: np.random.Generator, n_nodes: int, *, n_graphs: int = 0,
                   scripted_agents_ratio: float = 0.0,
                   fixed_interest_density: float | None = None) -> ResetTuple:
    """One training-mode reset.  ``n_graphs == 0`` means a fixed graph (no graph draw)."""
    episode_seed = int(np_random.integers(0, 1e9))
    ep = np.random.RandomState(episode_seed)
    graph_index = -1
    if n_graphs > 0:
        graph_index = int(np_random.choice(n_graphs, replace=True))
    movement_seed = int(ep.randint(0, 1e9))
    source = int(ep.randint(0, n_nodes))
    density = float(ep.uniform(0.1, 1.0)) if fixed_interest_density is None else float(fixed_interest_density)
    k = int(density * n_nodes)
    idx = ep.choice(n_nodes, size=k, replace=False)
    interested = np.zeros(n_nodes, dtype=bool)
    interested[idx] = True
    scripted = _scripted_mask(np_random, n_nodes, scripted_agents_ratio, source)
    return ResetTuple(graph_index, source, interested, scripted, movement_seed, episode_seed, density)


class TestingResetStream:
    """Testing-mode resets (reference core.py:182-187, 348-370)."""

    __test__ = False

    def __init__(self, n_nodes: int, num_test_episodes: int, n_graphs: int):
        gen = np.random.RandomState(TEST_SEED_STREAM)
        self.seeds = [int(gen.randint(0, 1e9)) for _ in range(num_test_episodes)]
        self.n_nodes, self.n_graphs, self.index = n_nodes, n_graphs, 0

    def next(self, np_random: np.random.Generator, scripted_agents_ratio: float = 0.0) -> ResetTuple:
        if not self.seeds:
            raise ValueError("No test seeds have been generated! Check num_test_episodes.")
        episode_seed = self.seeds[self.index]
        self.index = (self.index + 1) % len(self.seeds)
        ep = np.random.RandomState(episode_seed)
        if self.n_graphs <= 0:
            raise ValueError("No test graphs found!")
        graph_index = int(ep.choice(self.n_graphs))
        movement_seed = int(ep.randint(0, 1e9))
        source = int(ep.randint(0, self.n_nodes))
        density = [i / 10.0 for i in range(1, 11)][self.index % 10]
        k = int(density * self.n_nodes)
        idx = ep.choice(self.n_nodes, size=k, replace=False)
        interested = np.zeros(self.n_nodes, dtype=bool)
        interested[idx] = True
        scripted = _scripted_mask(np_random, self.n_nodes, scripted_agents_ratio, source)
        return ResetTuple(graph_index, source, interested, scripted, movement_seed, episode_seed, density)


def movement_offsets(movement_rng: np.random.RandomState, n_nodes: int, step: float = 0.06) -> np.ndarray:
    """One round of node movement, reference core.py:316-319: N x-offsets then N
    y-offsets, each ``step * uniform(-1, 1)``.  Returns float64 [2, N]."""
    ox = [step * movement_rng.uniform(-1, 1) for _ in range(n_nodes)]
    oy = [step * movement_rng.uniform(-1, 1) for _ in range(n_nodes)]
    return np.array([ox, oy], dtype=np.float64)


def episode_pool(base_seed: int, count: int, n_nodes: int, n_graphs: int,
                 scripted_agents_ratio: float = 0.0):
    """Reset tuples for ``count`` vector-env slots the way tianshou seeds them: env ``i``
    gets ``seed = base_seed + i`` and performs one seeded reset.  Graph ``i % n_graphs``
    is used instead of a graph draw when ``n_graphs`` > 0 (synthetic pool).
    Returns arrays (graph_index i32[count], source i32[count], interested bool[count,N],
    scripted bool[count,N], movement_seed i64[count])."""
    gi = np.zeros(count, dtype=np.int32)
    src = np.zeros(count, dtype=np.int32)
    inter = np.zeros((count, n_nodes), dtype=bool)
    scr = np.zeros((count, n_nodes), dtype=bool)
    mv = np.zeros(count, dtype=np.int64)
    for i in range(count):
        rng, _ = make_np_random(base_seed + i)
        t = training_reset(rng, n_nodes, n_graphs=0, scripted_agents_ratio=scripted_agents_ratio)
        gi[i] = (i % n_graphs) if n_graphs > 0 else -1
        src[i], inter[i], scr[i], mv[i] = t.source, t.interested, t.scripted, t.movement_seed
    return gi, src, inter, scr, mv


def testing_episode_pool(count: int, n_nodes: int, n_graphs: int, num_test_episodes: int, *, seed: int | None = None,
                         scripted_agents_ratio: float = 0.0, skip: int = 2):
    """Reset tuples of ONE testing-mode environment (``is_testing=True``, reference core.py:182-187,348-370)
    for its next ``count`` episodes: episode seeds from ``RandomState(17)`` used cyclically, graph drawn
    by the episode RNG from the sorted test graphs, interest density ladder 0.1 .. 1.0.
    ``skip`` = resets already consumed by the reference constructor (core.py:190, graph.py:117).
    Returns the same arrays as :func:`episode_pool` plus the densities."""
    stream = TestingResetStream(n_nodes, num_test_episodes, n_graphs)
    rng, _ = make_np_random(seed)
    for _ in range(skip):
        stream.next(rng, scripted_agents_ratio)
    gi = np.zeros(count, dtype=np.int32)
    src = np.zeros(count, dtype=np.int32)
    inter = np.zeros((count, n_nodes), dtype=bool)
    scr = np.zeros((count, n_nodes), dtype=bool)
    mv = np.zeros(count, dtype=np.int64)
    dens = np.zeros(count, dtype=np.float64)
    for k in range(count):
        t = stream.next(rng, scripted_agents_ratio)
        gi[k], src[k], inter[k], scr[k], mv[k], dens[k] = t.graph_index, t.source, t.interested, t.scripted, t.movement_seed, t.interest_density
    return gi, src, inter, scr, mv, dens
