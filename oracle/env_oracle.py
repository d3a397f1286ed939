"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the Melissa environment round.

Restates, vectorised over B independent episodes, exactly what the reference does in
``graph_env/env/utils/core.py`` (World) and ``graph_env/env/graph.py`` (GraphEnv):

* ``World.reset``            core.py:343-437   -> :meth:`BatchedEnvOracle.reset`
* ``World.step``             core.py:225-266   -> :meth:`BatchedEnvOracle._world_step`
* ``World.relay_message``    core.py:268-279   -> inside ``_world_step`` (ordered pass)
* ``move_graph/update_position`` core.py:281-319 -> :meth:`_move`
* 1-/2-hop masks             core.py:321-341   -> :func:`two_hop`
* heuristics                 heuristics/core.py:12-62, heuristics/mpr.py:7-72 -> ``_scripted``
* ``GraphEnv.step`` round assembly / TTL / active set  graph.py:303-359 -> :meth:`step`
* ``GraphEnv._execute_world_step`` / ``reward``        graph.py:361-463 -> :meth:`step`, :meth:`_reward`
* obs rows                   graph.py:254-271  -> :meth:`obs`
* ``get_info``               graph.py:149-179  -> :meth:`info`

PINNED: checked against the unmodified reference (oracle/ref_loader.py) in
tests/test_oracle_env_vs_reference.py and against tests/golden/*.npz generated from it.

Not product code: melissa_b200 never imports this module.
"""
from __future__ import annotations

import numpy as np

RADIUS = 0.20            # constants.py:1
MOVE_STEP = 0.06         # constants.py:4
TTL = 4                  # graph.py:332, selector.py:44

HEURISTICS = (None, "silent", "simple_broadcast", "broadcast_if_any_interested",
              "probabilistic_gossip", "probabilistic_relay", "mpr")


def two_hop(adj: np.ndarray) -> np.ndarray:
    """core.py:334-341: 2-hop row = 1-hop | OR of neighbours' 1-hop rows, self cleared.
    adj: bool [..., N, N]."""
    a = adj.astype(np.uint8)
    reach = (a @ a) > 0                       # i -> k -> j
    th = adj | reach
    n = adj.shape[-1]
    th = th & ~np.eye(n, dtype=bool)
    return th


def mpr_select(adj: np.ndarray, a: int) -> np.ndarray:
    """heuristics/mpr.py:7-72 (OLSR MPR set of node ``a``) on a bool adjacency [N, N].
    Returns bool [N] relay mask.  Tie rules as shipped: a 2-hop node with exactly one
    provider marks that provider and covers ONLY that 2-hop node (mpr.py:45-48); the
    greedy loop picks the max-reach neighbour, ties -> largest node id (``max()`` over a
    dict iterates keys, mpr.py:65)."""
    n = adj.shape[0]
    n1 = adj[a].copy()
    n2 = n1 | adj[n1].any(axis=0)
    n2[a] = False
    strict = n2 & ~n1
    nbrs = np.flatnonzero(n1)
    mpr = np.zeros(n, dtype=bool)
    covered = np.zeros(n, dtype=bool)
    for t in np.flatnonzero(strict):
        prov = [int(x) for x in nbrs if adj[x, t]]
        if len(prov) == 1:
            mpr[prov[0]] = True
            covered[t] = True
    unc = strict & ~covered
    while unc.any():
        reach = {int(x): int((adj[x] & unc).sum()) for x in nbrs}
        m = max(reach.values())
        cands = [k for k, v in reach.items() if v == m]
        pick = cands[0] if len(cands) == 1 else max(cands)
        mpr[pick] = True
        unc = unc & ~adj[pick]
    return mpr


class BatchedEnvOracle:
    """B independent episodes on N-node graphs.  All randomness is supplied by the
    caller (reset tuples, per-round movement offsets, scripted random bits) so that the
    CUDA path and this oracle can be driven with identical inputs."""

    def __init__(self, B: int, N: int, *, dynamic: bool = False, is_testing: bool = False,
                 heuristic: str | None = None, keep_matrices: bool = True):
        assert heuristic in HEURISTICS
        self.B, self.N = B, N
        self.dynamic, self.is_testing, self.heuristic = dynamic, is_testing, heuristic
        self.keep_matrices = keep_matrices
        z = lambda *s, dt=bool: np.zeros(s, dtype=dt)
        self.adj = z(B, N, N)
        self.pos = z(B, N, 2, dt=np.float64)
        self.has_message = z(B, N)
        self.origin = z(B, N)
        self.interested = z(B, N)
        self.scripted = z(B, N)
        self.has_taken_action = z(B, N)
        self.truncated = z(B, N)
        self.active = z(B, N)
        self.steps_taken = z(B, N, dt=np.int32)
        self.msgs = z(B, N, dt=np.int32)
        self.action = np.full((B, N), -1, dtype=np.int8)      # agent.action, -1 == None
        self.received_from = z(B, N, N, dt=np.int32)          # [b, receiver, sender]
        self.transmitted_to = z(B, N, N, dt=np.int32)         # [b, sender, receiver]
        self.nin = z(B, N, dt=np.int32)                       # number_interested_neighbors (core.py:323-328)
        self.source = z(B, dt=np.int32)
        self.world_msgs = z(B, dt=np.int32)
        self.num_moves = z(B, dt=np.int32)
        self.episode_rewards_sum = z(B, dt=np.float64)
        self._mpr_cache: dict[bytes, np.ndarray] = {}

    # ------------------------------------------------------------------ reset
    def reset(self, ids, adj, pos, source, interested, scripted, *, move_offsets=None,
              gossip_bits=None, relay_bits=None):
        """core.py:388-437 + graph.py:222-248 for the episodes ``ids``.
        adj bool [n,N,N]; pos f64 [n,N,2]; source int [n]; interested/scripted bool [n,N].
        Performs the forced first world step (core.py:437) in which the source
        broadcasts (core.py:246)."""
        ids = np.asarray(ids, dtype=np.int64)
        n = len(ids)
        ar = np.arange(n)
        self.adj[ids] = adj
        self.pos[ids] = pos
        for name in ("has_message", "origin", "has_taken_action", "truncated", "active"):
            getattr(self, name)[ids] = False
        self.interested[ids] = interested
        self.scripted[ids] = scripted
        self.steps_taken[ids] = 0
        self.msgs[ids] = 0
        self.action[ids] = -1
        self.received_from[ids] = 0
        self.transmitted_to[ids] = 0
        self.nin[ids] = 0                      # Agent.reset -> __init__ zeroes it (core.py:68,73-83)
        self.source[ids] = source
        self.world_msgs[ids] = 0
        self.num_moves[ids] = 0
        self.episode_rewards_sum[ids] = 0.0
        self.origin[ids, source] = True
        self.has_message[ids, source] = True
        self.steps_taken[ids, source] = 1      # core.py:435
        act = np.full((n, self.N), -1, dtype=np.int8)
        self._world_step(ids, act, move_offsets, gossip_bits, relay_bits)
        act_mask = self.has_message[ids] & (True if self.is_testing else ~self.scripted[ids])
        self.active[ids] = act_mask            # graph.py:242-246
        return self

    # ------------------------------------------------------------- heuristics
    def _has_callback(self, ids):
        # action_callback = heuristic_fn if is_scripted else None (core.py:428-429)
        if self.heuristic is None:
            return np.zeros((len(ids), self.N), dtype=bool)
        return self.scripted[ids]

    def _mpr_masks(self, ids):
        out = np.zeros((len(ids), self.N, self.N), dtype=bool)   # [k, selector i, chosen relay]
        for k, b in enumerate(ids):
            adj = self.adj[b]
            key = None
            if not self.dynamic:
                key = adj.tobytes()
                hit = self._mpr_cache.get(key)
                if hit is not None:
                    out[k] = hit
                    continue
            m = np.stack([mpr_select(adj, a) for a in range(self.N)])
            if key is not None:
                self._mpr_cache[key] = m
            out[k] = m
        return out

    def _world_step(self, ids, act, move_offsets=None, gossip_bits=None, relay_bits=None):
        """core.py:225-266.  ``act`` int8 [n,N] with -1 == None is agent.action on entry."""
        N = self.N
        n = len(ids)
        S = self._has_callback(ids)
        act = act.copy()
        hta = self.has_taken_action[ids]
        hm = self.has_message[ids].copy()
        org = self.origin[ids]
        relays_for = np.zeros((n, N, N), dtype=bool)           # [k, agent, selector]
        h = self.heuristic
        if h == "simple_broadcast":                            # heuristics/core.py:12-17
            act = np.where(S, np.where(hta, 0, 1), act).astype(np.int8)
        elif h == "silent":                                    # heuristics/core.py:57-62
            act = np.where(S, 0, act).astype(np.int8)
        elif h == "broadcast_if_any_interested":               # heuristics/core.py:46-54
            act = np.where(S, (self.nin[ids] > 0).astype(np.int8), act).astype(np.int8)
        elif h == "probabilistic_gossip":                      # heuristics/core.py:20-28
            assert gossip_bits is not None
            act = np.where(S, np.where(hta, 0, gossip_bits), act).astype(np.int8)
        elif h == "probabilistic_relay":                       # heuristics/core.py:31-42
            assert relay_bits is not None
            mask = relay_bits.astype(bool) & self.adj[ids] & S[:, :, None]
            relays_for = mask.transpose(0, 2, 1).copy()
        elif h == "mpr":
            mask = self._mpr_masks(ids) & S[:, :, None]
            relays_for = mask.transpose(0, 2, 1).copy()
        # second scripted loop (core.py:236-243)
        any_rf = relays_for.any(axis=2) & S
        act = np.where(any_rf, 0, act).astype(np.int8)
        rec_rel = ((self.received_from[ids] > 0) & relays_for).any(axis=2)
        fire = any_rf & ~hta & (hm | org) & (rec_rel | org)
        act = np.where(fire, 1, act).astype(np.int8)
        # source override (core.py:246)
        ar = np.arange(n)
        src = self.source[ids]
        first = self.msgs[ids, src] == 0
        act[ar[first], src[first]] = 1
        # ordered relay pass (core.py:249-254, 268-279)
        adj = self.adj[ids]
        msgs = self.msgs[ids]
        rf = self.received_from[ids]
        tt = self.transmitted_to[ids]
        wm = self.world_msgs[ids]
        hta = hta.copy()
        for i in range(N):
            tx = (act[:, i] > 0) & hm[:, i]
            if not tx.any():
                continue
            row = adj[:, i, :] & tx[:, None]
            tt[:, i, :] += row
            wm += tx
            msgs[:, i] += tx
            hta[:, i] |= tx
            rf[:, :, i] += row
            hm |= row
        self.has_message[ids] = hm
        self.has_taken_action[ids] = hta
        self.msgs[ids] = msgs
        self.received_from[ids] = rf
        self.transmitted_to[ids] = tt
        self.world_msgs[ids] = wm
        if self.dynamic:
            assert move_offsets is not None
            self._move(ids, move_offsets)
        # clear scripted relays/actions (core.py:264-266)
        act = np.where(S, 0, act).astype(np.int8)
        self.action[ids] = act

    def _move(self, ids, off):
        """core.py:281-319: pos += off (off = 0.06*U(-1,1), x offsets [n,0,N] then y [n,1,N]);
        edges = pairs with dist <= 0.2 (nx.geometric_edges, fp64)."""
        p = self.pos[ids]
        p = np.stack([p[:, :, 0] + off[:, 0, :], p[:, :, 1] + off[:, 1, :]], axis=2)
        self.pos[ids] = p
        dx = p[:, :, None, 0] - p[:, None, :, 0]
        dy = p[:, :, None, 1] - p[:, None, :, 1]
        adj = (dx * dx + dy * dy) <= RADIUS * RADIUS
        adj &= ~np.eye(self.N, dtype=bool)
        self.adj[ids] = adj
        self.nin[ids] = (adj & self.interested[ids][:, None, :]).sum(axis=2)   # core.py:323-328

    # ------------------------------------------------------------------ round
    def step(self, actions, *, move_offsets=None, gossip_bits=None, relay_bits=None):
        """One full round for all B episodes (graph.py:303-359).
        ``actions`` int8 [B,N]; only entries of currently-active agents are read.
        Returns (obs f32 [B,N,8], reward f64 [B,N], active, terminated, done)."""
        ids = np.arange(self.B)
        acted = self.active.copy()
        self.steps_taken += acted                                  # graph.py:316-318
        act = np.where(acted, actions, -1).astype(np.int8)         # graph.py:362-365
        self._world_step(ids, act, move_offsets, gossip_bits, relay_bits)
        reward = self._reward(acted)
        # id-ordered fp64 accumulation (graph.py:378-389)
        for i in range(self.N):
            m = acted[:, i]
            self.episode_rewards_sum[m] += reward[m, i]
        self.num_moves += 1
        self.truncated |= acted & (self.steps_taken >= TTL)        # graph.py:330-334
        elig = True if self.is_testing else ~self.scripted
        self.active = self.has_message & ~self.truncated & elig    # graph.py:336-345
        done = ~self.active.any(axis=1)
        return self.obs(), reward, self.active.copy(), self.truncated.copy(), done

    def _reward(self, acted):
        """graph.py:402-463, fp64, evaluated on the post-step state."""
        n1 = self.adj
        n2 = two_hop(self.adj)
        I = self.interested[:, None, :]
        M = self.has_message[:, None, :]
        O = self.origin[:, None, :]
        t = (n2 & I).sum(2)
        c = (n2 & I & (M | O)).sum(2)
        d = n1.sum(2)
        pu_n = (n1 & ~I).sum(2)
        pc_n = (n1 & M).sum(2)
        u = (n1 & I & ~M & ~O).sum(2)
        i1 = (n1 & I).sum(2)
        B, N = self.B, self.N
        rew = np.zeros((B, N), dtype=np.float64)
        tx = self.action > 0
        for b in range(B):
            for i in np.flatnonzero(acted[b]):
                r = (int(c[b, i]) / int(t[b, i])) if t[b, i] > 0 else 0.0
                if tx[b, i]:
                    di = int(d[b, i])
                    pu = (int(pu_n[b, i]) / di) if di > 0 else 0
                    pc = (int(pc_n[b, i]) / di) if di else 0
                    r -= (pu + pc)
                else:
                    if u[b, i] > 0:
                        r -= int(u[b, i]) / int(i1[b, i])
                rew[b, i] = r
        return rew

    def reward_vectorised(self, acted):
        """Same arithmetic as :meth:`_reward` with numpy fp64 ops (for big B)."""
        n1 = self.adj
        n2 = two_hop(self.adj)
        I = self.interested[:, None, :]
        M = self.has_message[:, None, :]
        O = self.origin[:, None, :]
        f = lambda x: x.sum(2).astype(np.float64)
        t, c, d = f(n2 & I), f(n2 & I & (M | O)), f(n1)
        pu_n, pc_n, u, i1 = f(n1 & ~I), f(n1 & M), f(n1 & I & ~M & ~O), f(n1 & I)
        with np.errstate(divide="ignore", invalid="ignore"):
            cov = np.where(t > 0, c / t, 0.0)
            pen_tx = np.where(d > 0, pu_n / d, 0.0) + np.where(d > 0, pc_n / d, 0.0)
            pen_no = np.where(u > 0, u / i1, 0.0)
        tx = self.action > 0
        r = np.where(tx, cov - pen_tx, np.where(u > 0, cov - pen_no, cov))
        return np.where(acted, r, 0.0)

    # ------------------------------------------------------------ observables
    def obs(self):
        """graph.py:254-271: [x, y, deg, msgs_tx, action(None->0), interested,
        has_msg|origin, not scripted] as float32."""
        o = np.zeros((self.B, self.N, 8), dtype=np.float32)
        o[:, :, 0] = self.pos[:, :, 0]
        o[:, :, 1] = self.pos[:, :, 1]
        o[:, :, 2] = self.adj.sum(2)
        o[:, :, 3] = self.msgs
        o[:, :, 4] = np.where(self.action < 0, 0, self.action)
        o[:, :, 5] = self.interested
        o[:, :, 6] = self.has_message | self.origin
        o[:, :, 7] = ~self.scripted
        return o

    def info(self):
        """graph.py:149-179 as integer counts (+ the fp64 reward sum); the fractions the
        reference reports are count/N and count/num_interested."""
        hm = self.has_message
        return dict(
            total_messages_transmitted=self.world_msgs.copy(),
            covered=hm.sum(1),
            messages_sent=self.msgs.sum(1),
            messages_received=self.received_from.sum((1, 2)),
            n_neighbours=self.adj.sum((1, 2)),
            interested_agents=self.interested.sum(1),
            coverage_interested_count=(hm & self.interested).sum(1),
            uninterested_with_message=(hm & ~self.interested).sum(1),
            episode_rewards_sum=self.episode_rewards_sum.copy(),
            num_moves=self.num_moves.copy(),
        )
