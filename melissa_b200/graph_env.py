"""AEC facade with the reference's ``GraphEnv`` surface (graph_env/env/graph.py:18-496) over the CUDA
environment-round kernel, for drop-in use where code expects the PettingZoo AEC protocol
(``reset / step / observe / last / agents / agent_selection / rewards / terminations / infos``).

One agent acts per ``step()``; the world only advances when every agent of the round has acted
(graph.py:324-345).  All of the dynamics run on the device (``mls_env_reset`` / ``mls_env_step``
through :class:`CudaRoundStepper`, batch of one episode); this module only keeps the AEC bookkeeping
-- selection order, dead steps, reward accumulation -- which is host logic in the reference too.
For throughput use :class:`melissa_b200.batched_env.BatchedGraphEnv` directly: the facade exists for
interface parity, not speed.

The reset RNG chain is :mod:`melissa_b200.reset_chain` (graph.py:222-225, core.py:343-395).
"""
from __future__ import annotations

import glob

import numpy as np

from . import reset_chain, topology
from .topology import RADIUS_OF_INFLUENCE  # noqa: F401  (re-exported like the reference's constants)

NUMBER_OF_FEATURES = 5          # reference constants.py:2
TTL = 4                         # graph.py:332, selector.py:44


class _Space:
    """Minimal stand-in used when gymnasium is not installed (shape/dtype/bounds only)."""

    def __init__(self, kind, **kw):
        self.kind = kind
        self.__dict__.update(kw)

    def __repr__(self):
        return f"{self.kind}({', '.join(f'{k}={v}' for k, v in self.__dict__.items() if k != 'kind')})"


def _spaces():
    try:
        import gymnasium
        return gymnasium.spaces.Box, gymnasium.spaces.Discrete, gymnasium.spaces.Dict
    except ImportError:
        box = lambda low, high, shape, dtype: _Space("Box", low=low, high=high, shape=shape, dtype=dtype)
        disc = lambda n: _Space("Discrete", n=n)
        dct = lambda d: _Space("Dict", spaces=d)
        return box, disc, dct


class CudaRoundStepper:
    """One episode on the device: a :class:`BatchedGraphEnv` with a single slot."""

    def __init__(self, n_nodes, *, dynamic_graph, is_testing, heuristic, device="cuda"):
        from .batched_env import BatchedGraphEnv
        self.N = n_nodes
        self.dynamic = dynamic_graph
        self._mk = lambda pool: BatchedGraphEnv(1, n_nodes, pool, dynamic_graph=dynamic_graph, is_testing=is_testing,
                                                heuristic=heuristic, device=device, want_info=False)
        self.env = None
        self.device = device

    def _adj(self):
        if self.dynamic:
            bits = self.env.adj.cpu().numpy().view(np.uint32)[0]
            return topology.unpack_adjacency(bits, self.N)
        return self._static_adj

    def _pos(self):
        return self.env.pos.cpu().numpy()[0] if self.dynamic else self._static_pos

    def has_taken_action(self):
        return self.env.flags()["has_taken_action"][0]

    def reset(self, adj, pos, source, interested, scripted, move_offsets=None, gossip_bits=None, relay_bits=None):
        from .batched_env import ResetTuplesDevice
        pool = topology.GraphPool(adj[None], pos[None])
        self.env = self._mk(pool)                       # a fresh single-slot environment on the given topology
        self._static_adj, self._static_pos = adj.copy(), pos.copy()
        tup = ResetTuplesDevice(np.zeros(1, np.int32), np.array([source], np.int32), interested[None], scripted[None],
                                self.N, self.device)
        obs, active = self.env.reset(tup, move_offsets=None if move_offsets is None else move_offsets[None],
                                     gossip_bits=None if gossip_bits is None else gossip_bits[None],
                                     relay_bits=None if relay_bits is None else relay_bits[None])
        return dict(obs=obs.cpu().numpy()[0], active=active.cpu().numpy()[0].astype(bool), adj=self._adj(), pos=self._pos())

    def step(self, actions, move_offsets=None, gossip_bits=None, relay_bits=None):
        obs, rew, active, term, done = self.env.step(np.asarray(actions, dtype=np.int8)[None],
                                                     move_offsets=None if move_offsets is None else move_offsets[None],
                                                     gossip_bits=None if gossip_bits is None else gossip_bits[None],
                                                     relay_bits=None if relay_bits is None else relay_bits[None])
        return dict(obs=obs.cpu().numpy()[0], reward=rew.cpu().numpy()[0], active=active.cpu().numpy()[0].astype(bool),
                    terminated=term.cpu().numpy()[0].astype(bool), done=bool(done.cpu().numpy()[0]), adj=self._adj(),
                    pos=self._pos())

    def info(self):
        inf = self.env.info()
        return {k: (v[0].item() if hasattr(v[0], "item") else v[0]) for k, v in inf.items()}


class _Selector:
    """Per-round agent ordering (reference utils/selector.py:1-52): agents are offered in id order, once
    per round, while they are active and have been offered fewer than four times."""

    def __init__(self, agents):
        self.reinit(agents)

    def reinit(self, agents):
        self.state = {a: {"steps": 0, "active": False, "selected_round": False} for a in agents}

    def next(self):
        for a, st in self.state.items():
            if st["active"] and not st["selected_round"]:
                st["steps"] += 1
                st["selected_round"] = True
                return a
        return False

    def disable(self, agent):
        self.state[agent]["active"] = False

    def enable(self, agents, on_reset=False, source_agent=None):
        if on_reset:
            self.state[source_agent]["steps"] += 1
        for a in agents:
            self.state[a]["active"] = self.state[a]["steps"] < TTL

    def start_new_round(self):
        for st in self.state.values():
            st["selected_round"] = False


class GraphEnv:
    """Same constructor as the reference (graph.py:25-42).  ``stepper`` is the object that advances the
    world by one round; by default the CUDA kernel (tests of the host logic may inject another)."""

    metadata = {"render_modes": ["human"], "name": "graph_environment", "is_parallelizable": False}

    def __init__(self, graph=None, render_mode=None, number_of_agents=10, radius=10, max_cycles=100, device="cuda",
                 local_ratio=None, scripted_agents_ratio=0.0, heuristic=None, heuristic_params=None, is_testing=False,
                 random_graph=False, dynamic_graph=False, all_agents_source=False, num_test_episodes=None, stepper=None):
        from ._lib import HEURISTIC_IDS
        if not (0.0 <= scripted_agents_ratio <= 1.0):
            raise ValueError("`scripted_agents_ratio` must be in [0.0, 1.0].")          # core.py:143-144
        if scripted_agents_ratio == 0.0 and heuristic is not None:
            raise ValueError("If `scripted_agents_ratio` is 0.0, no heuristic can be set.")
        if heuristic is not None and heuristic not in HEURISTIC_IDS:
            raise ValueError(f"Unknown heuristic policy: {heuristic}")
        if heuristic_params is not None and not isinstance(heuristic_params, dict):
            raise ValueError("Heuristic parameters must be a dictionary.")
        self.heuristic_params = dict(heuristic_params or {})
        if heuristic in ("probabilistic_gossip", "probabilistic_relay") and "prob" not in self.heuristic_params:
            raise TypeError(f"{heuristic}() missing 1 required positional argument: 'prob'")     # what the reference's partial raises
        self.random_graph = bool(random_graph)
        self.device, self.render_mode, self.local_ratio, self.radius = device, render_mode, local_ratio, radius
        self.number_of_agents = N = number_of_agents
        self.max_cycles = max_cycles
        self.is_testing, self.dynamic_graph = is_testing, dynamic_graph
        self.scripted_agents_ratio, self.heuristic = scripted_agents_ratio, heuristic
        self.is_new_round = None
        self.seed()
        # topology source: a fixed graph object, or the reference's directories (core.py:165-175)
        self.is_graph_fixed = graph is not None
        if self.is_graph_fixed:
            self._graph_adj, self._graph_pos = topology.graph_to_arrays(graph, N)
            self._graph_paths = []
        elif self.random_graph and not is_testing:
            self._graph_paths = []                      # every training episode draws a fresh graph (core.py:375-376)
        else:
            split = "testing" if is_testing else "training"
            self._graph_paths = topology.list_topology_dir(".", N, split)
        self._test_stream = reset_chain.TestingResetStream(N, num_test_episodes or 0, len(self._graph_paths)) if is_testing else None
        self.stepper = stepper if stepper is not None else CudaRoundStepper(
            N, dynamic_graph=dynamic_graph, is_testing=is_testing, heuristic=heuristic, device=device)
        self.possible_agents = [str(i) for i in range(N)]
        self.agents = self.possible_agents[:]
        self.agent_name_mapping = {a: i for i, a in enumerate(self.possible_agents)}
        self._agent_selector = _Selector(self.possible_agents)
        Box, Discrete, Dict = _spaces()
        obs_dim = N * (2 + NUMBER_OF_FEATURES + 1) + 1
        self.observation_spaces = {a: Dict({
            "observation": Box(low=-1e6, high=1e6, shape=(obs_dim,), dtype=np.float32),
            "action_mask": Box(low=0, high=1, shape=(2,), dtype=np.int8)}) for a in self.possible_agents}
        self.action_spaces = {a: Discrete(2) for a in self.possible_agents}
        self.state_space = Box(low=-1e6, high=1e6, shape=(obs_dim,), dtype=np.float32)
        self.obs_matrix = np.zeros((N, 2 + NUMBER_OF_FEATURES + 1), dtype=np.float32)
        self.num_moves = 0
        self.current_actions = [None] * N
        self._skip_agent_selection = None
        # the reference constructor performs two unseeded resets (core.py:190, graph.py:117); keep the
        # RNG / test-episode cursor in the same place
        t0 = self._draw_reset_tuple()
        if self.random_graph and not is_testing:
            self._topology_for(t0)                      # World.__init__'s own reset draws (and discards) a graph too
        if heuristic in ("probabilistic_gossip", "probabilistic_relay"):
            self._scripted = t0.scripted.copy()         # ... and its forced first step draws the heuristics' bits
            self._heuristic_bits(np.zeros(N, dtype=bool))
        self.reset()

    # ------------------------------------------------------------------ plumbing
    def observation_space(self, agent):
        return self.observation_spaces[agent]

    def action_space(self, agent):
        return self.action_spaces[agent]

    def seed(self, seed=None):
        self.np_random, _ = reset_chain.make_np_random(seed)

    def state(self):
        return self.obs_matrix.reshape(-1)

    def render(self):
        return None

    def close(self):
        return None

    def _heuristic_bits(self, has_taken_action):
        """The probabilistic heuristics draw from numpy's GLOBAL stream, one scripted agent after the other in id order
        (heuristics/core.py:20-42, World.step core.py:226-235): probabilistic_gossip one binomial(1, prob) per scripted
        agent that has not taken its action yet, probabilistic_relay binomial(1, prob, size=N) per scripted agent (the
        kernel masks it with the agent's 1-hop row).  Same draws, fed to the kernel as bits."""
        N, prob = self.number_of_agents, self.heuristic_params.get("prob")
        if self.heuristic == "probabilistic_gossip":
            bits = np.zeros(N, dtype=np.uint8)
            for i in np.flatnonzero(self._scripted):
                if not has_taken_action[i]:
                    bits[i] = np.random.binomial(1, prob)
            return dict(gossip_bits=bits)
        if self.heuristic == "probabilistic_relay":
            m = np.zeros((N, N), dtype=bool)
            for i in np.flatnonzero(self._scripted):
                m[i] = np.random.binomial(1, prob, size=N).astype(bool)
            return dict(relay_bits=m)
        return {}

    def _draw_reset_tuple(self):
        N = self.number_of_agents
        if self.is_testing:
            t = self._test_stream.next(self.np_random, self.scripted_agents_ratio)
        else:
            fixed = self.is_graph_fixed or self.random_graph          # neither draws np_random.choice(train_graphs)
            t = reset_chain.training_reset(self.np_random, N, n_graphs=0 if fixed else len(self._graph_paths),
                                           scripted_agents_ratio=self.scripted_agents_ratio)
        return t

    def _topology_for(self, t):
        if self.random_graph and not self.is_testing:
            # core.py:375-376 + create_connected_graph (core.py:440-447): nx.random_geometric_graph(n, radius) -- positions
            # from networkx's default (Python global `random`) stream, exactly as the reference draws them -- until connected
            import networkx as nx
            while True:
                g = nx.random_geometric_graph(n=self.number_of_agents, radius=self.radius)
                if nx.is_connected(g):
                    break
            return topology.graph_to_arrays(g, self.number_of_agents)
        if self.is_graph_fixed:
            return self._graph_adj, self._graph_pos
        return topology.graph_to_arrays(topology.load_graph(self._graph_paths[t.graph_index]), self.number_of_agents)

    # ------------------------------------------------------------------ AEC protocol
    def reset(self, seed=None, return_info=False, options=None):
        if seed is not None:
            self.seed(seed)
        N = self.number_of_agents
        self.agents = self.possible_agents[:]
        self._agent_selector.reinit(self.agents)
        self.rewards = {a: 0.0 for a in self.agents}
        self._cumulative_rewards = {a: 0.0 for a in self.agents}
        self.terminations = {a: False for a in self.agents}
        self.truncations = {a: False for a in self.agents}
        self.infos = {a: {} for a in self.agents}
        self.num_moves = 0
        t = self._draw_reset_tuple()
        adj, pos = self._topology_for(t)
        self._movement_rng = np.random.RandomState(t.movement_seed)
        mo = reset_chain.movement_offsets(self._movement_rng, N) if self.dynamic_graph else None
        self._scripted = t.scripted.copy()
        out = self.stepper.reset(adj, pos, t.source, t.interested, t.scripted, mo, **self._heuristic_bits(np.zeros(N, dtype=bool)))
        self._after_world_step(out)
        self.origin_agent = t.source
        self._truncated = np.zeros(N, dtype=bool)
        self.agents = [str(i) for i in np.flatnonzero(out["active"])]                       # graph.py:242-245
        self._agent_selector.enable(self.agents, on_reset=True, source_agent=str(t.source))
        self.agent_selection = self._agent_selector.next()
        self.current_actions = [None] * N
        self._skip_agent_selection = None

    def _after_world_step(self, out):
        self.obs_matrix = np.array(out["obs"], dtype=np.float32)
        self._adj = out["adj"]
        if self.dynamic_graph and self.is_graph_fixed:
            # the reference mutates the caller's graph object in place (core.py:300-314): the next
            # episode starts from the moved topology
            self._graph_adj, self._graph_pos = out["adj"].copy(), np.array(out["pos"], dtype=np.float64)

    def observe(self, agent):
        idx = self.agent_name_mapping[agent]
        obs = np.concatenate([self.obs_matrix.reshape(-1), [idx]]).astype(np.float32)
        dead = self.terminations[agent] or self.truncations[agent]
        mask = np.array([0, 0] if dead else [1, 1], dtype=np.int8)
        info = self.infos[agent]
        info["env_step"], info["environment_step"], info["explicit_reset"] = self.num_moves, False, False
        aoh = self._adj[idx].copy()
        for k in np.flatnonzero(aoh):
            if self._truncated[k] and str(k) not in self.agents:
                aoh[k] = False
        info["active_one_hop_neighbors"] = aoh.astype(np.bool_)
        if len(self.agents) == 1 and all(self.terminations[a] for a in self.agents):
            self.is_new_round = False
            info["explicit_reset"] = True
        if self.is_new_round:
            info["environment_step"] = True
            self.is_new_round = False
        return {"observation": obs, "action_mask": mask}

    def last(self, observe=True):
        a = self.agent_selection
        return (self.observe(a) if observe else None, self._cumulative_rewards[a], self.terminations[a],
                self.truncations[a], self.infos[a])

    def _dead_step(self):
        agent = self.agent_selection
        for d in (self.terminations, self.truncations, self.rewards, self._cumulative_rewards, self.infos):
            del d[agent]
        self.agents.remove(agent)
        dead = [a for a in self.agents if self.terminations[a] or self.truncations[a]]
        if dead:
            if self._skip_agent_selection is None:
                self._skip_agent_selection = self.agent_selection
            self.agent_selection = dead[0]
        else:
            if self._skip_agent_selection is not None:
                self.agent_selection = self._skip_agent_selection
            self._skip_agent_selection = None

    def step(self, action):
        sel = self.agent_selection
        if self.terminations[sel] or self.truncations[sel]:
            if action is not None:
                raise ValueError("When an agent is dead, the only valid action is None")
            self._agent_selector.disable(sel)
            self._dead_step()
            return
        self.current_actions[self.agent_name_mapping[sel]] = action
        self._cumulative_rewards[sel] = 0
        self.agent_selection = self._agent_selector.next()
        if not self.agent_selection:                    # every agent of the round has acted: advance the world
            for a, r in self.rewards.items():
                self._cumulative_rewards[a] += r
            for a in self.rewards:
                self.rewards[a] = 0
            N = self.number_of_agents
            acts = np.array([-1 if a is None else int(a) for a in self.current_actions], dtype=np.int8)
            mo = reset_chain.movement_offsets(self._movement_rng, N) if self.dynamic_graph else None
            hb = self._heuristic_bits(self.stepper.has_taken_action()) if self.heuristic in ("probabilistic_gossip", "probabilistic_relay") else {}
            out = self.stepper.step(acts, mo, **hb)
            self._after_world_step(out)
            for a in self.agents:
                self.rewards[a] = float(out["reward"][int(a)])
            self.num_moves += 1
            for a in self.agents:                       # TTL (graph.py:330-334)
                i = int(a)
                if out["terminated"][i] and not self._truncated[i]:
                    self._truncated[i] = True
                    self.terminations[a] = True
            has_msg = self.obs_matrix[:, 6] > 0
            self.agents = [str(i) for i in range(N) if has_msg[i] and str(i) in self.terminations
                           and (self.is_testing or not self._scripted[i])]
            self._agent_selector.enable(self.agents)
            self._agent_selector.start_new_round()
            self.is_new_round = True
            self.agent_selection = self._agent_selector.next()
            self.current_actions = [None] * N
        self.infos[self.agent_selection] = self.get_info(self.agent_selection)
        dead = [a for a in self.agents if self.terminations[a] or self.truncations[a]]      # _deads_step_first
        if dead:
            self._skip_agent_selection = self.agent_selection
            self.agent_selection = dead[0]

    def get_info(self, agent):
        s = self.stepper.info()
        n_int = int(s["interested_agents"])
        return {"logger_stats": {
            "total_messages_transmitted": int(s["total_messages_transmitted"]),
            "coverage": int(s["covered"]) / self.number_of_agents,
            "messages_sent": int(s["messages_sent"]),
            "messages_received": float(s["messages_received"]),
            "n_neighbours": float(s["n_neighbours"]),
            "interested_agents": n_int,
            "coverage_interested_fraction": (int(s["coverage_interested_count"]) / n_int) if n_int > 0 else 0.0,
            "coverage_interested_count": int(s["coverage_interested_count"]),
            "uninterested_with_message": int(s["uninterested_with_message"]),
            "episode_rewards_sum": float(s["episode_rewards_sum"]),
        }}


def env(**kwargs):
    """``graph_env_v0.env(**kwargs)`` (reference graph.py:487-496; the PettingZoo wrappers it applies are
    argument checks only)."""
    return GraphEnv(**kwargs)
