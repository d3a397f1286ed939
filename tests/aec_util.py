"""Replay helper for tests/golden/aec_*.npz (AEC-level traces of the unmodified reference) through
melissa_b200.graph_env.GraphEnv, with either the CUDA round stepper or an oracle-backed one."""
import glob
import os

import numpy as np

from melissa_b200 import topology
from melissa_b200.graph_env import GraphEnv

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def aec_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "aec_*.npz")))


class OracleRoundStepper:
    """Test double for the device round: oracle/env_oracle.py behind the stepper interface."""

    def __init__(self, N, dynamic, is_testing, heuristic):
        self.N, self.dynamic, self.is_testing, self.heuristic = N, dynamic, is_testing, heuristic
        self.o = None

    def _pack(self, extra=None):
        o = self.o
        d = dict(obs=o.obs()[0], active=o.active[0].copy(), adj=o.adj[0].copy(), pos=o.pos[0].copy())
        d.update(extra or {})
        return d

    def has_taken_action(self):
        return self.o.has_taken_action[0].copy()

    def reset(self, adj, pos, source, interested, scripted, move_offsets=None, gossip_bits=None, relay_bits=None):
        from oracle.env_oracle import BatchedEnvOracle
        self.o = BatchedEnvOracle(1, self.N, dynamic=self.dynamic, is_testing=self.is_testing, heuristic=self.heuristic)
        self.o.reset([0], adj[None], pos[None], np.array([source]), interested[None], scripted[None],
                     move_offsets=None if move_offsets is None else move_offsets[None],
                     gossip_bits=None if gossip_bits is None else gossip_bits[None],
                     relay_bits=None if relay_bits is None else relay_bits[None])
        return self._pack()

    def step(self, actions, move_offsets=None, gossip_bits=None, relay_bits=None):
        obs, rew, active, term, done = self.o.step(np.asarray(actions, dtype=np.int8)[None],
                                                   move_offsets=None if move_offsets is None else move_offsets[None],
                                                   gossip_bits=None if gossip_bits is None else gossip_bits[None],
                                                   relay_bits=None if relay_bits is None else relay_bits[None])
        return self._pack(dict(reward=rew[0], terminated=term[0], done=bool(done[0])))

    def info(self):
        return {k: v[0] for k, v in self.o.info().items()}


def replay(path, make_stepper):
    z = np.load(path)
    N = int(z["n_nodes"])
    dynamic, heuristic, ratio = bool(z["dynamic"]), (str(z["heuristic"]) or None), float(z["scripted_ratio"])
    ptr = z["step_ptr"]
    n_steps = 0
    for e in range(len(ptr) - 1):
        g = topology.arrays_to_graph(z["adj0"][e], z["pos0"][e])
        env = GraphEnv(graph=g, number_of_agents=N, radius=0.2, dynamic_graph=dynamic, heuristic=heuristic,
                       scripted_agents_ratio=ratio, stepper=make_stepper(N, dynamic, False, heuristic))
        # the constructor's own (unseeded) reset has already moved a dynamic fixed graph, exactly as the
        # reference does; the golden episode starts from the recorded pre-reset topology
        env._graph_adj, env._graph_pos = z["adj0"][e].copy(), z["pos0"][e].copy()
        env.reset(seed=int(z["seed"][e]))
        for s in range(ptr[e], ptr[e + 1]):
            ctx = f"{os.path.basename(path)} episode {e} aec step {s - ptr[e]}"
            assert env.agents and env.agent_selection is not False, ctx
            assert int(env.agent_selection) == int(z["agent"][s]), ctx
            assert len(env.agents) == int(z["n_agents"][s]), ctx
            o, cum, term, trunc, info = env.last()
            np.testing.assert_array_equal(o["observation"].view(np.uint32), z["obs"][s].view(np.uint32), err_msg=ctx)
            np.testing.assert_array_equal(o["action_mask"], z["mask"][s], err_msg=ctx)
            assert o["observation"].dtype == np.float32 and o["action_mask"].dtype == np.int8
            assert np.float64(cum).view(np.uint64) == z["cum"][s].view(np.uint64), f"{ctx}: {cum} vs {z['cum'][s]}"
            assert bool(term) == bool(z["term"][s]) and bool(trunc) == bool(z["trunc"][s]), ctx
            assert info["env_step"] == z["env_step"][s], ctx
            assert info["environment_step"] == bool(z["environment_step"][s]), ctx
            assert info["explicit_reset"] == bool(z["explicit_reset"][s]), ctx
            np.testing.assert_array_equal(info["active_one_hop_neighbors"], z["aoh"][s], err_msg=ctx)
            a = int(z["action"][s])
            env.step(None if a < 0 else a)
            n_steps += 1
        assert not env.agents or env.agent_selection is False, f"{path} episode {e}: facade still has agents"
    return n_steps
