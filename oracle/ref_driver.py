"""TEST INFRASTRUCTURE ONLY.  Drive the UNMODIFIED reference GraphEnv through its AEC API
and record one trace entry per completed round, in the batched layout the CUDA path and
``env_oracle`` use.  Used to pin the oracle and to generate ``tests/golden``.
"""
from __future__ import annotations

import numpy as np

from .ref_loader import load_reference


def world_snapshot(env):
    w = env.world
    N = w.num_agents
    g = lambda f: np.array([f(a) for a in w.agents])
    return dict(
        has_message=g(lambda a: bool(a.state.has_message)),
        origin=g(lambda a: bool(a.state.message_origin)),
        interested=g(lambda a: bool(a.is_interested)),
        scripted=g(lambda a: bool(a.is_scripted)),
        has_taken_action=g(lambda a: bool(a.state.has_taken_action)),
        steps_taken=g(lambda a: int(a.steps_taken or 0)),
        msgs=g(lambda a: int(a.messages_transmitted)),
        received_from=np.stack([np.asarray(a.state.received_from) for a in w.agents]).astype(np.int32),
        transmitted_to=np.stack([np.asarray(a.state.transmitted_to) for a in w.agents]).astype(np.int32),
        adj=np.stack([np.asarray(a.one_hop_neighbours_ids) for a in w.agents]).astype(bool),
        two_hop=np.stack([np.asarray(a.two_hop_neighbours_ids) for a in w.agents]).astype(bool),
        pos=np.array([[float(a.pos[0]), float(a.pos[1])] for a in w.agents], dtype=np.float64),
        world_msgs=int(w.messages_transmitted),
        source=int(w.origin_agent),
        obs=env.obs_matrix.copy(),
        # env.agents still lists agents flagged terminated this round until their dead step
        # (graph.py:304-310); "active" = the ones that will actually be asked to act next round
        active=np.isin(np.arange(N), [int(x) for x in env.agents
                                      if not (env.terminations[x] or env.truncations[x])]),
        episode_rewards_sum=float(env.episode_rewards_sum),
        num_moves=int(env.num_moves),
    )


class MovementTap:
    """Observes World.compute_random_movement (core.py:316-319) without changing it."""

    def __init__(self, world):
        self.log = []
        orig = world.compute_random_movement

        def tapped(step):
            ox, oy = orig(step)
            self.log.append(np.array([ox, oy], dtype=np.float64))
            return ox, oy

        world.compute_random_movement = tapped

    def pop(self):
        assert len(self.log) == 1, len(self.log)
        return self.log.pop()


def run_reference_episode(env, seed, action_fn, max_rounds=64, tap: MovementTap | None = None):
    """Reset ``env`` with ``seed`` and play one episode.  ``action_fn(round, agent_idx)``
    returns 0/1.  Returns (reset_snapshot, [round records])."""
    N = env.number_of_agents
    if tap is not None:
        tap.log.clear()
    env.reset(seed=seed)
    snap0 = world_snapshot(env)
    if tap is not None:
        snap0["move_offsets"] = tap.pop()
    rounds = []
    cur_actions = np.full(N, -1, dtype=np.int8)
    rnd = 0
    guard = 0
    while env.agents and env.agent_selection is not False and rnd < max_rounds:
        guard += 1
        assert guard < 100000
        agent = env.agent_selection
        if env.terminations[agent] or env.truncations[agent]:
            env.step(None)
            continue
        idx = int(agent)
        a = int(action_fn(rnd, idx))
        cur_actions[idx] = a
        before = env.num_moves
        env.step(a)
        if env.num_moves > before:
            rec = world_snapshot(env)
            rec["actions"] = cur_actions.copy()
            rew = np.zeros(N, dtype=np.float64)
            acted = cur_actions >= 0
            for i in np.flatnonzero(acted):
                rew[i] = float(env.rewards.get(str(i), 0.0))
            rec["reward"] = rew
            rec["terminated"] = np.array([bool(a_.truncated) for a_ in env.world.agents])
            if tap is not None:
                rec["move_offsets"] = tap.pop()
            rec["logger_stats"] = dict(env.get_info(None)["logger_stats"])   # graph.py:149-179 (pure)
            rounds.append(rec)
            cur_actions[:] = -1
            rnd += 1
    return snap0, rounds
