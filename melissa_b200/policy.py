"""Policy and collector surfaces of the rollout path.

Reference:
* tianshou ``DQNPolicy.forward / compute_q_value / exploration_noise / set_eps`` as driven by
  ``MultiAgentSharedPolicy.forward`` / ``.exploration_noise``
  (graph_env/env/utils/policies/multi_agent_managers/shared_policy.py:81-183): every agent-observation
  row goes through ONE shared Q-network, the greedy action is ``argmax`` and, with probability
  ``eps``, it is replaced by ``argmax(rand(2) + mask)``.
* ``MultiAgentCollector.collect`` (graph_env/env/utils/collectors/multi_agent_collector.py:89-353): the
  step counter is the number of agent transitions (``:274``), ``collect_speed`` their rate (``:343``),
  episode returns / lengths and the per-episode ``logger_stats`` are summarised at the end
  (``collector.py:15-36``).

Here the Q-network forward, the action selection and the environment round are CUDA kernels; these
classes only sequence them and gather statistics.  Training-side methods (``process_fn`` / ``learn``,
replay buffers) are out of scope of the accelerated path.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np
import torch

from .batched_env import BatchedGraphEnv, ResetTuplesDevice
from .networks.common import DGNBase


class DQNPolicy:
    """The slice of tianshou's ``DQNPolicy`` the rollout uses.  ``model`` is one of the three networks."""

    def __init__(self, model: DGNBase, eps: float = 0.0, seed: int = 9):
        self.model = model
        self.eps = float(eps)
        self.seed = int(seed)
        self._calls = 0

    def set_eps(self, eps: float) -> None:
        self.eps = float(eps)

    def state_dict(self):
        # reference checkpoints are DQNPolicy.state_dict(): keys prefixed "model." (and "model_old.")
        return {f"model.{k}": v for k, v in self.model.state_dict().items()}

    def load_state_dict(self, sd):
        self.model.load_state_dict({k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")})

    def forward(self, obs, mask=None):
        """obs [bs, 8N+1] (numpy or tensor) -> dict(logits [bs,2] tensor, act [bs] numpy int64), greedy.
        ``mask`` follows tianshou: logits + (1 - mask) * (min - max - 1)."""
        q, _ = self.model(obs)
        logits = q
        if mask is not None:
            m = torch.as_tensor(np.asarray(mask), device=q.device, dtype=q.dtype)
            logits = q + (1 - m) * (q.min() - q.max() - 1.0)
        return {"logits": logits, "act": logits.argmax(dim=1).cpu().numpy()}

    def exploration_noise(self, act: np.ndarray, mask=None, rng: np.random.Generator | None = None) -> np.ndarray:
        """tianshou ``exploration_noise`` on host arrays (for callers that keep actions on the host)."""
        if np.isclose(self.eps, 0.0):
            return act
        rng = rng or np.random.default_rng(self.seed + self._calls)
        self._calls += 1
        bsz = len(act)
        rand_mask = rng.random(bsz) < self.eps
        q = rng.random((bsz, 2))
        if mask is not None:
            q = q + np.asarray(mask)
        act = act.copy()
        act[rand_mask] = q.argmax(axis=1)[rand_mask]
        return act


class MultiAgentSharedPolicy:
    """Parameter sharing manager (shared_policy.py:14-31): every agent uses the same ``policy``.  The
    reference loops over agent ids and calls the network once per id group; the rows are independent,
    so one batched call gives the same actions."""

    def __init__(self, policy: DQNPolicy, agents):
        self.policy = policy
        self.agents = list(agents)

    def forward(self, obs, mask=None):
        return self.policy.forward(obs, mask)

    __call__ = forward

    def exploration_noise(self, act, mask=None, rng=None):
        return self.policy.exploration_noise(act, mask, rng)


@dataclass
class SequenceSummaryStats:
    mean: float
    std: float
    max: float
    min: float

    @classmethod
    def from_sequence(cls, seq):
        a = np.asarray(seq, dtype=np.float64)
        return cls(float(a.mean()), float(a.std()), float(a.max()), float(a.min()))


@dataclass
class CollectStatsWithInfo:
    """Field for field the reference's collect statistics (collectors/collector.py:34, multi_agent_collector.py:339-353)."""
    n_collected_episodes: int = 0
    n_collected_steps: int = 0
    collect_time: float = 0.0
    collect_speed: float = 0.0
    returns: np.ndarray = field(default_factory=lambda: np.array([]))
    returns_stat: SequenceSummaryStats | None = None
    lens: np.ndarray = field(default_factory=lambda: np.array([], dtype=int))
    lens_stat: SequenceSummaryStats | None = None
    info: dict = field(default_factory=dict)


LOGGER_KEYS = ["total_messages_transmitted", "coverage", "messages_sent", "messages_received", "n_neighbours",
               "interested_agents", "coverage_interested_fraction", "coverage_interested_count",
               "uninterested_with_message", "episode_rewards_sum"]


class BatchedCollector:
    """``collect(n_step=... | n_episode=...)`` over a :class:`BatchedGraphEnv` (which must have been built with
    ``want_info=True``).  One iteration = forward + action selection + one environment round for every
    episode; finished episodes are restarted inside the step kernel from ``tuples``."""

    def __init__(self, policy: MultiAgentSharedPolicy | DQNPolicy, env: BatchedGraphEnv, tuples: ResetTuplesDevice,
                 exploration_noise: bool = False):
        if env.info_buf is None:
            raise ValueError("BatchedCollector needs BatchedGraphEnv(..., want_info=True)")
        self.policy = policy.policy if isinstance(policy, MultiAgentSharedPolicy) else policy
        self.env, self.tuples = env, tuples
        self.exploration_noise = exploration_noise
        self.collect_step = self.collect_episode = 0
        self.collect_time = 0.0
        self._round = 0
        B, N = env.B, env.N
        self.q = torch.zeros(B, N, 2, dtype=torch.float32, device=env.device)
        self.act = torch.full((B, N), -1, dtype=torch.int8, device=env.device)
        self.feature_errors = torch.zeros(1, dtype=torch.int32, device=env.device)
        self.reset()

    def reset(self):
        if self.tuples.count < self.env.B:
            raise ValueError(f"need at least {self.env.B} reset tuples (one per episode), got {self.tuples.count}")
        first = ResetTuplesDevice.__new__(ResetTuplesDevice)
        first.count = self.env.B
        for k in ("graph_index", "source", "interested", "scripted"):
            setattr(first, k, getattr(self.tuples, k)[: self.env.B])
        self.env.reset(first)
        self.env.set_recycling(self.tuples)
        self.env.transitions.zero_()

    def iterate(self, eps: float = 0.0, random: bool = False):
        """One collector iteration: policy forward + action selection + environment round, all on the device."""
        env = self.env
        if random:
            self.act.copy_(torch.randint(0, 2, self.act.shape, device=env.device, dtype=torch.int8))
        else:
            self.policy.model.forward_graphs(env.obs, env.active, eps=eps, philox_seed=self.policy.seed,
                                             philox_offset=self._round, q_out=self.q, act_out=self.act,
                                             discrete_features=True, feature_errors=self.feature_errors)
        env.step_device(self.act)
        self._round += 1

    def collect(self, n_step: int | None = None, n_episode: int | None = None, random: bool = False) -> CollectStatsWithInfo:
        if (n_step is None) == (n_episode is None):
            raise TypeError("Please specify exactly one of n_step or n_episode")
        env = self.env
        start = time.time()
        t0 = int(env.transitions.item())
        steps = episodes = 0
        returns, lens = [], []
        stats = {k: [] for k in LOGGER_KEYS}
        eps = self.policy.eps if self.exploration_noise else 0.0
        while True:
            self.iterate(eps, random)
            done = env.done.cpu().numpy().astype(bool)
            if done.any():
                inf = env.last_info()                    # describes the round that just ended (before the restart)
                ids = np.flatnonzero(done)
                episodes += len(ids)
                returns.extend(inf["episode_rewards_sum"][ids].tolist())
                lens.extend(inf["num_moves"][ids].tolist())
                for k in LOGGER_KEYS:
                    stats[k].extend(np.asarray(inf[k])[ids].tolist())
            steps = int(env.transitions.item()) - t0
            if (n_step and steps >= n_step) or (n_episode and episodes >= n_episode):
                break
        dt = max(time.time() - start, 1e-9)
        self.collect_step += steps
        self.collect_episode += episodes
        self.collect_time += dt
        return CollectStatsWithInfo(
            n_collected_episodes=episodes, n_collected_steps=steps, collect_time=dt, collect_speed=steps / dt,
            returns=np.array(returns), returns_stat=SequenceSummaryStats.from_sequence(returns) if returns else None,
            lens=np.array(lens, dtype=int), lens_stat=SequenceSummaryStats.from_sequence(lens) if lens else None,
            info={k: SequenceSummaryStats.from_sequence(v) for k, v in stats.items() if v})
