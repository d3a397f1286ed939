"""Generates tests/golden/aec_*.npz: AEC-level traces of the UNMODIFIED reference GraphEnv
(every `last()` tuple, every `step(action)`), recorded in the builder container.

    python tests/golden/make_golden_aec.py

Per file (one configuration, E episodes, S AEC steps in total, flattened):
  adj0/pos0/seed per episode; step_ptr[E+1];
  per AEC step: agent (index selected before the step), action (-1 = None for a dead agent),
  obs f32 [S, 8N+1], mask i8 [S,2], cum_reward f64, terminated/truncated, env_step, environment_step,
  explicit_reset, active_one_hop [S,N], n_agents (len(env.agents)).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from melissa_b200 import topology  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402


def record(name, N, side, graph_seeds, seeds, *, dynamic=False, heuristic=None, ratio=0.0):
    ref = load_reference()
    eps = []
    for gs in graph_seeds:
        g = topology.make_connected_graph(N, gs, side)
        env = ref.GraphEnv(graph=g, number_of_agents=N, radius=0.2, dynamic_graph=dynamic, heuristic=heuristic,
                           scripted_agents_ratio=ratio)
        for s in seeds:
            seed = 100 * gs + s
            adj0, pos0 = topology.graph_to_arrays(env.world.graph, N)
            env.reset(seed=seed)
            rng = np.random.default_rng(seed)
            steps = []
            guard = 0
            while env.agents and env.agent_selection is not False:
                guard += 1
                assert guard < 5000
                agent = env.agent_selection
                o, cum, term, trunc, info = env.last()
                a = -1 if (term or trunc) else int(rng.integers(0, 2))
                steps.append(dict(
                    agent=int(agent), action=a, obs=o["observation"].copy(), mask=o["action_mask"].copy(), cum=float(cum),
                    term=bool(term), trunc=bool(trunc), env_step=int(info["env_step"]),
                    environment_step=bool(info["environment_step"]), explicit_reset=bool(info["explicit_reset"]),
                    aoh=np.asarray(info["active_one_hop_neighbors"], dtype=bool).copy(), n_agents=len(env.agents)))
                env.step(None if a < 0 else a)
            eps.append((adj0, pos0, seed, steps))
    S = [st for *_, sts in eps for st in sts]
    out = dict(
        n_nodes=np.int32(N), dynamic=np.bool_(dynamic), heuristic=np.str_(heuristic or ""), scripted_ratio=np.float64(ratio),
        adj0=np.stack([e[0] for e in eps]), pos0=np.stack([e[1] for e in eps]), seed=np.array([e[2] for e in eps]),
        step_ptr=np.concatenate([[0], np.cumsum([len(e[3]) for e in eps])]).astype(np.int32),
        agent=np.array([s["agent"] for s in S], dtype=np.int32), action=np.array([s["action"] for s in S], dtype=np.int8),
        obs=np.stack([s["obs"] for s in S]), mask=np.stack([s["mask"] for s in S]),
        cum=np.array([s["cum"] for s in S]), term=np.array([s["term"] for s in S]), trunc=np.array([s["trunc"] for s in S]),
        env_step=np.array([s["env_step"] for s in S], dtype=np.int32),
        environment_step=np.array([s["environment_step"] for s in S]), explicit_reset=np.array([s["explicit_reset"] for s in S]),
        aoh=np.stack([s["aoh"] for s in S]), n_agents=np.array([s["n_agents"] for s in S], dtype=np.int32))
    path = os.path.join(HERE, f"aec_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {len(eps)} episodes, {len(S)} AEC steps -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    record("static_n20", 20, 0.6, range(3), range(3))
    record("static_n12_mixed_bcast", 12, 0.45, range(3, 5), range(2), heuristic="simple_broadcast", ratio=0.4)
    record("dynamic_n20", 20, 0.6, range(5, 7), range(2), dynamic=True)
