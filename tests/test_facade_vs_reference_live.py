"""CPU, builder container only (needs /root/reference): the AEC facade's host logic (with the oracle round stepper
behind it) against the UNMODIFIED reference GraphEnv driven live with the same seeds and the same action sequence, for the
constructor options whose randomness lives in GLOBAL streams -- ``random_graph=True`` (networkx draws from Python's
``random``) and the probabilistic heuristics (numpy's global stream).  Both environments are constructed and stepped
after re-seeding those streams identically."""
import random

import numpy as np
import pytest

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")


def _drive(make_env, seed, n_episodes, n_steps):
    """-> list of (agent_selection, obs bits, cumulative reward bits, terminated) at every AEC step."""
    random.seed(seed)
    np.random.seed(seed)
    env = make_env()
    # the constructor's two UNSEEDED resets consume a run-dependent number of global draws (whether the source falls into
    # the scripted set depends on OS entropy, in the reference too): re-seed once the environment exists
    random.seed(seed + 100)
    np.random.seed(seed + 100)
    rng = np.random.default_rng(seed + 1)            # action stream (a private generator: not the global one)
    trace = []
    for ep in range(n_episodes):
        env.reset(seed=seed + ep)
        for _ in range(n_steps):
            if not env.agents or not env.agent_selection:
                break
            a = env.agent_selection
            obs, cum, term, trunc, info = env.last()
            trace.append((str(a), np.asarray(obs["observation"], dtype=np.float32).view(np.uint32).copy(),
                          np.float64(cum).view(np.uint64), bool(term), bool(trunc)))
            env.step(None if (term or trunc) else int(rng.integers(0, 2)))
    return trace


def _compare(kwargs, seed=7, n_episodes=4, n_steps=150):
    from aec_util import OracleRoundStepper
    from melissa_b200.graph_env import GraphEnv
    ref = ref_loader.load_reference()
    N = kwargs["number_of_agents"]
    want = _drive(lambda: ref.GraphEnv(**kwargs), seed, n_episodes, n_steps)
    got = _drive(lambda: GraphEnv(stepper=OracleRoundStepper(N, kwargs.get("dynamic_graph", False), False, kwargs.get("heuristic")),
                                  **kwargs), seed, n_episodes, n_steps)
    assert len(want) == len(got) and len(want) > 40
    for k, (w, g) in enumerate(zip(want, got)):
        assert w[0] == g[0], f"step {k}: agent {w[0]} vs {g[0]}"
        assert np.array_equal(w[1], g[1]), f"step {k}: observation"
        assert w[2] == g[2] and w[3:] == g[3:], f"step {k}: reward / flags"


def test_random_graph_draws_the_same_topologies_as_the_reference():
    _compare(dict(number_of_agents=12, radius=0.45, random_graph=True))


@pytest.mark.parametrize("heuristic,prob", [("probabilistic_gossip", 0.5), ("probabilistic_relay", 0.6)])
def test_probabilistic_heuristics_consume_the_global_numpy_stream_like_the_reference(heuristic, prob):
    import networkx as nx
    g = nx.random_geometric_graph(14, 0.45, seed=3)
    while not nx.is_connected(g):
        g = nx.random_geometric_graph(14, 0.5, seed=4)
    for n, d in g.nodes(data=True):
        d["pos"] = list(d["pos"])
    _compare(dict(graph=g, number_of_agents=14, radius=0.45, scripted_agents_ratio=0.4, heuristic=heuristic,
                  heuristic_params={"prob": prob}))


def test_probabilistic_heuristic_without_prob_fails_like_the_reference():
    from melissa_b200.graph_env import GraphEnv
    with pytest.raises(TypeError):
        GraphEnv(number_of_agents=8, scripted_agents_ratio=0.5, heuristic="probabilistic_gossip", random_graph=True, radius=0.9)
