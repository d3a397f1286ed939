"""TEST INFRASTRUCTURE ONLY.  Import the UNMODIFIED reference env under stub modules.

The reference's ``graph_env/env/graph.py`` imports gymnasium, pettingzoo and
matplotlib, none of which is installed (or installable offline) in this image.  The
stubs below provide exactly the few names the reference touches; the bodies of the
pettingzoo helpers restate pettingzoo's ``AECEnv``.

Only usable where ``/root/reference`` exists (the builder container).  The GPU box
never has it: tests that need it skip there and rely on ``tests/golden``.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np

def _find_reference_root() -> str:
    """/root/reference in the builder container; on a GPU box the driver may have left an installed copy of the
    reference under baseline/_ref (git-ignored, travels with the snapshot) -- use it when it holds the env sources."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cands = [os.environ.get("MELISSA_REFERENCE_ROOT"), "/root/reference", os.path.join(here, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "graph_env", "env", "graph.py")):
            return c
    return cands[1]


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "graph_env", "env", "graph.py"))


def _install_stubs() -> None:
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")
        spaces = types.ModuleType("gymnasium.spaces")

        class _Space:
            def __init__(self, *a, **k):
                self.args, self.kwargs = a, k
                self.shape = k.get("shape")
                self.dtype = k.get("dtype")

        class Box(_Space):
            pass

        class Discrete(_Space):
            def __init__(self, n, *a, **k):
                super().__init__(n, *a, **k)
                self.n = n

        class Dict(_Space):
            def __init__(self, d=None, **k):
                super().__init__(d, **k)
                self.spaces = d

        class MultiDiscrete(_Space):
            pass

        spaces.Box, spaces.Discrete, spaces.Dict, spaces.MultiDiscrete = Box, Discrete, Dict, MultiDiscrete
        logger = types.ModuleType("gymnasium.logger")
        logger.warn = lambda *a, **k: None
        utils = types.ModuleType("gymnasium.utils")
        seeding = types.ModuleType("gymnasium.utils.seeding")

        def np_random(seed=None):
            # gymnasium.utils.seeding.np_random: Generator(PCG64(SeedSequence(seed)))
            ss = np.random.SeedSequence(seed)
            return np.random.Generator(np.random.PCG64(ss)), ss.entropy

        seeding.np_random = np_random
        utils.seeding = seeding
        gym.spaces, gym.logger, gym.utils = spaces, logger, utils
        sys.modules.update({
            "gymnasium": gym, "gymnasium.spaces": spaces, "gymnasium.logger": logger,
            "gymnasium.utils": utils, "gymnasium.utils.seeding": seeding,
        })
    if "pettingzoo" not in sys.modules:
        pz = types.ModuleType("pettingzoo")

        class AECEnv:
            """Restates the pettingzoo.AECEnv helpers GraphEnv relies on."""

            def __init__(self):
                pass

            def _deads_step_first(self):
                dead = [a for a in self.agents if (self.terminations[a] or self.truncations[a])]
                if dead:
                    self._skip_agent_selection = self.agent_selection
                    self.agent_selection = dead[0]
                return self.agent_selection

            def _clear_rewards(self):
                for a in self.rewards:
                    self.rewards[a] = 0

            def _accumulate_rewards(self):
                for a, r in self.rewards.items():
                    self._cumulative_rewards[a] += r

            def last(self, observe=True):
                a = self.agent_selection
                obs = self.observe(a) if observe else None
                return (obs, self._cumulative_rewards[a], self.terminations[a],
                        self.truncations[a], self.infos[a])

        pz.AECEnv = AECEnv
        pzu = types.ModuleType("pettingzoo.utils")
        wr = types.ModuleType("pettingzoo.utils.wrappers")
        wr.AssertOutOfBoundsWrapper = lambda e: e
        wr.OrderEnforcingWrapper = lambda e: e
        pzu.wrappers = wr
        pz.utils = pzu
        sys.modules.update({"pettingzoo": pz, "pettingzoo.utils": pzu, "pettingzoo.utils.wrappers": wr})
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        plt.clf = plt.pause = lambda *a, **k: None
        mpl.pyplot = plt
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})


_LOADED = {}


def load_reference():
    """Return a namespace with the reference's World, GraphEnv, heuristics (unmodified
    code objects).  Adds the one-line MPR adapter documented in SURVEY.md section 8c as a
    SEPARATE registry key ``"mpr"`` override (the shipped ``mpr_heuristic`` returns a
    bare ndarray while ``World.step`` reads ``result.relay_mask``: reference
    heuristics/mpr.py:72 vs utils/core.py:229)."""
    if _LOADED:
        return _LOADED["ns"]
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from graph_env.env.utils import core as ref_core
    from graph_env.env.utils import heuristics as ref_heur
    from graph_env.env.utils.heuristics import mpr as ref_mpr
    from graph_env.env import graph as ref_graph
    from graph_env.env.utils import selector as ref_selector

    def _mpr_adapter(agent):
        return ref_heur.HeuristicResult(relay_mask=ref_mpr.mpr_heuristic(agent), action=None)

    ref_heur.HEURISTIC_REGISTRY["mpr"] = _mpr_adapter
    # core.py did `from .heuristics import HEURISTIC_REGISTRY` (same dict object) -> patched too.
    ns = types.SimpleNamespace(
        core=ref_core, graph=ref_graph, heuristics=ref_heur, selector=ref_selector,
        World=ref_core.World, GraphEnv=ref_graph.GraphEnv, Agent=ref_core.Agent, State=ref_core.State,
        HeuristicResult=ref_heur.HeuristicResult, HEURISTIC_REGISTRY=ref_heur.HEURISTIC_REGISTRY,
        np_random=sys.modules["gymnasium.utils.seeding"].np_random,
    )
    _LOADED["ns"] = ns
    return ns


@contextlib.contextmanager
def in_dir(path):
    """The reference globs ``graph_topologies/...`` relative to the CWD (core.py:165-175)."""
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)
