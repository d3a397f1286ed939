"""Versioned entry point, like the reference's ``graph_env/graph_env_v0.py``:
``from melissa_b200.graph_env_v0 import env, GraphEnv``."""
from .graph_env import GraphEnv, env

__all__ = ["env", "GraphEnv"]
