"""Replay helper for tests/golden/env_*.npz (see tests/golden/make_golden.py).

All E episodes of a file are replayed as ONE batch (B = E).  Episodes that ended earlier
than the longest one keep receiving "no action" rounds and are simply no longer compared.
"""
from __future__ import annotations

import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_env_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "env_*.npz")))


class GoldenEnv:
    def __init__(self, path):
        z = np.load(path)
        self.name = os.path.basename(path)[4:-4]
        self.z = z
        self.N = int(z["n_nodes"])
        self.dynamic = bool(z["dynamic"])
        self.is_testing = bool(z["is_testing"])
        self.heuristic = str(z["heuristic"]) or None
        self.E = z["adj0"].shape[0]
        self.ptr = z["round_ptr"]
        self.lens = np.diff(self.ptr)
        self.max_rounds = int(self.lens.max())

    def reset_args(self):
        z = self.z
        kw = dict(adj=z["adj0"], pos=z["pos0"], source=z["source"], interested=z["interested"],
                  scripted=z["scripted"])
        kw["move_offsets"] = z["reset_move_offsets"] if self.dynamic else None
        return kw

    def round_inputs(self, r):
        """-> (alive bool[E], actions i8[E,N], move_offsets f64[E,2,N] | None)"""
        alive = self.lens > r
        idx = self.ptr[:-1] + r
        idx = np.where(alive, idx, 0)
        actions = np.where(alive[:, None], self.z["actions"][idx], -1).astype(np.int8)
        mo = None
        if self.dynamic:
            mo = np.where(alive[:, None, None], self.z["move_offsets"][idx], 0.0)
        return alive, actions, mo

    def round_expected(self, r, key):
        alive = self.lens > r
        idx = np.where(alive, self.ptr[:-1] + r, 0)
        return self.z[key][idx]

    def adj_expected(self, r):
        a = np.unpackbits(self.round_expected(r, "adj"), axis=-1)[..., : self.N].astype(bool)
        return a


def check_reset(g: GoldenEnv, obs, active, has_message=None, msgs=None):
    z = g.z
    np.testing.assert_array_equal(np.asarray(obs).view(np.uint32), z["reset_obs"].view(np.uint32),
                                  err_msg=f"{g.name}: reset obs")
    np.testing.assert_array_equal(np.asarray(active).astype(bool), z["reset_active"], err_msg=f"{g.name}: reset active")
    if has_message is not None:
        np.testing.assert_array_equal(np.asarray(has_message).astype(bool), z["reset_has_message"])
    if msgs is not None:
        np.testing.assert_array_equal(np.asarray(msgs), z["reset_msgs"])


def check_round(g: GoldenEnv, r, alive, obs, reward, active, terminated, done=None, rewards_sum=None,
                world_msgs=None):
    ctx = f"{g.name}: round {r}"
    e = lambda k: g.round_expected(r, k)[alive]
    np.testing.assert_array_equal(np.asarray(obs)[alive].view(np.uint32), e("obs").view(np.uint32), err_msg=ctx + " obs")
    np.testing.assert_array_equal(np.asarray(reward)[alive].view(np.uint64), e("reward").view(np.uint64),
                                  err_msg=ctx + " reward")
    np.testing.assert_array_equal(np.asarray(active)[alive].astype(bool), e("active"), err_msg=ctx + " active")
    np.testing.assert_array_equal(np.asarray(terminated)[alive].astype(bool), e("terminated"), err_msg=ctx + " terminated")
    if done is not None:
        np.testing.assert_array_equal(np.asarray(done)[alive].astype(bool), ~e("active").any(axis=1), err_msg=ctx + " done")
    if rewards_sum is not None:
        np.testing.assert_array_equal(np.asarray(rewards_sum)[alive].view(np.uint64), e("rewards_sum").view(np.uint64),
                                      err_msg=ctx + " rewards_sum")
    if world_msgs is not None:
        np.testing.assert_array_equal(np.asarray(world_msgs)[alive], e("world_msgs"), err_msg=ctx + " world_msgs")
