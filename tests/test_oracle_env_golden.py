"""CPU: the oracle replays every golden trace recorded from the unmodified reference
(tests/golden/make_golden.py) bit-exactly, all episodes of a file as one batch."""
import numpy as np
import pytest

from golden_util import GoldenEnv, check_reset, check_round, golden_env_files
from oracle.env_oracle import BatchedEnvOracle

FILES = golden_env_files()


def test_golden_files_present():
    assert len(FILES) >= 10


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("env_")[-1][:-4])
def test_oracle_replays_golden(path):
    g = GoldenEnv(path)
    o = BatchedEnvOracle(g.E, g.N, dynamic=g.dynamic, is_testing=g.is_testing, heuristic=g.heuristic)
    kw = g.reset_args()
    o.reset(np.arange(g.E), kw["adj"], kw["pos"], kw["source"], kw["interested"], kw["scripted"],
            move_offsets=kw["move_offsets"])
    check_reset(g, o.obs(), o.active, o.has_message, o.msgs)
    np.testing.assert_array_equal(o.received_from.sum(2), g.z["reset_recv_count"])
    for r in range(g.max_rounds):
        alive, actions, mo = g.round_inputs(r)
        obs, rew, active, term, done = o.step(actions, move_offsets=mo)
        check_round(g, r, alive, obs, rew, active, term, done, o.episode_rewards_sum, o.world_msgs)
        np.testing.assert_array_equal(o.adj[alive], g.adj_expected(r)[alive])
        np.testing.assert_array_equal(o.received_from.sum(2)[alive], g.round_expected(r, "recv_count")[alive])
        inf = o.info()
        st = g.round_expected(r, "stats")[alive]
        for k, key in enumerate(["total_messages_transmitted", "messages_sent", "messages_received", "n_neighbours",
                                 "interested_agents", "coverage_interested_count", "uninterested_with_message"]):
            np.testing.assert_array_equal(inf[key][alive], st[:, k], err_msg=key)
        np.testing.assert_array_equal(inf["covered"][alive] / g.N, g.round_expected(r, "stats_coverage")[alive])
