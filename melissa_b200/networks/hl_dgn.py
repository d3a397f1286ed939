"""HL-DGN: MLP encoder -> GATv2Conv -> decision-maker mask -> global pool -> dueling Q/V.
Mirror of the reference ``HLDGNNetwork`` (graph_env/env/utils/networks/hl_dgn.py:14-119);
its output does not depend on the controlling index."""
from typing import Any, Dict, Optional, Tuple

from .common import DGNBase, GATv2Params, MLPParams


class HLDGNNetwork(DGNBase):
    KIND = "hl_dgn"

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int, num_heads: int, agents_num: int,
                 aggregator: str = "mean", dueling_param: Optional[Tuple[Dict[str, Any], Dict[str, Any]]] = None,
                 device: str = "cpu", edge_attributes: bool = False):
        super().__init__()
        self._init_common(input_dim, hidden_dim, output_dim, num_heads, agents_num, dueling_param, device,
                          edge_attributes)
        if aggregator not in ("mean", "add", "max"):
            raise KeyError(aggregator)                                           # hl_dgn.py:56-60
        self.aggregator_name = aggregator
        self.encoder = MLPParams(input_dim, hidden_dim, [hidden_dim])            # hl_dgn.py:41-46
        self.conv1 = GATv2Params(hidden_dim, hidden_dim, num_heads)              # hl_dgn.py:49-53
        self._build_heads(hidden_dim * num_heads, dueling_param, output_dim)

    def _conv_tensors(self, c):
        return [c.lin_l.weight, c.lin_l.bias, c.lin_r.weight, c.lin_r.bias, None, None, c.att, c.bias]
