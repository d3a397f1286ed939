"""L-DGN: MLP encoder -> GATv2Conv x2 -> three controlling-node snapshots -> dueling Q/V.
Same constructor, attribute names and state_dict keys as the reference
``LDGNNetwork`` (graph_env/env/utils/networks/l_dgn.py:12-151); forward runs in CUDA."""
from typing import Any, Dict, Optional, Tuple

from .common import DGNBase, GATv2Params, MLPParams


class LDGNNetwork(DGNBase):
    KIND = "l_dgn"

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int, num_heads: int, agents_num: int,
                 dueling_param: Optional[Tuple[Dict[str, Any], Dict[str, Any]]] = None, device: str = "cpu",
                 edge_attributes=False):
        super().__init__()
        self._init_common(input_dim, hidden_dim, output_dim, num_heads, agents_num, dueling_param, device,
                          edge_attributes)
        self.final_latent_dim = hidden_dim + hidden_dim * num_heads * 2          # l_dgn.py:44
        self.encoder = MLPParams(input_dim, hidden_dim, [hidden_dim])            # l_dgn.py:49-54
        self.conv1 = GATv2Params(hidden_dim, hidden_dim, num_heads)              # l_dgn.py:56-60
        self.conv2 = GATv2Params(hidden_dim * num_heads, hidden_dim, num_heads)  # l_dgn.py:61-65
        self._build_heads(self.final_latent_dim, dueling_param, output_dim)

    def _conv_tensors(self, c):
        return [c.lin_l.weight, c.lin_l.bias, c.lin_r.weight, c.lin_r.bias, None, None, c.att, c.bias]

    def _state_out(self, state):
        return None                                                              # l_dgn.py:151
