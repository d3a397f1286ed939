"""DGN-R: MLP encoder -> TransformerConv(root_weight=False) x2 -> three snapshots -> dueling Q/V.
Mirror of the reference ``DGNRNetwork`` (graph_env/env/utils/networks/dgn_r.py:13-129)."""
from typing import Any, Dict, Optional, Tuple

from .common import DGNBase, MLPParams, TransformerConvParams


class DGNRNetwork(DGNBase):
    KIND = "dgn_r"

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int, num_heads: int, agents_num: int,
                 dueling_param: Optional[Tuple[Dict[str, Any], Dict[str, Any]]] = None, device: str = "cpu",
                 edge_attributes: bool = False):
        super().__init__()
        self._init_common(input_dim, hidden_dim, output_dim, num_heads, agents_num, dueling_param, device,
                          edge_attributes)
        self.final_latent_dim = hidden_dim + hidden_dim * num_heads * 2                    # dgn_r.py:63
        self.encoder = MLPParams(input_dim, hidden_dim, [hidden_dim])                      # dgn_r.py:39-44
        self.conv1 = TransformerConvParams(hidden_dim, hidden_dim, num_heads)              # dgn_r.py:47-52
        self.conv2 = TransformerConvParams(hidden_dim * num_heads, hidden_dim, num_heads)  # dgn_r.py:53-58
        self._build_heads(self.final_latent_dim, dueling_param, output_dim)

    def _conv_tensors(self, c):
        # order q | k | v (the kernel's projection row layout); lin_skip is unused (root_weight=False)
        return [c.lin_query.weight, c.lin_query.bias, c.lin_key.weight, c.lin_key.bias,
                c.lin_value.weight, c.lin_value.bias, None, None]
