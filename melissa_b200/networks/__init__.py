from .dgn_r import DGNRNetwork
from .hl_dgn import HLDGNNetwork
from .l_dgn import LDGNNetwork

NETWORKS = {"dgn_r": DGNRNetwork, "l_dgn": LDGNNetwork, "hl_dgn": HLDGNNetwork}

__all__ = ["DGNRNetwork", "HLDGNNetwork", "LDGNNetwork", "NETWORKS"]
