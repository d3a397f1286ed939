"""melissa_b200 -- B200-native rollout hot path for Melissa (graph message dissemination MARL).

Host side is Python/PyTorch (device memory, streams, torch.distributed); the hot path is
hand-written sm_100a CUDA behind the C ABI in include/melissa_b200.h, loaded with ctypes
from melissa_b200/lib/libmelissa_b200.so.  There is no CPU fallback.
"""
__version__ = "0.1.0"
