"""cmflow_b200 -- B200-native (sm_100a) implementation of CMFlow's scene-flow inference hot path.

Host side is Python/PyTorch (device memory, streams, torch.distributed); all compute is hand-written
CUDA behind the C ABI declared in include/cmflow_b200.h.  There is no CPU fallback: importing an
operator without the built library raises.
"""
__version__ = "0.1.0"
