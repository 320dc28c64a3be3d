"""interactvlm_b200: B200-native (sm_100a) implementation of InteractVLM's inference hot path.

Host code is Python/PyTorch (memory, streams, torch.distributed); all computation runs in hand-written CUDA
kernels behind the C ABI of include/ivlm_b200.h (interactvlm_b200/csrc, loaded with ctypes).  There is no CPU
or eager-PyTorch fallback: without the built library and a CUDA device the product path raises.
"""
__version__ = "0.1.0"
