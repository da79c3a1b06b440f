"""esrecsys_b200: B200-native embedding-training hot path for the ESRecsys trainers.

Host side is Python/PyTorch; every kernel lives in ``csrc/`` behind the C ABI
declared in ``include/esr.h`` (loaded with ctypes by ``esrecsys_b200._lib``).
There is no CPU fallback: anything that computes raises if ``libesr.so`` or a
CUDA device is missing.
"""
__version__ = "0.1.0"
