"""boa_b200 - B200-native drop-in for the segmentation + body-composition hot path of UMEssen/Body-and-Organ-Analysis.

Host code is thin Python over the C ABI of `libboa_b200.so` (include/boa_b200.h); PyTorch tensors are containers for
device memory only.  There is no CPU implementation of any compute entry: without the CUDA library or a CUDA device
the calls raise.
"""
__version__ = "0.1.0"
