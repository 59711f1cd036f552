"""kzero_b200 -- B200-native (sm_100a) batched AlphaZero-ResNet evaluation for kZero's self-play hot path.

The product is the C-ABI library `libkzb200.so` (include/kzb200.h, sources in kzero_b200/csrc);
this package is the thin Python host mirror of the reference's `Network` interface over that ABI.
There is no CPU fallback: anything that evaluates a network raises if the CUDA library is missing.
"""
__all__ = ["netgen", "mapping"]
