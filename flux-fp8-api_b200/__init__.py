"""flux-fp8-api_b200: a B200 (sm_100a) native FP8 Flux-DiT denoise hot path.

Python host code mirroring the reference's module surface for this path (F8Linear, Modulation,
DoubleStreamBlock, SingleStreamBlock, attention, rope, QKNorm, Flux) on top of hand-written CUDA
kernels reached through the C ABI in include/flux_b200.h (libflux_b200.so).
"""
__version__ = "0.1.0"
