"""Diagnostics: the fused-epilogue GEMMs of a Flux block at full size, a few launches each (for ncu / timing).
    python tools/gemm_epi_probe.py [gelu|gate|all]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flux_fp8_api_b200 import ops  # noqa: E402

E4M3, E5M2, BF16 = torch.float8_e4m3fn, torch.float8_e5m2, torch.bfloat16
which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = "cuda"
sa, sw = torch.tensor(1 / 64., device=dev), torch.tensor(1 / 32., device=dev)


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


if which in ("gelu", "all"):
    M, N, K = 4608, 12288, 3072
    a = (torch.randn(M, K, device=dev) * 4).to(BF16).to(E5M2)
    w = torch.randn(N, K, device=dev).to(BF16).to(E4M3)
    bias = (torch.randn(N, device=dev) * 0.5).to(BF16)
    so = torch.tensor(64.0, device=dev)
    out = torch.empty(M, N, dtype=E5M2, device=dev)
    ms = timed(lambda: ops.f8_gemm_gelu_quant(a, w, bias, sa, sw, so, E5M2, out=out))
    print(f"gelu_quant {M}x{N}x{K}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TFLOP/s")
if which in ("gate", "all"):
    M, N, K = 4608, 3072, 15360
    a = (torch.randn(M, K, device=dev) * 4).to(BF16).to(E5M2)
    w = torch.randn(N, K, device=dev).to(BF16).to(E4M3)
    bias = (torch.randn(N, device=dev) * 0.5).to(BF16)
    resid = torch.randn(M, N, device=dev).to(BF16)
    gate = torch.randn(1, N, device=dev).to(BF16)
    out = torch.empty(M, N, dtype=BF16, device=dev)
    ms = timed(lambda: ops.f8_gemm_gate_residual(a, w, bias, sa, sw, resid, gate, M, out=out))
    print(f"gate_residual {M}x{N}x{K}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TFLOP/s")
if which in ("qkv", "all"):
    import math
    from oracle import flux_oracle as O
    B, L, T, H, K = 1, 4096, 512, 24, 3072
    S, D = L + T, H * 128
    N, M = 3 * D, B * L
    a = (torch.randn(M, K, device=dev) * 4).to(BF16).to(E5M2)
    w = (torch.randn(N, K, device=dev) * 0.5).to(BF16).to(E4M3)
    bias = (torch.randn(N, device=dev) * 0.5).to(BF16)
    qw = (1 + 0.05 * torch.randn(128, device=dev)).float()
    kw = (1 + 0.05 * torch.randn(128, device=dev)).float()
    ids = torch.cat((torch.zeros(1, T, 3), O.make_img_ids(1, 64, 64, torch.float32)), 1).to(BF16).to(dev)
    pe = O.embed_nd(ids, [16, 56, 56], 10000, BF16)
    cos, sin = pe[:, 0, :, :, 0, 0].contiguous(), pe[:, 0, :, :, 1, 0].contiguous()
    q = torch.zeros(B, H, S, 128, dtype=BF16, device=dev)
    k, v = torch.zeros_like(q), torch.zeros_like(q)
    ms = timed(lambda: ops.f8_gemm_qkv_rope(a, w, bias, sa, sw, q, k, v, qw, kw, cos, sin, L, T))
    print(f"qkv_rope {M}x{N}x{K}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TFLOP/s")
