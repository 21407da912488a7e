"""GEMM timing probe (diagnostics): python tools/gemm_probe.py ; honours FLUXB200_GEMM_CG / FLUXB200_GEMM_DEBUG."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flux_fp8_api_b200 import ops, _cabi as cabi
E4M3, E5M2, BF16 = torch.float8_e4m3fn, torch.float8_e5m2, torch.bfloat16
def timed(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
print("CG", os.environ.get("FLUXB200_GEMM_CG"), "DEBUG", os.environ.get("FLUXB200_GEMM_DEBUG"))
for (M, N, K) in [(4608, 3072, 3072), (4608, 21504, 3072), (4608, 3072, 15360), (4096, 12288, 3072), (16384, 8192, 8192)]:
    a = (torch.randn(M, K, device="cuda") * 4).to(BF16).to(E5M2)
    w = torch.randn(N, K, device="cuda").to(BF16).to(E4M3)
    sa = torch.tensor(1 / 64., device="cuda"); sw = torch.tensor(1 / 32., device="cuda")
    out = torch.empty(M, N, dtype=BF16, device="cuda")
    g = ops.gemm_args(a, w, None, sa, sw, cabi.EPI_PLAIN); g.out, g.ldo = out.data_ptr(), N
    ms = timed(lambda: ops.run_gemm(g))
    print(f"  M{M} N{N} K{K}: {ms*1e3:8.1f} us {2*M*N*K/ms/1e9:8.1f} TFLOP/s")
