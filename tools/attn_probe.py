"""Attention phase-timing probe (diagnostics): python tools/attn_probe.py [variant ...]"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flux_fp8_api_b200 import ops, _cabi as cabi
BF16 = torch.bfloat16
def timed(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
B, H, S = 1, 24, 4608
q = torch.randn(B, H, S, 128, device="cuda").to(BF16); k = torch.randn_like(q); v = torch.randn_like(q)
out = torch.empty(B, S, H * 128, dtype=BF16, device="cuda")
for variant in [int(a) for a in sys.argv[1:]] or [0, 1]:
    ms = timed(lambda: ops.attention(q, k, v, out=out, variant=variant))
    buf = (C.c_ulonglong * 16)()
    cabi.load().fluxb200_debug_counters(buf)
    c = list(buf)
    n = max(c[6], 1)
    print(f"variant {variant}: {ms*1e3:.1f} us {4*B*H*S*S*128/ms/1e9:.0f} TFLOP/s")
    print(f"   softmax warp per step (cycles): wait_S={c[0]/n:.0f} tmem_ld={c[1]/n:.0f} max={c[2]/n:.0f} exp={c[3]/n:.0f} wait_O={c[4]/n:.0f} store_P={c[5]/n:.0f}  steps={n}  total={sum(c[:6])/n:.0f}")
    print(f"   MMA issuer per step (cycles): wait_P={c[8]/n:.0f} wait_KV={c[9]/n:.0f} issue={c[10]/n:.0f}")
