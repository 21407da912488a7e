"""Event timeline of CTA (0,0,0) of the step-interleaved attention kernel (diagnostics; needs tools/build_probe.sh):
FLUXB200_LIB=$PWD/tools/ab/libflux_probe.so python tools/attn_trace.py [variant]"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flux_fp8_api_b200 import ops, _cabi as cabi
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 18
B, H, S = 1, 24, 4608
q = torch.randn(B, H, S, 128, device="cuda").to(torch.bfloat16); k = torch.randn_like(q); v = torch.randn_like(q)
for _ in range(3):
    ops.attention(q, k, v, variant=variant)
buf = (C.c_ulonglong * 576)()
lib = C.CDLL(os.environ["FLUXB200_LIB"])
assert lib.fluxb200_debug_trace(buf) == 0
t = [[[buf[(r * 24 + j) * 8 + e] for e in range(8)] for j in range(24)] for r in range(3)]
t0 = min(x for r in t for j in r for x in j if x)
names = {0: "issuers: QK[top s_free k_full issued] PV[p_lo PVlo p_ready PVhi+commit]",
         1: "WG0:    top s_ready ld max m_wait o_wait P_lo P_hi", 2: "WG1:    top s_ready ld max m_wait o_wait P_lo P_hi"}
for r in range(3):
    print(names[r])
    for j in range(24):
        if any(t[r][j]):
            print(f"  j={j:2d} " + " ".join(f"{(x - t0) if x else -1:7d}" for x in t[r][j]))
