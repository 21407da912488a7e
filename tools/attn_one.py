"""One attention launch per listed variant at S=4608 (for ncu): python tools/attn_one.py 18"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flux_fp8_api_b200 import ops
S = int(os.environ.get("S", "4608"))
q = torch.randn(1, 24, S, 128, device="cuda").to(torch.bfloat16); k = torch.randn_like(q); v = torch.randn_like(q)
for var in [int(a) for a in sys.argv[1:]] or [0]:
    for _ in range(3):
        ops.attention(q, k, v, variant=var)
torch.cuda.synchronize()
