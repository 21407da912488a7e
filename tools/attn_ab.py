"""A/B of attention kernel variants (diagnostics): python tools/attn_ab.py [variant ...]   (default: 0 15)
Checks every variant against variant 0 bit for bit (and variant 0 against torch SDPA), then times each isolated and
back to back at the model's sequence lengths."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flux_fp8_api_b200 import ops  # noqa: E402

BF16 = torch.bfloat16
variants = [int(a) for a in sys.argv[1:]] or [17, 16]
SIZES = [int(x) for x in os.environ.get("ATTN_AB_SIZES", "4608,1000,9728,2816").split(",")]


def timed(fn, iters, warm=3):
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


for S in SIZES:
    B, H = 1, 24
    g = torch.Generator(device="cuda").manual_seed(S)
    q = (torch.randn(B, H, S, 128, device="cuda", generator=g) * 1.5).to(BF16)
    k = (torch.randn(B, H, S, 128, device="cuda", generator=g) * 1.5).to(BF16)
    v = torch.randn(B, H, S, 128, device="cuda", generator=g).to(BF16)
    base = ops.attention(q, k, v, variant=17)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, S, H * 128)
    print(f"S={S}: variant 0 vs torch SDPA max|d| = {(base.float() - ref.float()).abs().max().item():.4g}")
    flops = 4.0 * B * H * S * S * 128
    for var in variants:
        out = ops.attention(q, k, v, variant=var)
        torch.cuda.synchronize()
        same = torch.equal(out, base)
        d = (out.float() - base.float()).abs().max().item()
        iso = timed(lambda: ops.attention(q, k, v, variant=var), 10)
        b2b = timed(lambda: ops.attention(q, k, v, variant=var), max(20, int(600.0 / iso)))
        print(f"   variant {var:2d}: bit-identical to 0: {same} (max|d| {d:.3g})   isolated {iso * 1e3:7.1f} us "
              f"{flops / iso / 1e9:6.0f} TF/s | back-to-back {b2b * 1e3:7.1f} us {flops / b2b / 1e9:6.0f} TF/s", flush=True)
