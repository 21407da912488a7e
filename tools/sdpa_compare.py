"""Our attention kernel against torch's fused SDPA backends on the same box (the reference calls
F.scaled_dot_product_attention, modules/flux_model.py:43): python tools/sdpa_compare.py [S ...]

Times the launch the reference makes -- SDPA on [B,H,S,128] bf16 followed by transpose(1,2).reshape to [B,S,H*128]
(materialised, as the next linear needs it contiguous) -- against fluxb200_attention (which writes that layout
directly), isolated (GPU otherwise idle, clocks high) and back to back for ~1 s (power-capped regime of the step)."""
import json
import os
import sys
import time

import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flux_fp8_api_b200 import ops  # noqa: E402

BF16 = torch.bfloat16


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


rows = []
for S in [int(a) for a in sys.argv[1:]] or [4608, 4352, 2816, 9728]:
    B, H = 1, 24
    g = torch.Generator(device="cuda").manual_seed(S)
    q = torch.randn(B, H, S, 128, device="cuda", generator=g).to(BF16)
    k = torch.randn(B, H, S, 128, device="cuda", generator=g).to(BF16)
    v = torch.randn(B, H, S, 128, device="cuda", generator=g).to(BF16)
    out = torch.empty(B, S, H * 128, dtype=BF16, device="cuda")
    flops = 4.0 * B * H * S * S * 128
    cands = {"ours (fluxb200_attention)": lambda: ops.attention(q, k, v, out=out)}
    for name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION),
                     ("efficient", SDPBackend.EFFICIENT_ATTENTION), ("default", None)):
        def run(be=be):
            if be is None:
                o = F.scaled_dot_product_attention(q, k, v)
            else:
                with sdpa_kernel([be]):
                    o = F.scaled_dot_product_attention(q, k, v)
            return o.transpose(1, 2).reshape(B, S, H * 128)
        try:
            run()
            cands[f"torch SDPA {name}"] = run
        except RuntimeError as ex:
            rows.append({"S": S, "impl": f"torch SDPA {name}", "error": str(ex)[:80]})
    for name, fn in cands.items():
        time.sleep(0.5)
        iso = timed(fn, 10)
        sustained = timed(fn, max(20, int(1000.0 / iso)))
        rows.append({"S": S, "impl": name, "isolated_us": iso * 1e3, "isolated_tflops": flops / iso / 1e9,
                     "sustained_us": sustained * 1e3, "sustained_tflops": flops / sustained / 1e9})
        print(f"S={S:5d} {name:28s} isolated {iso * 1e3:8.1f} us {flops / iso / 1e9:7.0f} TFLOP/s | back-to-back "
              f"{sustained * 1e3:8.1f} us {flops / sustained / 1e9:7.0f} TFLOP/s", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/sdpa_compare.json", "w"), indent=1)
