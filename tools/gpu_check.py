"""Stand-alone GPU bring-up / diagnostics script (run on the B200 box through gpurun).

    python tools/gpu_check.py            # every group, each in its own subprocess with a timeout
    python tools/gpu_check.py gemm       # one group in-process

Each group compares the CUDA kernels (through the C ABI) with the oracle evaluated on the same
tensors and prints max-abs / mismatch statistics plus a CUDA-event timing.  The pytest suite
(tests/test_gpu_*.py) asserts the same things; this script exists to get as much diagnostic text as
possible out of a single GPU session.
"""
from __future__ import annotations

import math
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BF16 = torch.bfloat16
E4M3, E5M2 = torch.float8_e4m3fn, torch.float8_e5m2
DEV = "cuda"
FAILS = []


def timed(fn, iters=10, warm=3):
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


def cmp(name, got, ref, tol, frac_tol=1.0, exact=False):
    g, r = got.float(), ref.float()
    bad_nan = torch.isnan(g).sum().item()
    d = (g - r).abs()
    mx = d.max().item()
    frac = (d > 0).float().mean().item()
    ok = (mx <= tol and frac <= frac_tol and bad_nan == 0) if not exact else (mx == 0 and bad_nan == 0)
    print(f"  [{'PASS' if ok else 'FAIL'}] {name:58s} max|d|={mx:.4e} mismatch={frac*100:.4f}% nan={bad_nan} "
          f"ref_amax={r.abs().max().item():.3f}")
    if not ok:
        FAILS.append(name)
    return ok


def block_report(got, ref, bs=32):
    d = (got.float() - ref.float()).abs()
    M, N = d.shape
    Mb, Nb = min(M // bs, 8), min(N // bs, 16)
    print("    per-32x32 block max|diff| (rows x cols):")
    for i in range(Mb):
        print("     ", " ".join(f"{d[i*bs:(i+1)*bs, j*bs:(j+1)*bs].max().item():8.2e}" for j in range(Nb)))


def rand_fp8(shape, dtype, std=1.0, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(shape, device=DEV, generator=g) * std).to(BF16).to(dtype)


def scalar(v):
    return torch.tensor(v, dtype=torch.float32, device=DEV)


# ------------------------------------------------------------------------------------------------
def group_elementwise():
    from flux_fp8_api_b200 import ops
    from flux_fp8_api_b200.f8linear import mul_scale
    from oracle import flux_oracle as O
    print("device SMs:", ops.device_check())
    g = torch.Generator(device=DEV).manual_seed(1)
    x = (torch.randn(1 << 20, device=DEV, generator=g) * 3).to(BF16)
    gold = torch.load(os.path.join(ROOT, "tests/golden/quantize.pt"))
    xs = gold["x"].to(DEV)
    for dt in (E5M2, E4M3):
        for amax in (0.003, 0.7, 5.0, 300.0):
            s = O.amax_to_scale(scalar(amax), torch.finfo(dt).max)
            # torch on CUDA rounds a 0-dim CUDA fp32 scale to the tensor dtype (bf16) before the multiply;
            # torch on CPU keeps it in fp32.  mul_scale() applies the CUDA convention (numerics).
            sq = mul_scale(s)
            cmp(f"quantize rand {dt} amax={amax} (cuda semantics)", ops.quantize(x, sq, dt).view(torch.uint8), O.quantize(x, s, dt).view(torch.uint8), 0, exact=True)
            cmp(f"quantize special {dt} amax={amax} (cuda semantics)", ops.quantize(xs, sq, dt).view(torch.uint8), O.quantize(xs, s, dt).view(torch.uint8), 0, exact=True)
            cmp(f"quantize rand {dt} amax={amax} (cpu semantics)", ops.quantize(x, s, dt).view(torch.uint8).cpu(), O.quantize(x.cpu(), s.cpu(), dt).view(torch.uint8), 0, exact=True)
    for c in gold["cases"]:
        dt = E5M2 if "e5m2" in c["dtype"] else E4M3
        cmp(f"quantize golden {c['dtype']} amax={c['amax']}", ops.quantize(xs, c["scale"].to(DEV), dt).view(torch.uint8).cpu(), c["y"], 0, exact=True)
    cmp("amax", ops.amax(x), x.abs().max().float(), 0, exact=True)
    cmp("amax odd n", ops.amax(x[:1000003 % x.numel()][:12345]), x[:12345].abs().max().float(), 0, exact=True)
    ms = timed(lambda: ops.quantize(x, s, E5M2))
    print(f"  quantize 1Mi: {ms*1e3:.1f} us")

    # silu + quant
    v = (torch.randn(4, 3072, device=DEV, generator=g) * 2).to(BF16)
    s = O.amax_to_scale(scalar(4.0), 57344.0)
    yq, yb = ops.silu_quant(v, s, E5M2, want_bf16=True)
    cmp("silu bf16", yb, torch.nn.functional.silu(v), 2 ** -7, 0.01)
    cmp("silu quant", yq.float(), O.quantize(torch.nn.functional.silu(v), s, E5M2).float(), 1e9, 0.01)

    # LN + modulate + quant
    for (B, L, D) in [(1, 512, 3072), (2, 384, 3072), (2, 100, 256)]:
        xx = (torch.randn(B, L, D, device=DEV, generator=g) * 2 + 0.3).to(BF16)
        mod = (torch.randn(B, 1, 6 * D, device=DEV, generator=g) * 0.3).to(BF16)
        shift, scale = mod[..., :D], mod[..., D:2 * D]
        ref = O.layernorm_modulate(xx, shift, scale)
        s = O.amax_to_scale(ref.abs().max().float(), 57344.0)
        yq, yb = ops.ln_mod_quant(xx, shift, scale, mul_scale(s), E5M2, want_bf16=True)
        cmp(f"ln_mod bf16 B{B} L{L} D{D}", yb, ref, 2 ** -4, 0.02)
        dq = (yq.float() - O.quantize(ref, s, E5M2).float()).abs() / s
        print(f"    ln_mod fp8 dequantised max|d|={dq.max().item():.4e} mismatch={(dq > 0).float().mean().item()*100:.3f}%")
    xx = (torch.randn(1, 4608, 3072, device=DEV, generator=g)).to(BF16)
    mod = (torch.randn(1, 1, 6 * 3072, device=DEV, generator=g) * 0.3).to(BF16)
    ms = timed(lambda: ops.ln_mod_quant(xx, mod[..., :3072], mod[..., 3072:6144], s, E5M2))
    print(f"  ln_mod_quant 4608x3072: {ms*1e3:.1f} us  ({4608*3072*3/ms/1e6:.0f} GB/s)")

    # qknorm + rope (golden from the reference)
    ops_g = torch.load(os.path.join(ROOT, "tests/golden/ops.pt"))
    pe = ops_g["pe"].to(DEV)
    cos, sin = pe[:, 0, :, :, 0, 0].contiguous(), pe[:, 0, :, :, 1, 0].contiguous()
    q = ops_g["q"].to(DEV)
    cmp("rope only (golden)", ops.qknorm_rope(q, None, cos, sin).cpu(), ops_g["q_rope"], 0, exact=True)
    cmp("qknorm only (golden)", ops.qknorm_rope(q, ops_g["qnorm_w"].float().to(DEV), None, None).cpu(), ops_g["q_norm"], 2 ** -6, 0.01)

    # gemv
    for (M, N, K) in [(1, 18432, 3072), (4, 9216, 3072), (3, 512, 256), (6, 3072, 3072)]:
        a, w = rand_fp8((M, K), E5M2, 4.0, 3), rand_fp8((N, K), E4M3, 1.0, 4)
        bias = (torch.randn(N, device=DEV, generator=g) * 0.1).to(BF16)
        sa, sw = scalar(1 / 900.0), scalar(1 / 50.0)
        ref = O.scaled_mm(a, w, sa, sw, bias)
        cmp(f"gemv M{M} N{N} K{K}", ops.f8_gemv(a, w, bias, sa, sw), ref, 2 ** -6 * max(1, ref.abs().max().item()), 0.01)
    a, w = rand_fp8((1, 3072), E5M2, 4.0, 3), rand_fp8((18432, 3072), E4M3, 1.0, 4)
    ms = timed(lambda: ops.f8_gemv(a, w, None, sa, sw))
    print(f"  gemv 1x18432x3072: {ms*1e3:.1f} us ({18432*3072/ms/1e6:.0f} GB/s)")


def gemm_ref(a, w, sa, sw, bias):
    from oracle import flux_oracle as O
    return O.scaled_mm(a, w, sa, sw, bias)


def group_gemm():
    from flux_fp8_api_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(2)
    torch.backends.cuda.matmul.allow_tf32 = False
    shapes = [(128, 256, 128, E5M2), (128, 128, 128, E5M2), (256, 512, 384, E4M3), (200, 384, 320, E5M2),
              (384, 1024, 1024, E5M2), (100, 72, 64, E5M2), (512, 9216, 3072, E5M2), (4608, 3072, 3072, E5M2),
              (4096, 3072, 12288, E5M2), (4608, 21504, 3072, E5M2), (4608, 3072, 15360, E5M2)]
    for (M, N, K, adt) in shapes:
        a, w = rand_fp8((M, K), adt, 4.0, 5), rand_fp8((N, K), E4M3, 1.0, 6)
        bias = (torch.randn(N, device=DEV, generator=g) * 0.5).to(BF16)
        sa, sw = scalar(1 / 64.0), scalar(1 / 32.0)
        ref = gemm_ref(a, w, sa, sw, bias)
        out = torch.empty((M, N), dtype=BF16, device=DEV)
        from flux_fp8_api_b200 import _cabi as cabi
        gg = ops.gemm_args(a, w, bias, sa, sw, cabi.EPI_PLAIN)
        gg.out, gg.ldo = out.data_ptr(), out.stride(0)
        ops.run_gemm(gg)
        ok = cmp(f"gemm plain M{M} N{N} K{K} {adt}", out, ref, 2 ** -7 * max(1, ref.abs().max().item()), 0.01)
        if not ok and M * N <= 256 * 1024:
            block_report(out, ref)
        if M * N * K >= 2 ** 30:
            ms = timed(lambda: ops.run_gemm(gg), iters=20)
            print(f"      {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s")
    # no-bias
    a, w = rand_fp8((256, 512), E5M2, 4.0, 5), rand_fp8((256, 512), E4M3, 1.0, 6)
    cmp("gemm plain no bias", ops.f8_gemm(a, w, None, sa, sw), gemm_ref(a, w, sa, sw, None), 2 ** -7 * 8, 0.01)


def group_epilogue():
    from flux_fp8_api_b200 import ops
    from flux_fp8_api_b200.f8linear import mul_scale
    from oracle import flux_oracle as O
    import torch.nn.functional as F
    g = torch.Generator(device=DEV).manual_seed(3)
    sa, sw = scalar(1 / 64.0), scalar(1 / 32.0)
    # gate residual
    for (B, L, N, K) in [(2, 256, 512, 256), (1, 4608, 3072, 3072)]:
        M = B * L
        a, w = rand_fp8((M, K), E5M2, 4.0, 7), rand_fp8((N, K), E4M3, 1.0, 8)
        bias = (torch.randn(N, device=DEV, generator=g) * 0.5).to(BF16)
        resid = torch.randn(M, N, device=DEV, generator=g).to(BF16)
        mod = torch.randn(B, 3 * N, device=DEV, generator=g).to(BF16)
        gate = mod[:, N:2 * N]
        y = gemm_ref(a, w, sa, sw, bias)
        ref = (resid.view(B, L, N) + gate[:, None, :] * y.view(B, L, N)).view(M, N)
        out = ops.f8_gemm_gate_residual(a, w, bias, sa, sw, resid, gate, L)
        cmp(f"gate_residual B{B} L{L} N{N} K{K}", out, ref, 2 ** -6 * max(1, ref.abs().max().item()), 0.01)
        r2 = resid.clone()
        ops.f8_gemm_gate_residual(a, w, bias, sa, sw, r2, gate, L, out=r2)
        cmp("gate_residual in-place", r2, ref, 2 ** -6 * max(1, ref.abs().max().item()), 0.01)
        if M * N * K >= 2 ** 30:
            ms = timed(lambda: ops.f8_gemm_gate_residual(a, w, bias, sa, sw, resid, gate, L, out=out), iters=20)
            print(f"      {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s")
    # gelu + quant
    for (M, N, K) in [(256, 512, 256), (4608, 12288, 3072)]:
        a, w = rand_fp8((M, K), E5M2, 4.0, 9), rand_fp8((N, K), E4M3, 1.0, 10)
        bias = (torch.randn(N, device=DEV, generator=g) * 0.5).to(BF16)
        y = gemm_ref(a, w, sa, sw, bias)
        ge = F.gelu(y, approximate="tanh")
        so = O.amax_to_scale(ge.abs().max().float(), 57344.0)
        ref = O.quantize(ge, so, E5M2)
        out = ops.f8_gemm_gelu_quant(a, w, bias, sa, sw, mul_scale(so), E5M2)
        d = (out.float() - ref.float()).abs() / so
        mism = (d > 0).float().mean().item()
        ok = mism < 0.02 and not torch.isnan(out.float()).any()
        print(f"  [{'PASS' if ok else 'FAIL'}] gelu_quant M{M} N{N} K{K}: dequantised max|d|={d.max().item():.4e} mismatch={mism*100:.3f}%")
        if not ok:
            FAILS.append("gelu_quant")
        if M * N * K >= 2 ** 30:
            ms = timed(lambda: ops.f8_gemm_gelu_quant(a, w, bias, sa, sw, mul_scale(so), E5M2, out=out), iters=20)
            print(f"      {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s")
    # qkv + rope  /  linear1
    for (B, L, T, H, K, mlp) in [(2, 128, 64, 2, 256, 0), (1, 256, 128, 2, 256, 512), (1, 4096, 512, 24, 3072, 12288)]:
        S = L + T
        D = H * 128
        N = 3 * D + mlp
        M = B * L
        a, w = rand_fp8((M, K), E5M2, 4.0, 11), rand_fp8((N, K), E4M3, 0.5, 12)
        bias = (torch.randn(N, device=DEV, generator=g) * 0.5).to(BF16)
        qw = (1 + 0.05 * torch.randn(128, device=DEV, generator=g)).to(BF16)
        kw = (1 + 0.05 * torch.randn(128, device=DEV, generator=g)).to(BF16)
        hh = 1 << (int(math.log2(L)) // 2)
        ids = torch.cat((torch.zeros(1, T, 3), O.make_img_ids(1, hh, L // hh, torch.float32)), 1).to(BF16).to(DEV)
        pe = O.embed_nd(ids, [16, 56, 56], 10000, BF16)  # [1,1,S,64,2,2]
        cos, sin = pe[:, 0, :, :, 0, 0].contiguous(), pe[:, 0, :, :, 1, 0].contiguous()
        q = torch.zeros(B, H, S, 128, dtype=BF16, device=DEV)
        k, v = torch.zeros_like(q), torch.zeros_like(q)
        y = gemm_ref(a, w, sa, sw, bias).view(B, L, N)
        rq, rk, rv = O.split_heads(y[..., :3 * D], H)
        rq, rk = O.rms_norm(rq, qw), O.rms_norm(rk, kw)
        rq, rk = O.apply_rope(rq, rk, pe[:, :, T:])
        mlp_out, so = None, None
        if mlp:
            ge = F.gelu(y[..., 3 * D:], approximate="tanh")
            so = O.amax_to_scale(ge.abs().max().float(), 57344.0)
            mlp_out = torch.zeros(M, D + mlp, dtype=E5M2, device=DEV)
        ops.f8_gemm_qkv_rope(a, w, bias, sa, sw, q, k, v, qw.float(), kw.float(), cos, sin, L, T,
                             mlp_out=mlp_out, mlp_scale=mul_scale(so) if so is not None else None, mlp_col_offset=D)
        tag = f"B{B} L{L} T{T} H{H} K{K} mlp{mlp}"
        cmp(f"qkv_rope q {tag}", q[:, :, T:], rq, 2 ** -5 * max(1, rq.abs().max().item()), 0.02)
        cmp(f"qkv_rope k {tag}", k[:, :, T:], rk, 2 ** -5 * max(1, rk.abs().max().item()), 0.02)
        cmp(f"qkv_rope v {tag}", v[:, :, T:], rv, 2 ** -6 * max(1, rv.abs().max().item()), 0.02)
        cmp(f"qkv_rope untouched txt rows {tag}", q[:, :, :T], torch.zeros_like(q[:, :, :T]), 0, exact=True)
        if mlp:
            ref = O.quantize(ge, so, E5M2).view(M, mlp)
            d = (mlp_out[:, D:].float() - ref.float()).abs() / so
            print(f"    linear1 mlp part: dequantised max|d|={d.max().item():.4e} mismatch={(d > 0).float().mean().item()*100:.3f}%  "
                  f"attn cols untouched={bool((mlp_out[:, :D].float() == 0).all())}")
        if M * N * K >= 2 ** 30:
            ms = timed(lambda: ops.f8_gemm_qkv_rope(a, w, bias, sa, sw, q, k, v, qw.float(), kw.float(), cos, sin, L, T,
                                                    mlp_out=mlp_out, mlp_scale=so, mlp_col_offset=D), iters=20)
            print(f"      {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s")


def attn_ref(q, k, v):
    from oracle import flux_oracle as O
    B, H, S, D = q.shape
    outs = []
    for h0 in range(0, H, 4):
        outs.append(O.sdpa(q[:, h0:h0 + 4], k[:, h0:h0 + 4], v[:, h0:h0 + 4]))
    x = torch.cat(outs, 1).transpose(1, 2)
    return x.reshape(B, S, H * D)


def group_attention(variants=(0, 1, 2, 3, 4, 5, 6)):
    from flux_fp8_api_b200 import ops
    from oracle import flux_oracle as O
    g = torch.Generator(device=DEV).manual_seed(4)
    for variant in variants:
        print(f" -- attention variant {variant}")
        for (B, H, S, std) in [(1, 1, 128, 1.0), (1, 2, 256, 1.0), (2, 2, 384, 2.0), (1, 2, 320, 1.0), (1, 3, 1000, 1.5),
                               (1, 24, 4608, 1.0)]:
            q = (torch.randn(B, H, S, 128, device=DEV, generator=g) * std).to(BF16)
            k = (torch.randn(B, H, S, 128, device=DEV, generator=g) * std).to(BF16)
            v = torch.randn(B, H, S, 128, device=DEV, generator=g).to(BF16)
            ref = attn_ref(q, k, v)
            try:
                out = ops.attention(q, k, v, variant=variant)
                torch.cuda.synchronize()
            except Exception as ex:  # noqa: BLE001
                print(f"  [FAIL] attention v{variant} B{B} H{H} S{S}: {ex}")
                FAILS.append(f"attention v{variant}")
                return
            ok = cmp(f"attention v{variant} B{B} H{H} S{S} std{std}", out, ref, 2 ** -6, 1.0)
            if not ok and S <= 256:
                block_report(out[0], ref[0])
            if S >= 4096:
                ms = timed(lambda: ops.attention(q, k, v, out=out, variant=variant), iters=10)
                print(f"      {ms*1e3:8.1f} us  {4*B*H*S*S*128/ms/1e9:8.1f} TFLOP/s")
                # fp8 output
                s0, s1 = O.amax_to_scale(scalar(1.0), 57344.0), O.amax_to_scale(scalar(2.0), 57344.0)
                o8 = torch.zeros(B, S, H * 128 + 256, dtype=E5M2, device=DEV)
                from flux_fp8_api_b200.f8linear import mul_scale
                ops.attention(q, k, v, out=o8[..., : H * 128], out_scale0=mul_scale(s0), out_scale1=mul_scale(s1), split_row=512, variant=variant)
                ref8 = torch.cat((O.quantize(ref[:, :512], s0, E5M2), O.quantize(ref[:, 512:], s1, E5M2)), 1)
                d = (o8[..., : H * 128].float() - ref8.float()).abs()
                d[:, :512] /= s0
                d[:, 512:] /= s1
                print(f"      fp8 out: dequantised max|d|={d.max().item():.4e} mismatch={(d > 0).float().mean().item()*100:.2f}% "
                      f"pad untouched={bool((o8[..., H*128:].float() == 0).all())}")


def group_model():
    """Block / model level: tiny golden model (reference outputs minted on CPU) through fused and eager paths."""
    import flux_fp8_api_b200.f8linear as f8
    from flux_fp8_api_b200 import blocks, model as M, pipeline as PL
    from oracle import flux_oracle as O
    gold = torch.load(os.path.join(ROOT, "tests/golden/flux_tiny.pt"))
    tiny = gold["tiny"]
    spec = M.FluxSpec(params=M.FluxParams(**tiny), prequantized_flow=True, quantize_modulation=True,
                      quantize_flow_embedder_layers=False)
    f8.SCALE_SEMANTICS = "cpu"  # the golden outputs come from the reference run on CPU
    with torch.device(DEV):
        net = M.Flux(spec, dtype=BF16).to(BF16)
    net.load_state_dict(gold["state"], strict=True)
    net = net.to(DEV).eval()
    print("  all frozen:", PL.all_frozen(net))
    inp = {k: v.to(DEV) for k, v in gold["inputs"].items()}
    bi = {k: v.to(DEV) for k, v in gold["block_in"].items()}
    with torch.inference_mode():
        for mode in ("fused", "eager"):
            if mode == "eager":
                blocks.DoubleStreamBlock._fusable = lambda self, a, b: False
                blocks.SingleStreamBlock._fusable = lambda self, a: False
            d_img, d_txt = net.double_blocks[0](img=bi["img"], txt=bi["txt"], vec=bi["vec"], pe=bi["pe"])
            cmp(f"[{mode}] DoubleStreamBlock img vs reference", d_img.cpu(), gold["double_img"], 2 ** -4, 0.2)
            cmp(f"[{mode}] DoubleStreamBlock txt vs reference", d_txt.cpu(), gold["double_txt"], 2 ** -4, 0.2)
            xs = torch.cat((bi["txt"], bi["img"]), 1)
            s_out = net.single_blocks[0](xs, vec=bi["vec"], pe=bi["pe"])
            cmp(f"[{mode}] SingleStreamBlock vs reference", s_out.cpu(), gold["single"], 2 ** -4, 0.2)
            y = net(**inp)
            cmp(f"[{mode}] Flux.forward fp8 vs reference", y.cpu(), gold["y_fp8"], 2 ** -4, 0.6)
            y2 = net(**inp)
            cmp(f"[{mode}] Flux.forward deterministic", y2, y, 0, exact=True)
    f8.SCALE_SEMANTICS = "cuda"
    # batched modulation (one launch pair for every Modulation.lin) vs the per-layer path
    with torch.inference_mode():
        mods = [net.double_blocks[0].img_mod, net.double_blocks[0].txt_mod, net.single_blocks[0].modulation]
        bank = blocks.ModulationBank(mods)
        for B in (1, 2, 5):
            vec = torch.randn(B, net.hidden_size, device=DEV).to(BF16)
            got = bank(vec)
            for m, (o1, o2) in zip(mods, got):
                r1, r2 = m(vec)
                for a_, b_ in zip(o1, r1):
                    cmp(f"ModulationBank B{B} vs Modulation.forward", a_, b_, 2 ** -7 * max(1, b_.abs().max().item()), 0.01)
                if r2 is not None:
                    for a_, b_ in zip(o2, r2):
                        cmp(f"ModulationBank B{B} (second half)", a_, b_, 2 ** -7 * max(1, b_.abs().max().item()), 0.01)


def group_modbank():
    """Full-width batched modulation (K = 3072; 4 double-type + 8 single-type layers) against the per-layer GEMV path,
    batch 1 / 3 / 8 / 11, and its weight-streaming rate.  FLUXB200_GEMV_MMA=0 selects the CUDA-core kernel."""
    from flux_fp8_api_b200 import blocks
    from flux_fp8_api_b200.f8linear import F8Linear
    torch.manual_seed(5)
    mods = []
    for i in range(12):
        m = blocks.Modulation(3072, double=i < 4).to(DEV).to(BF16)
        torch.nn.init.normal_(m.lin.weight, std=0.01)
        torch.nn.init.normal_(m.lin.bias, std=0.02)
        m.lin = F8Linear.from_linear(m.lin, input_float8_dtype=torch.float8_e4m3fn if i % 2 else torch.float8_e5m2)
        mods.append(m)
    with torch.inference_mode():
        cal = torch.randn(1, 3072, device=DEV).to(BF16)
        for _ in range(13):
            for m in mods:
                m(cal)
        for dt in (torch.float8_e5m2, torch.float8_e4m3fn):
            sel = [m for m in mods if m.lin.input_float8_dtype == dt]
            bank = blocks.ModulationBank(sel)
            for B in (1, 3, 8, 11):
                vec = torch.randn(B, 3072, device=DEV).to(BF16)
                got = bank(vec)
                for m, (o1, o2) in zip(sel, got):
                    r1, r2 = m(vec)
                    a_ = torch.cat(tuple(o1) + (tuple(o2) if o2 else ()), -1)
                    b_ = torch.cat(tuple(r1) + (tuple(r2) if r2 else ()), -1)
                    cmp(f"bank {str(dt)[-4:]} B{B} N{m.lin.out_features}", a_, b_, 2 ** -7 * max(1, b_.abs().max().item()), 0.01)
        bank = blocks.ModulationBank([m for m in mods if m.lin.input_float8_dtype == torch.float8_e5m2])
        vec = torch.randn(1, 3072, device=DEV).to(BF16)
        nbytes = bank.total_n * bank.K
        for _ in range(3):
            bank(vec)
        torch.cuda.synchronize()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # stream more than L2 between repeats: 6 layers x ~40 MB = 240 MB per call, still flush explicitly
        flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=DEV)
        tot = 0.0
        for _ in range(10):
            flush.zero_()
            s_.record()
            bank(vec)
            e_.record()
            torch.cuda.synchronize()
            tot += s_.elapsed_time(e_)
        print(f"  batched modulation {nbytes/1e6:.0f} MB: {tot/10*1e3:.1f} us ({nbytes/(tot/10*1e-3)/1e9:.0f} GB/s) "
              f"[FLUXB200_GEMV_MMA={os.environ.get('FLUXB200_GEMV_MMA', '1')}]")


def group_flux():
    """Full-size Flux-dev 1024x1024: synthetic weights -> quantise -> calibrate (eager) -> fused denoise timing."""
    from flux_fp8_api_b200 import model as M, pipeline as PL
    t0 = time.time()
    spec = M.flux_dev_spec()
    net = PL.build_synthetic_flux(spec, DEV)
    torch.cuda.synchronize()
    print(f"  build+quantise: {time.time()-t0:.1f} s, mem {torch.cuda.memory_allocated()/2**30:.1f} GiB")
    req = PL.synthetic_request(spec.params, 1024, 1024, 1, 512, DEV, seed=0)
    t0 = time.time()
    PL.calibrate(net, req, num_steps=13)
    torch.cuda.synchronize()
    print(f"  calibration (13 eager steps): {time.time()-t0:.1f} s")
    scales = [m.input_scale.item() for m in net.modules() if hasattr(m, "input_scale") and m.input_scale is not None]
    print(f"  input scales: n={len(scales)} min={min(scales):.1f} max={max(scales):.1f}")
    sched = PL.get_schedule(28, req["img"].shape[1])
    with torch.inference_mode():
        def one_step(img, t):
            tv = torch.full((1,), t, dtype=BF16, device=DEV)
            return net(img=img, img_ids=req["img_ids"], txt=req["txt"], txt_ids=req["txt_ids"], y=req["y"], timesteps=tv,
                       guidance=req["guidance"])
        pred = one_step(req["img"], sched[0])
        print(f"  pred: amax={pred.abs().max().item():.3f} std={pred.float().std().item():.3f} nan={torch.isnan(pred.float()).sum().item()}")
        ms = timed(lambda: one_step(req["img"], sched[0]), iters=5, warm=2)
        print(f"  fused eager-launch step: {ms:.2f} ms  ({1000/ms:.2f} it/s)")
        gs = PL.GraphedStep(net, req)
        tv = torch.full((1,), sched[0], dtype=BF16, device=DEV)
        out_g = gs(req["img"], tv, sched[1] - sched[0])
        out_e = req["img"] + (sched[1] - sched[0]) * one_step(req["img"], sched[0])
        cmp("graphed step == eager-launch step", out_g, out_e, 0, exact=True)
        ms = timed(lambda: gs(req["img"], tv, sched[1] - sched[0]), iters=10, warm=2)
        print(f"  CUDA-graph step: {ms:.2f} ms  ({1000/ms:.2f} it/s)")
        t0 = time.time()
        img = PL.denoise(net, req, sched, step_fn=gs)
        torch.cuda.synchronize()
        print(f"  28-step denoise (graph): {time.time()-t0:.2f} s; final latent amax={img.abs().max().item():.3f}")


def stats(name, got, ref):
    d = (got.float() - ref.float()).abs().flatten()
    r = ref.float().abs()
    q = torch.quantile(d[: 4_000_000], torch.tensor([0.5, 0.99, 0.9999], device=d.device)).tolist()
    print(f"  {name:34s} ref amax {r.max().item():8.3f} rms {ref.float().pow(2).mean().sqrt().item():7.3f} | err mean {d.mean().item():.3e} "
          f"p50 {q[0]:.2e} p99 {q[1]:.2e} p99.99 {q[2]:.2e} max {d.max().item():.3e} | frac>2^-6 {(d > 2**-6).float().mean().item():.2e}")


def group_fullwidth():
    """Full-width (3072 x 24 heads) depth-1 model: fused path vs the oracle on the same GPU tensors, stage by stage."""
    from flux_fp8_api_b200 import model as M, pipeline as PL, blocks
    from oracle import flux_oracle as O
    params = M.FluxParams(depth=1, depth_single_blocks=1)
    net = PL.build_synthetic_flux(M.FluxSpec(params=params), DEV, seed=7)
    req = PL.synthetic_request(params, 1024, 1024, 1, 512, DEV, seed=3)
    PL.calibrate(net, req, num_steps=13)
    sd = {k: v for k, v in net.state_dict().items() if v is not None}
    g = torch.Generator(device=DEV).manual_seed(11)
    with torch.inference_mode():
        img = torch.randn(1, 4096, 3072, device=DEV, generator=g).to(BF16)
        txt = torch.randn(1, 512, 3072, device=DEV, generator=g).to(BF16)
        vec = torch.randn(1, 3072, device=DEV, generator=g).to(BF16)
        pe = net.pe_embedder(torch.cat((req["txt_ids"], req["img_ids"]), 1))
        for mode in ("fused", "eager"):
            if mode == "eager":
                blocks.DoubleStreamBlock._fusable = lambda self, a, b: False
                blocks.SingleStreamBlock._fusable = lambda self, a: False
            print(f" -- {mode}")
            d_img, d_txt = net.double_blocks[0](img=img, txt=txt, vec=vec, pe=pe)
            o_img, o_txt = O.double_block(img, txt, vec, pe, sd, "double_blocks.0.", 24)
            stats("double block img", d_img, o_img)
            stats("double block txt", d_txt, o_txt)
            x = torch.cat((txt, img), 1)
            s_out = net.single_blocks[0](x, vec=vec, pe=pe)
            stats("single block", s_out, O.single_block(x, vec, pe, sd, "single_blocks.0.", 24))
            t = torch.full((1,), 0.7, dtype=BF16, device=DEV)
            ours = net(img=req["img"], img_ids=req["img_ids"], txt=req["txt"], txt_ids=req["txt_ids"], timesteps=t, y=req["y"], guidance=req["guidance"])
            cfg = dict(num_heads=24, depth=1, depth_single_blocks=1, axes_dim=[16, 56, 56], theta=10_000, guidance_embed=True)
            ref = O.flux_forward(sd, cfg, req["img"], req["img_ids"], req["txt"], req["txt_ids"], t, req["y"], req["guidance"])
            stats("Flux.forward (1+1 blocks)", ours, ref)
            if mode == "fused":
                # the reference's own noise floor: the same oracle with the other legitimate SDPA rounding
                O.SDPA_P_DTYPE = "bf16"
                n_img, n_txt = O.double_block(img, txt, vec, pe, sd, "double_blocks.0.", 24)
                n_single = O.single_block(x, vec, pe, sd, "single_blocks.0.", 24)
                n_ref = O.flux_forward(sd, cfg, req["img"], req["img_ids"], req["txt"], req["txt_ids"], t, req["y"], req["guidance"])
                O.SDPA_P_DTYPE = "fp32"
                stats("noise floor: double block img", n_img, o_img)
                stats("noise floor: single block", n_single, O.single_block(x, vec, pe, sd, "single_blocks.0.", 24))
                stats("noise floor: Flux.forward", n_ref, ref)
                stats("ours vs bf16-P oracle: dbl img", d_img, n_img)
                stats("ours vs bf16-P oracle: forward", ours, n_ref)


GROUPS = {"elementwise": group_elementwise, "fullwidth": group_fullwidth, "modbank": group_modbank, "model": group_model, "flux": group_flux, "gemm": group_gemm, "epilogue": group_epilogue,
          "attention": group_attention}


def main():
    if len(sys.argv) > 1:
        name = sys.argv[1]
        if name.startswith("attention") and len(name) > len("attention"):
            tail = name[len("attention"):]
            group_attention(tuple(int(c) for c in tail[1:].split(",")) if tail[0] == ":" else tuple(int(c) for c in tail))
        else:
            GROUPS[name]()
        print(f"== {name}: {'OK' if not FAILS else 'FAILED: ' + ', '.join(FAILS)}")
        sys.exit(1 if FAILS else 0)
    rc = 0
    base_env = dict(os.environ)
    for name in ["attention0", "attention1", "model", "flux"]:
        print(f"===== {name} =====", flush=True)
        t0 = time.time()
        try:
            env = dict(base_env)
            if name.endswith("_cg2"):
                env["FLUXB200_GEMM_CG"] = "2"
            p = subprocess.run([sys.executable, os.path.abspath(__file__), name.replace("_cg2", "")], timeout=420, env=env)
            rc |= p.returncode
        except subprocess.TimeoutExpired:
            print(f"== {name}: TIMEOUT")
            rc |= 1
        print(f"   ({time.time()-t0:.0f} s)", flush=True)
    sys.exit(rc)


if __name__ == "__main__":
    main()
