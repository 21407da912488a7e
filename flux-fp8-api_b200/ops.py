"""Tensor-level wrappers over the C ABI (one function per entry point of include/flux_b200.h).

These take torch CUDA tensors, allocate outputs with torch, and pass raw pointers plus the current
stream to libflux_b200.so.  They do no arithmetic themselves.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch

from . import _cabi as cabi

Tensor = torch.Tensor
BF16 = torch.bfloat16

#: when set to a list, the GEMM / attention / LayerNorm wrappers append (kind, work, start_event, end_event) per launch
#: (work = flops, or bytes for the HBM-bound kernels): CUDA-event timing of individual kernels on the launching
#: stream (bench.py's roofline pass)
KERNEL_TIMELINE = None


def _want(t: Optional[Tensor], dtype: torch.dtype, what: str) -> None:
    """The kernels take raw pointers: an operand of another dtype would be silently mis-read, so refuse it here."""
    if t is not None and t.dtype != dtype:
        raise ValueError(f"{what} must be {dtype}, got {t.dtype}")


def _contig(t: Tensor, what: str) -> Tensor:
    if not t.is_contiguous():
        raise ValueError(f"{what} must be contiguous")
    return t


def device_check() -> int:
    n = C.c_int(0)
    cabi.check(cabi.load().fluxb200_device_check(C.byref(n)), "fluxb200_device_check")
    return n.value


def gemm_force_tiling(cta_group: int = 0, pairs_per_cluster: int = 0) -> None:
    """Measurement / test hook (fluxb200_gemm_force_tiling): (0, 0) restores the library's own choice."""
    rc = cabi.load().fluxb200_gemm_force_tiling(cta_group, pairs_per_cluster)
    if rc == cabi.ERR_UNSUPPORTED:
        raise NotImplementedError(cabi.load().fluxb200_last_error().decode())
    if rc:
        raise ValueError(cabi.load().fluxb200_last_error().decode())


def quantize(x: Tensor, scale: Tensor, dtype: torch.dtype, out: Optional[Tensor] = None) -> Tensor:
    """F8Linear.to_fp8_saturated(...).to(fp8)  (reference float8_quantize.py:217-218, 274-276)."""
    cabi.require_cuda(x, scale)
    if x.dtype != BF16:
        raise ValueError(f"quantize expects bfloat16 activations, got {x.dtype}")
    x = x.contiguous()
    if out is None:
        out = torch.empty(x.shape, dtype=dtype, device=x.device)
    if x.numel() == 0:
        return out
    cabi.check(cabi.load().fluxb200_quantize(x.data_ptr(), out.data_ptr(), x.numel(), scale.data_ptr(),
                                             cabi.fp8_fmt(dtype), cabi.stream_ptr()), "fluxb200_quantize")
    return out


def amax(x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """torch.max(torch.abs(x)).float()  (reference float8_quantize.py:197, 227)."""
    cabi.require_cuda(x)
    if x.dtype != BF16:
        raise ValueError(f"amax expects bfloat16, got {x.dtype}")
    x = x.contiguous()
    if out is None:
        out = torch.zeros((), dtype=torch.float32, device=x.device)
    if x.numel() == 0:
        return out
    cabi.check(cabi.load().fluxb200_amax(x.data_ptr(), x.numel(), out.data_ptr(), cabi.stream_ptr()), "fluxb200_amax")
    return out


def lora_fuse(w_fp8: Tensor, w_scale_recip: Tensor, lora_down: Tensor, lora_up: Tensor, coeff: float,
              unfuse: bool = False, chunks: int = 1):
    """Dequantise + rank-R update + bf16 rounding + amax in one pass (fluxb200_lora_fuse; reference
    lora_loading.py:615-626, 509-577 and the amax of float8_quantize.py:196-198).
    lora_down [N, R] fp32, lora_up [chunks*R, K] fp32.  Returns (W' bf16 [N, K], amax 0-dim fp32)."""
    cabi.require_cuda(w_fp8, w_scale_recip, lora_down, lora_up)
    N, K = w_fp8.shape
    R = lora_down.shape[1]
    if lora_down.dtype != torch.float32 or lora_up.dtype != torch.float32:
        raise ValueError("lora_fuse expects fp32 LoRA factors (the reference computes the delta in fp32)")
    if lora_down.shape != (N, R) or lora_up.shape != (chunks * R, K):
        raise ValueError(f"lora_fuse: factors {tuple(lora_down.shape)} x {tuple(lora_up.shape)} do not match a "
                         f"[{N}, {K}] weight with {chunks} chunk(s)")
    w_fp8, lora_down, lora_up = w_fp8.contiguous(), lora_down.contiguous(), lora_up.contiguous()
    out = torch.empty((N, K), dtype=BF16, device=w_fp8.device)
    amax = torch.empty((), dtype=torch.float32, device=w_fp8.device)
    cabi.check(cabi.load().fluxb200_lora_fuse(w_fp8.data_ptr(), cabi.fp8_fmt(w_fp8.dtype), w_scale_recip.data_ptr(),
                                              lora_down.data_ptr(), lora_up.data_ptr(), N, K, R, chunks, float(coeff),
                                              1 if unfuse else 0, out.data_ptr(), amax.data_ptr(), cabi.stream_ptr()),
               "fluxb200_lora_fuse")
    return out, amax


def silu_quant(x: Tensor, scale: Optional[Tensor], dtype: Optional[torch.dtype], want_bf16: bool = False):
    """Modulation prologue: (fp8 quantised silu(x), bf16 silu(x))."""
    cabi.require_cuda(x)
    x = x.contiguous()
    yq = torch.empty(x.shape, dtype=dtype, device=x.device) if dtype is not None else None
    yb = torch.empty_like(x) if want_bf16 else None
    cabi.check(cabi.load().fluxb200_silu_quant(x.data_ptr(), cabi.ptr(yq), cabi.ptr(yb), x.numel(), cabi.ptr(scale),
                                               cabi.fp8_fmt(dtype) if dtype is not None else 0, cabi.stream_ptr()),
               "fluxb200_silu_quant")
    return yq, yb


def bf16_gemv(x: Tensor, weight: Tensor, bias: Optional[Tensor], silu_input: bool = False,
              add0: Optional[Tensor] = None, add1: Optional[Tensor] = None) -> Tensor:
    """F.linear(f(x), weight, bias) [+ add0] [+ add1] for a skinny bf16 input x [B <= 16, K] (fluxb200_bf16_gemv):
    f = silu when silu_input; every addition is followed by its own bf16 rounding, as the eager ops would."""
    cabi.require_cuda(x, weight)
    for t, what in ((x, "x"), (weight, "weight"), (bias, "bias"), (add0, "add0"), (add1, "add1")):
        _want(t, BF16, f"bf16_gemv: {what}")
    B, K = x.shape
    N = weight.shape[0]
    if weight.shape[1] != K:
        raise ValueError(f"bf16_gemv: x {tuple(x.shape)} does not match weight {tuple(weight.shape)}")
    x, weight = x.contiguous(), _contig(weight, "weight")
    adds = [a.contiguous() if a is not None else None for a in (add0, add1)]
    for a in adds:
        if a is not None and tuple(a.shape) != (B, N):
            raise ValueError(f"bf16_gemv: addend {tuple(a.shape)} is not [{B}, {N}]")
    out = torch.empty((B, N), dtype=BF16, device=x.device)
    _timed("bf16_gemv", 2.0 * N * K,
           lambda: cabi.check(cabi.load().fluxb200_bf16_gemv(x.data_ptr(), weight.data_ptr(), cabi.ptr(bias), cabi.ptr(adds[0]),
                                                             cabi.ptr(adds[1]), N, out.data_ptr(), N, B, N, K,
                                                             1 if silu_input else 0, cabi.stream_ptr()),
                              "fluxb200_bf16_gemv"))
    return out


def bf16_gemm_small(x: Tensor, weight: Tensor, bias: Optional[Tensor], euler_img: Optional[Tensor] = None,
                    euler_dt: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
    """F.linear(x, weight, bias) for the small un-quantised linears around the block stack (fluxb200_bf16_gemm_small);
    with euler_img / euler_dt the result p becomes euler_img + euler_dt * p with eager torch's roundings."""
    cabi.require_cuda(x, weight)
    for t, what in ((x, "x"), (weight, "weight"), (bias, "bias"), (euler_img, "euler_img")):
        _want(t, BF16, f"bf16_gemm_small: {what}")
    _want(euler_dt, torch.float32, "bf16_gemm_small: euler_dt")
    lead, K = x.shape[:-1], x.shape[-1]
    N = weight.shape[0]
    if weight.shape[1] != K:
        raise ValueError(f"bf16_gemm_small: x {tuple(x.shape)} does not match weight {tuple(weight.shape)}")
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    _contig(weight, "weight")
    M = x2.shape[0]
    if out is None:
        out = torch.empty((*lead, N), dtype=BF16, device=x.device)
    _contig(out, "out")
    if euler_img is not None:
        euler_img = euler_img.contiguous()
        if euler_img.numel() != M * N:
            raise ValueError("bf16_gemm_small: euler_img must have the output's shape")
    _timed("bf16_gemm_small", 2.0 * M * N * K,
           lambda: cabi.check(cabi.load().fluxb200_bf16_gemm_small(x2.data_ptr(), x2.stride(0), weight.data_ptr(), cabi.ptr(bias),
                                                                   out.data_ptr(), N, cabi.ptr(euler_img), cabi.ptr(euler_dt),
                                                                   M, N, K, cabi.stream_ptr()), "fluxb200_bf16_gemm_small"))
    return out


_freqs_cache = {}


def timestep_embedding(t: Tensor, dim: int = 256, max_period: int = 10000, time_factor: float = 1000.0) -> Tensor:
    """modules/flux_model.py:95-116 followed by .type(bfloat16): [B] bf16 timesteps -> [B, dim] bf16."""
    cabi.require_cuda(t)
    _want(t, BF16, "timestep_embedding: t")
    if dim % 2:
        raise ValueError("timestep_embedding: odd dim is not supported by the kernel")
    key = (t.device, dim, max_period)
    freqs = _freqs_cache.get(key)
    if freqs is None:
        half = dim // 2
        freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half)
        _freqs_cache[key] = freqs
    t = t.contiguous()
    out = torch.empty((t.shape[0], dim), dtype=BF16, device=t.device)
    cabi.check(cabi.load().fluxb200_timestep_embedding(t.data_ptr(), freqs.data_ptr(), out.data_ptr(), t.shape[0], dim,
                                                       float(time_factor), cabi.stream_ptr()), "fluxb200_timestep_embedding")
    return out


def euler_update(img: Tensor, pred: Tensor, dt: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """img + dt * pred with eager torch's rounding (flux_pipeline.py:651); dt is a 0-dim fp32 device tensor."""
    cabi.require_cuda(img, pred, dt)
    _want(img, BF16, "euler_update: img"), _want(pred, BF16, "euler_update: pred"), _want(dt, torch.float32, "euler_update: dt")
    if img.shape != pred.shape:
        raise ValueError("euler_update: img and pred shapes differ")
    img, pred = img.contiguous(), pred.contiguous()
    if out is None:
        out = torch.empty_like(img)
    _contig(out, "out")
    cabi.check(cabi.load().fluxb200_euler_update(img.data_ptr(), pred.data_ptr(), dt.data_ptr(), out.data_ptr(), img.numel(),
                                                 cabi.stream_ptr()), "fluxb200_euler_update")
    return out


def ln_mod_quant(x: Tensor, shift: Tensor, scale: Tensor, in_scale: Optional[Tensor], dtype: Optional[torch.dtype],
                 out_fp8: Optional[Tensor] = None, want_bf16: bool = False, eps: float = 1e-6):
    """LayerNorm(no affine) -> (1+scale)*x+shift -> quantise.  x [B,L,D]; shift/scale [B,1,D] or [B,D]
    (possibly strided views into the modulation output)."""
    cabi.require_cuda(x, shift, scale)
    _want(x, BF16, "ln_mod_quant: x"), _want(shift, BF16, "ln_mod_quant: shift"), _want(scale, BF16, "ln_mod_quant: scale")
    _want(in_scale, torch.float32, "ln_mod_quant: in_scale")
    B, L, D = x.shape
    if x.stride(-1) != 1 or x.stride(0) != L * x.stride(1):
        x = x.contiguous()
    sh = shift.reshape(B, D) if shift.dim() == 3 else shift
    sc = scale.reshape(B, D) if scale.dim() == 3 else scale
    if sh.stride(-1) != 1 or sc.stride(-1) != 1 or sh.stride(0) != sc.stride(0):
        sh, sc = sh.contiguous(), sc.contiguous()
    yq = out_fp8
    if yq is None and dtype is not None:
        yq = torch.empty((B, L, D), dtype=dtype, device=x.device)
    yb = torch.empty((B, L, D), dtype=BF16, device=x.device) if want_bf16 else None
    _timed("ln_mod_quant", B * L * D * (2.0 + (1.0 if yq is not None else 0.0) + (2.0 if yb is not None else 0.0)),
           lambda: cabi.check(cabi.load().fluxb200_ln_mod_quant(
               x.data_ptr(), x.stride(1), sh.data_ptr(), sc.data_ptr(), sh.stride(0) if B > 1 else D,
               cabi.ptr(yq), yq.stride(-2) if yq is not None else 0, cabi.ptr(yb), D, cabi.ptr(in_scale),
               cabi.fp8_fmt(dtype) if dtype is not None else 0, B, L, D, eps, cabi.stream_ptr()),
               "fluxb200_ln_mod_quant"))
    return yq, yb


def _ln_args(items, dtype: torch.dtype):
    """fluxb200_ln_args array + freshly allocated fp8 outputs for 1 or 2 (x [B,L,D], shift, scale, in_scale) row sets."""
    args = (cabi.LnArgs * len(items))()
    outs, keep = [], []
    D = items[0][0].shape[-1]
    total = 0.0
    for i, (x, shift, scale, in_scale) in enumerate(items):
        cabi.require_cuda(x, shift, scale, in_scale)
        _want(x, BF16, "ln_mod_quant: x"), _want(shift, BF16, "ln_mod_quant: shift")
        _want(scale, BF16, "ln_mod_quant: scale"), _want(in_scale, torch.float32, "ln_mod_quant: in_scale")
        B, L, Dx = x.shape
        if Dx != D:
            raise ValueError("ln_mod_quant: the row sets must share the hidden size")
        if x.stride(-1) != 1 or x.stride(0) != L * x.stride(1):
            x = x.contiguous()
        sh = shift.reshape(B, D) if shift.dim() == 3 else shift
        sc = scale.reshape(B, D) if scale.dim() == 3 else scale
        if sh.stride(-1) != 1 or sc.stride(-1) != 1 or sh.stride(0) != sc.stride(0):
            sh, sc = sh.contiguous(), sc.contiguous()
        yq = torch.empty((B, L, D), dtype=dtype, device=x.device)
        a = args[i]
        a.x, a.shift, a.scale, a.y_fp8, a.in_scale = x.data_ptr(), sh.data_ptr(), sc.data_ptr(), yq.data_ptr(), in_scale.data_ptr()
        a.ldx, a.ldy, a.mod_batch_stride, a.B, a.L = x.stride(1), D, (sh.stride(0) if B > 1 else D), B, L
        outs.append(yq)
        keep += [x, sh, sc, in_scale]
        total += B * L * D * 3.0
    return args, outs, keep, D, total


def ln_mod_quant_pair(items, dtype: torch.dtype, eps: float = 1e-6):
    """Two independent LN-modulate-quantise problems (the txt and img streams of a DoubleStreamBlock) in ONE launch.
    items: [(x [B,L,D], shift, scale, in_scale)] * 2 with the same D; returns the two fp8 tensors."""
    args, outs, keep, D, total = _ln_args(items, dtype)
    _timed("ln_mod_quant", total,
           lambda: cabi.check(cabi.load().fluxb200_ln_mod_quant_grouped(args, len(items), cabi.fp8_fmt(dtype), D, eps,
                                                                        cabi.stream_ptr()),
                              "fluxb200_ln_mod_quant_grouped"))
    return outs


_grid_barrier_ws = {}


def grid_barrier_ws(device) -> Tensor:
    """The two self-re-arming words of the fused LN -> GEMM launches' grid barrier: one pair per (device, stream) --
    launches that share a pair must be stream-ordered (include/flux_b200.h, fluxb200_f8_gemm_ln)."""
    dev = torch.device(device)
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (index, torch.cuda.current_stream(device).cuda_stream)
    ws = _grid_barrier_ws.get(key)
    if ws is None:
        with torch.cuda.stream(torch.cuda.default_stream(device)):
            ws = torch.zeros(2, dtype=torch.int32, device=device)
        torch.cuda.default_stream(device).synchronize()
        _grid_barrier_ws[key] = ws
    return ws


#: (epilogue, N, K, tuple of M) -> False once the library answered FLUXB200_ERR_UNSUPPORTED for that fused launch
_ln_fused_unsupported = set()
#: LayerNorm-modulate-quantise as a prologue phase of the consuming GEMM launch (fluxb200_f8_gemm_ln: 210 instead of
#: 286 launches per Flux-dev step) or as its own grouped launch.  Measured in the captured c2 step (tools/ablate_step.py,
#: profiles/r2_ablation.md): 40.94 ms fused vs 40.64 ms separate -- programmatic dependent launch already hides the
#: kernel boundary, and inside the GEMM grid the LayerNorm has 20 warps per SM to hide its L2 latency instead of 32.
#: The separate launch is therefore the default; env FLUXB200_FUSE_LN=1 (or setting this flag) selects the fused form.
FUSE_LN_INTO_GEMM = __import__("os").environ.get("FLUXB200_FUSE_LN", "0") not in ("0", "")


def ln_gemm_group(items, dtype: torch.dtype, build_gemms, eps: float = 1e-6):
    """LayerNorm-modulate-quantise of 1-2 row sets + the GEMM(s) that consume them, as ONE launch when the library has a
    fused kernel for the shape (fluxb200_f8_gemm_ln), else as the two grouped launches.  `build_gemms(a8s)` receives
    the fp8 A tensors ([B,L,D] each) and returns the list of deferred GemmArgs (ops.f8_gemm_*(..., defer=list))."""
    args, outs, keep, D, total = _ln_args(items, dtype)
    gs = list(build_gemms(outs))
    key = (gs[0].epilogue, gs[0].N, gs[0].K, tuple(g.M for g in gs))
    same = all(g.N == gs[0].N and g.K == gs[0].K and g.a_fmt == gs[0].a_fmt and g.w_fmt == gs[0].w_fmt and
               g.out_fmt == gs[0].out_fmt for g in gs)
    if FUSE_LN_INTO_GEMM and same and D == 3072 and key not in _ln_fused_unsupported:
        arr = (cabi.GemmArgs * len(gs))(*gs)
        ws = grid_barrier_ws(items[0][0].device)
        done = []

        def launch():
            rc = cabi.load().fluxb200_f8_gemm_ln(arr, len(gs), args, len(items), cabi.fp8_fmt(dtype), D, eps,
                                                 ws.data_ptr(), cabi.stream_ptr())
            if rc == cabi.ERR_UNSUPPORTED:
                return
            cabi.check(rc, "fluxb200_f8_gemm_ln")
            done.append(True)

        _timed("f8_gemm", sum(2.0 * g.M * g.N * g.K for g in gs), launch,
               "ln+" + _gemm_detail(gs) if KERNEL_TIMELINE is not None else "")
        if done:
            return outs
        _ln_fused_unsupported.add(key)
        if KERNEL_TIMELINE is not None and KERNEL_TIMELINE:
            KERNEL_TIMELINE.pop()  # the refused call launched nothing
    _timed("ln_mod_quant", total,
           lambda: cabi.check(cabi.load().fluxb200_ln_mod_quant_grouped(args, len(items), cabi.fp8_fmt(dtype), D, eps,
                                                                        cabi.stream_ptr()),
                              "fluxb200_ln_mod_quant_grouped"))
    if same and len(gs) > 1:
        run_gemm_group(gs)
    else:
        for g in gs:
            run_gemm(g)
    return outs


def qknorm_rope(x: Tensor, norm_w: Optional[Tensor], cos: Optional[Tensor], sin: Optional[Tensor]) -> Tensor:
    """Stand-alone QKNorm + RoPE on [B,H,S,128] (cos/sin: bf16 [Bp,S,64], Bp in {1,B})."""
    cabi.require_cuda(x)
    _want(x, BF16, "qknorm_rope: x"), _want(norm_w, torch.float32, "qknorm_rope: norm_w")
    _want(cos, BF16, "qknorm_rope: cos"), _want(sin, BF16, "qknorm_rope: sin")
    B, H, S, D = x.shape
    if D != 128:
        raise ValueError("head_dim must be 128")
    x = x.contiguous()
    y = torch.empty_like(x)
    bstride = 0
    if cos is not None:
        cos, sin = cos.contiguous(), sin.contiguous()
        bstride = cos.stride(0) if cos.shape[0] > 1 else 0
    cabi.check(cabi.load().fluxb200_qknorm_rope(x.data_ptr(), y.data_ptr(), cabi.ptr(norm_w), cabi.ptr(cos),
                                                cabi.ptr(sin), bstride, B, H, S, 1e-6, cabi.stream_ptr()),
               "fluxb200_qknorm_rope")
    return y


def f8_gemv(a: Tensor, w: Tensor, bias: Optional[Tensor], a_scale_recip: Tensor, w_scale_recip: Tensor) -> Tensor:
    cabi.require_cuda(a, w)
    _want(bias, BF16, "f8_gemv: bias"), _want(a_scale_recip, torch.float32, "f8_gemv: a_scale_recip")
    _want(w_scale_recip, torch.float32, "f8_gemv: w_scale_recip")
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=BF16, device=a.device)
    cabi.check(cabi.load().fluxb200_f8_gemv(a.data_ptr(), cabi.fp8_fmt(a.dtype), w.data_ptr(), cabi.fp8_fmt(w.dtype),
                                            cabi.ptr(bias), a_scale_recip.data_ptr(), w_scale_recip.data_ptr(),
                                            out.data_ptr(), M, N, K, cabi.stream_ptr()), "fluxb200_f8_gemv")
    return out


def gemm_args(a: Tensor, w: Tensor, bias: Optional[Tensor], a_scale_recip: Tensor, w_scale_recip: Tensor,
              epilogue: int) -> cabi.GemmArgs:
    cabi.require_cuda(a, w, bias, a_scale_recip, w_scale_recip)
    if a.dim() != 2 or w.dim() != 2 or a.shape[1] != w.shape[1]:
        raise ValueError(f"f8_gemm: incompatible shapes {tuple(a.shape)} x {tuple(w.shape)}^T")
    _contig(a, "A"), _contig(w, "W")
    _want(bias, BF16, "f8_gemm: bias"), _want(a_scale_recip, torch.float32, "f8_gemm: a_scale_recip")
    _want(w_scale_recip, torch.float32, "f8_gemm: w_scale_recip")
    g = cabi.GemmArgs()
    g.a, g.w, g.bias = a.data_ptr(), w.data_ptr(), cabi.ptr(bias)
    g.a_scale_recip, g.w_scale_recip = a_scale_recip.data_ptr(), w_scale_recip.data_ptr()
    g.M, g.K = a.shape
    g.N = w.shape[0]
    g.a_fmt, g.w_fmt = cabi.fp8_fmt(a.dtype), cabi.fp8_fmt(w.dtype)
    g.epilogue = epilogue
    # deferred (grouped) launches only hold raw pointers: keep the operands alive until the launch has been enqueued
    g._keep = [a, w, bias, a_scale_recip, w_scale_recip]
    return g


def _timed(kind: str, flops: float, launch, detail: str = "") -> None:
    if KERNEL_TIMELINE is None:
        launch()
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    launch()
    e.record()
    KERNEL_TIMELINE.append((kind, flops, s, e, detail))


_EPI_NAMES = {0: "plain", 1: "gate_residual", 2: "gelu_quant", 3: "qkv_rope", 4: "linear1"}


def _gemm_detail(gs) -> str:
    return _EPI_NAMES.get(gs[0].epilogue, "?") + " " + " + ".join(f"{g.M}x{g.N}x{g.K}" for g in gs)


def run_gemm(g: cabi.GemmArgs) -> None:
    _timed("f8_gemm", 2.0 * g.M * g.N * g.K,
           lambda: cabi.check(cabi.load().fluxb200_f8_gemm(C.byref(g), cabi.stream_ptr()), "fluxb200_f8_gemm"),
           _gemm_detail([g]) if KERNEL_TIMELINE is not None else "")


def run_gemm_group(gs) -> None:
    """One persistent launch for up to two GEMM problems sharing N, K, formats and epilogue."""
    gs = list(gs)
    if len(gs) == 1:
        return run_gemm(gs[0])
    arr = (cabi.GemmArgs * len(gs))(*gs)
    _timed("f8_gemm", sum(2.0 * g.M * g.N * g.K for g in gs),
           lambda: cabi.check(cabi.load().fluxb200_f8_gemm_grouped(arr, len(gs), cabi.stream_ptr()),
                              "fluxb200_f8_gemm_grouped"),
           _gemm_detail(gs) if KERNEL_TIMELINE is not None else "")


def f8_gemm(a: Tensor, w: Tensor, bias: Optional[Tensor], a_scale_recip: Tensor, w_scale_recip: Tensor,
            out: Optional[Tensor] = None) -> Tensor:
    """bf16( (A.W^T) * sa * sw + bias ): the torch._scaled_mm call of F8Linear.forward."""
    g = gemm_args(a, w, bias, a_scale_recip, w_scale_recip, cabi.EPI_PLAIN)
    if g.M <= 16 and g.K % 16 == 0 and g.K * 8 <= 200 * 1024:
        return f8_gemv(a, w, bias, a_scale_recip, w_scale_recip)
    if out is None:
        out = torch.empty((g.M, g.N), dtype=BF16, device=a.device)
    g.out, g.ldo = out.data_ptr(), out.stride(0)
    run_gemm(g)
    return out


def f8_gemm_gate_residual(a, w, bias, sa, sw, resid: Tensor, gate: Tensor, rows_per_batch: int,
                          out: Optional[Tensor] = None, defer: Optional[list] = None) -> Tensor:
    """out = resid + gate[b] * linear(a)   (resid [M,N] bf16 view, gate [B,N] view).  With `defer` (a list) the
    launch arguments are appended to it instead of being launched (see run_gemm_group)."""
    g = gemm_args(a, w, bias, sa, sw, cabi.EPI_GATE_RESIDUAL)
    _want(resid, BF16, "f8_gemm_gate_residual: resid"), _want(gate, BF16, "f8_gemm_gate_residual: gate")
    _want(out, BF16, "f8_gemm_gate_residual: out")
    if out is None:
        out = torch.empty((g.M, g.N), dtype=BF16, device=a.device)
    g.out, g.ldo = out.data_ptr(), out.stride(0)
    g.resid, g.ldr = resid.data_ptr(), resid.stride(0)
    g.gate, g.gate_batch_stride = gate.data_ptr(), gate.stride(0)
    g.rows_per_batch = rows_per_batch
    g._keep += [out, resid, gate]
    if defer is not None:
        defer.append(g)
    else:
        run_gemm(g)
    return out


def f8_gemm_gelu_quant(a, w, bias, sa, sw, out_scale: Tensor, out_dtype: torch.dtype, out: Optional[Tensor] = None,
                       out_col_offset: int = 0, defer: Optional[list] = None) -> Tensor:
    g = gemm_args(a, w, bias, sa, sw, cabi.EPI_GELU_QUANT)
    _want(out_scale, torch.float32, "f8_gemm_gelu_quant: out_scale")
    if out is None:
        out = torch.empty((g.M, g.N), dtype=out_dtype, device=a.device)
    g.out, g.ldo = out.data_ptr(), out.stride(0)
    g.out_scale, g.out_fmt, g.out_col_offset = out_scale.data_ptr(), cabi.fp8_fmt(out_dtype), out_col_offset
    g._keep += [out, out_scale]
    if defer is not None:
        defer.append(g)
    else:
        run_gemm(g)
    return out


def f8_gemm_qkv_rope(a, w, bias, sa, sw, q: Tensor, k: Tensor, v: Tensor, q_norm_w: Tensor, k_norm_w: Tensor,
                     cos: Tensor, sin: Tensor, rows_per_batch: int, seq_offset: int,
                     mlp_out: Optional[Tensor] = None, mlp_scale: Optional[Tensor] = None,
                     mlp_col_offset: int = 0, defer: Optional[list] = None) -> None:
    """QKV (or SingleStreamBlock.linear1 when mlp_out is given) GEMM writing normalised+rotated q,k and v
    straight into the joint [B,H,S,128] buffers."""
    epi = cabi.EPI_LINEAR1 if mlp_out is not None else cabi.EPI_QKV_ROPE
    g = gemm_args(a, w, bias, sa, sw, epi)
    for t, what in ((q, "q"), (k, "k"), (v, "v"), (cos, "cos"), (sin, "sin")):
        _want(t, BF16, f"f8_gemm_qkv_rope: {what}")
    _want(q_norm_w, torch.float32, "f8_gemm_qkv_rope: q_norm_w"), _want(k_norm_w, torch.float32, "f8_gemm_qkv_rope: k_norm_w")
    _want(mlp_scale, torch.float32, "f8_gemm_qkv_rope: mlp_scale")
    B, H, S, D = q.shape
    g.q, g.k, g.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    g.num_heads, g.seq_total, g.seq_offset = H, S, seq_offset
    g.rows_per_batch = rows_per_batch
    g.q_norm_w, g.k_norm_w = q_norm_w.data_ptr(), k_norm_w.data_ptr()
    g.rope_cos, g.rope_sin = cos.data_ptr(), sin.data_ptr()
    g.rope_batch_stride = cos.stride(0) if cos.shape[0] > 1 else 0
    if mlp_out is not None:
        g.out, g.ldo = mlp_out.data_ptr(), mlp_out.stride(0)
        g.out_scale, g.out_fmt, g.out_col_offset = mlp_scale.data_ptr(), cabi.fp8_fmt(mlp_out.dtype), mlp_col_offset
    g._keep += [q, k, v, q_norm_w, k_norm_w, cos, sin, mlp_out, mlp_scale]
    if defer is not None:
        defer.append(g)
    else:
        run_gemm(g)


def attention(q: Tensor, k: Tensor, v: Tensor, out: Optional[Tensor] = None, out_scale0: Optional[Tensor] = None,
              out_scale1: Optional[Tensor] = None, split_row: int = 0, variant: int = 0,
              out1: Optional[Tensor] = None) -> Tensor:
    """softmax(q k^T / sqrt(d)) v on [B,H,S,128] -> [B,S,H*128] (bf16, or fp8 when `out` is an fp8 tensor
    and scales are given; `out` may be a column-slice view of a wider buffer).  With `out1`, rows
    [0, split_row) go to `out` and rows [split_row, S) to `out1` (txt / img streams of a double block)."""
    cabi.require_cuda(q, k, v)
    _want(q, BF16, "attention: q"), _want(k, BF16, "attention: k"), _want(v, BF16, "attention: v")
    _want(out_scale0, torch.float32, "attention: out_scale0"), _want(out_scale1, torch.float32, "attention: out_scale1")
    B, H, S, D = q.shape
    if D != 128:
        raise ValueError("head_dim must be 128")
    _contig(q, "q"), _contig(k, "k"), _contig(v, "v")
    if out is None:
        out = torch.empty((B, S, H * D), dtype=BF16, device=q.device)
    else:
        rows = split_row if out1 is not None else S
        if out.dim() != 3 or out.shape[0] != B or out.shape[1] < rows or out.shape[2] != H * D or out.stride(2) != 1:
            raise ValueError(f"attention: out {tuple(out.shape)} (strides {out.stride()}) does not hold [{B}, {rows}, {H * D}] "
                             "rows with unit column stride")
        if out.dtype != BF16 and out_scale0 is None:
            raise ValueError("attention: an fp8 `out` needs out_scale0")
    if out1 is not None and (out1.dim() != 3 or out1.shape[0] != B or out1.shape[1] < S - split_row
                             or out1.shape[2] != H * D or out1.stride(2) != 1):
        raise ValueError(f"attention: out1 {tuple(out1.shape)} does not hold [{B}, {S - split_row}, {H * D}] rows")
    a = cabi.AttentionArgs()
    a.q, a.k, a.v, a.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.ldo, a.out_batch_stride = out.stride(1), out.stride(0)
    a.B, a.H, a.S = B, H, S
    a.softmax_scale = 1.0 / math.sqrt(D)
    a.variant = variant
    if out.dtype == BF16:
        a.out_kind = 0
    else:
        a.out_kind, a.out_fmt = 1, cabi.fp8_fmt(out.dtype)
        a.out_scale0 = out_scale0.data_ptr()
        a.out_scale1 = (out_scale1 if out_scale1 is not None else out_scale0).data_ptr()
    a.split_row = split_row
    if out1 is not None:
        # rows >= split_row go to out1 ([B, S - split_row, H*128], same dtype as out)
        if out1.dtype != out.dtype:
            raise ValueError("attention: out and out1 must share a dtype")
        a.out1, a.ldo1, a.out1_batch_stride = out1.data_ptr(), out1.stride(1), out1.stride(0)
    _timed("attention", 4.0 * B * H * S * S * D,
           lambda: cabi.check(cabi.load().fluxb200_attention(C.byref(a), cabi.stream_ptr()), "fluxb200_attention"))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# VAE decoder ops.  Activations are channels-last bf16 tensors [B, H, W, C] between these calls.
# ---------------------------------------------------------------------------------------------------------------------
def pack_conv_weight(weight: Tensor, cin_pad: Optional[int] = None) -> Tensor:
    """nn.Conv2d weight [N, C, kh, kw] (kh = kw = 1 or 3) -> the kernel's bf16 [N, kh*kw*Cp] with
    w[n, (ky*kw + kx)*Cp + c] = weight[n, c, ky, kx]; Cp = C rounded up to a multiple of 64 (zero filled)."""
    N, Cin, kh, kw = weight.shape
    if (kh, kw) not in ((1, 1), (3, 3)):
        raise ValueError(f"pack_conv_weight: kernel {kh}x{kw} is not supported (1x1 or 3x3)")
    cp = cin_pad or ((Cin + 63) // 64) * 64
    w = torch.zeros((N, kh * kw, cp), dtype=BF16, device=weight.device)
    w[:, :, :Cin] = weight.detach().to(BF16).permute(0, 2, 3, 1).reshape(N, kh * kw, Cin)
    return w.reshape(N, kh * kw * cp)


def conv2d_nhwc(x: Tensor, w_packed: Tensor, bias: Optional[Tensor], taps: int, residual: Optional[Tensor] = None,
                out: Optional[Tensor] = None, out_mode: int = 0, alpha: float = 1.0, nchw_plane: Optional[int] = None,
                gn_stats: Optional[Tensor] = None, stride: int = 1) -> Tensor:
    """fluxb200_conv2d_nhwc.  x bf16 [B, H, W, Cin] (Cin % 64 == 0); w_packed from pack_conv_weight.
    out_mode 0 -> bf16 [B, H, W, N] (+ bias, + residual); 1 -> fp32 [B, H, W, N] = alpha * acc; 2 -> bf16 NCHW [B, N, H, W].
    gn_stats (fp64 [B, 32, 2], out_mode 0): filled with the GroupNorm(32) sums of the stored output by the epilogue.
    stride 2 (3x3 only): the reference's Downsample = F.pad(x, (0, 1, 0, 1)) then stride-2 convolution without padding."""
    cabi.require_cuda(x, w_packed)
    _want(x, BF16, "conv2d_nhwc: x"), _want(w_packed, BF16, "conv2d_nhwc: w"), _want(bias, BF16, "conv2d_nhwc: bias")
    _want(residual, BF16, "conv2d_nhwc: residual")
    if x.dim() != 4 or x.stride(3) != 1:
        raise ValueError("conv2d_nhwc: x must be [B, H, W, C] with contiguous channels")
    B, H, W, Cin = x.shape
    if (H > 1 and x.stride(1) != W * x.stride(2)) or (B > 1 and x.stride(0) != H * W * x.stride(2)):
        raise ValueError("conv2d_nhwc: x must have dense B / H / W strides (a pixel stride >= C is allowed)")
    N = w_packed.shape[0]
    if w_packed.shape[1] != taps * Cin or w_packed.stride(1) != 1:
        raise ValueError(f"conv2d_nhwc: packed weight {tuple(w_packed.shape)} does not match taps={taps} Cin={Cin}")
    if stride not in (1, 2) or (stride == 2 and (taps != 9 or out_mode == 2 or H < 2 or W < 2)):
        raise ValueError("conv2d_nhwc: stride 2 is the 3x3 Downsample convolution (NHWC / fp32 output)")
    Ho, Wo = (H, W) if stride == 1 else ((H - 2) // 2 + 1, (W - 2) // 2 + 1)
    if out is None:
        if out_mode == 0:
            out = torch.empty((B, Ho, Wo, N), dtype=BF16, device=x.device)
        elif out_mode == 1:
            out = torch.empty((B, Ho, Wo, N), dtype=torch.float32, device=x.device)
        else:
            out = torch.empty((B, N, H, W), dtype=BF16, device=x.device)
    _contig(out, "conv2d_nhwc: out")
    _want(out, torch.float32 if out_mode == 1 else BF16, "conv2d_nhwc: out")
    a = cabi.ConvArgs()
    a.x, a.w, a.bias, a.out = x.data_ptr(), w_packed.data_ptr(), cabi.ptr(bias), out.data_ptr()
    a.ldx, a.ldw = x.stride(2), w_packed.stride(0)
    if residual is not None:
        _contig(residual, "conv2d_nhwc: residual")
        if residual.numel() != B * Ho * Wo * N:
            raise ValueError("conv2d_nhwc: residual must have the output's shape")
        a.residual, a.ld_res = residual.data_ptr(), N
    a.ldo = N if out_mode != 2 else (nchw_plane or H * W)
    a.B, a.H, a.W, a.Cin, a.N, a.taps, a.out_mode, a.alpha, a.stride = B, H, W, Cin, N, taps, out_mode, alpha, stride
    if out.numel() != B * Ho * Wo * N and out_mode != 2:
        raise ValueError("conv2d_nhwc: out has the wrong number of elements")
    if gn_stats is not None:
        _want(gn_stats, torch.float64, "conv2d_nhwc: gn_stats")
        if gn_stats.numel() != B * 64 or not gn_stats.is_contiguous():
            raise ValueError("conv2d_nhwc: gn_stats must be a contiguous fp64 [B, 32, 2]")
        a.gn_stats = gn_stats.data_ptr()
    _timed("conv2d", 2.0 * B * Ho * Wo * N * taps * Cin,
           lambda: cabi.check(cabi.load().fluxb200_conv2d_nhwc(C.byref(a), cabi.stream_ptr()), "fluxb200_conv2d_nhwc"),
           f"{B}x{H}x{W} {Cin}->{N} taps {taps}{' stride 2' if stride == 2 else ''}" if KERNEL_TIMELINE is not None else "")
    return out


#: the fused statistics of fluxb200_conv2d_nhwc are kept per CTA for at most this many images
CONV_STATS_MAX_BATCH = 4


def conv_can_fuse_gn_stats(B: int, N: int) -> bool:
    return B <= CONV_STATS_MAX_BATCH and N in (128, 256, 512, 1024)


def group_norm_nhwc(x: Tensor, gamma: Tensor, beta: Tensor, eps: float, swish: bool, out: Optional[Tensor] = None,
                    stats: Optional[Tensor] = None) -> Tensor:
    """nn.GroupNorm(32, C, eps) (+ x * sigmoid(x)) on a channels-last bf16 tensor, fp32 arithmetic, bf16 result.
    stats: the fp64 [B, 32, 2] sums a producing conv2d_nhwc(gn_stats=...) already left behind (skips the reduction pass)."""
    cabi.require_cuda(x, gamma, beta)
    _want(x, BF16, "group_norm_nhwc: x"), _want(gamma, BF16, "group_norm_nhwc: weight"), _want(beta, BF16, "group_norm_nhwc: bias")
    _want(stats, torch.float64, "group_norm_nhwc: stats")
    _contig(x, "group_norm_nhwc: x")
    B, H, W, Cn = x.shape
    if out is None:
        out = torch.empty_like(x)
    ready = stats is not None
    if ready and (stats.numel() != B * 64 or not stats.is_contiguous()):
        raise ValueError("group_norm_nhwc: stats must be a contiguous fp64 [B, 32, 2]")
    ws = stats if ready else torch.empty(64 * B, dtype=torch.float64, device=x.device)
    _timed("group_norm", 0.0,
           lambda: cabi.check(cabi.load().fluxb200_group_norm_nhwc(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(),
                                                                   ws.data_ptr(), int(ready), B, H * W, Cn, eps, int(swish),
                                                                   cabi.stream_ptr()), "fluxb200_group_norm_nhwc"),
           f"{B}x{H}x{W}x{Cn}{' (stats from the conv)' if ready else ''}" if KERNEL_TIMELINE is not None else "")
    return out


def upsample2x_nhwc(x: Tensor) -> Tensor:
    cabi.require_cuda(x)
    _want(x, BF16, "upsample2x_nhwc: x")
    _contig(x, "upsample2x_nhwc: x")
    B, H, W, Cn = x.shape
    out = torch.empty((B, 2 * H, 2 * W, Cn), dtype=BF16, device=x.device)
    _timed("upsample2x", 0.0,
           lambda: cabi.check(cabi.load().fluxb200_upsample2x_nhwc(x.data_ptr(), out.data_ptr(), B, H, W, Cn, cabi.stream_ptr()),
                              "fluxb200_upsample2x_nhwc"))
    return out


def softmax_rows(scores: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """bf16(softmax(scores, -1)) of an fp32 [rows, n] matrix."""
    cabi.require_cuda(scores)
    _want(scores, torch.float32, "softmax_rows: scores")
    if scores.dim() != 2 or scores.stride(1) != 1:
        raise ValueError("softmax_rows: scores must be a 2-D row-major matrix")
    rows, n = scores.shape
    if out is None:
        out = torch.empty((rows, n), dtype=BF16, device=scores.device)
    _timed("softmax_rows", 0.0,
           lambda: cabi.check(cabi.load().fluxb200_softmax_rows(scores.data_ptr(), scores.stride(0), out.data_ptr(), out.stride(0),
                                                                rows, n, cabi.stream_ptr()), "fluxb200_softmax_rows"))
    return out


def vae_latent_prep(z: Tensor, scale_factor: float, shift_factor: float, cpad: int = 64) -> Tensor:
    """z fp32 [B, C, H, W] -> bf16 [B, H, W, cpad] = z / scale_factor + shift_factor, extra channels zero."""
    cabi.require_cuda(z)
    _want(z, torch.float32, "vae_latent_prep: z")
    _contig(z, "vae_latent_prep: z")
    B, Cn, H, W = z.shape
    out = torch.empty((B, H, W, cpad), dtype=BF16, device=z.device)
    _timed("latent_prep", 0.0,
           lambda: cabi.check(cabi.load().fluxb200_vae_latent_prep(z.data_ptr(), out.data_ptr(), B, Cn, H * W, cpad, scale_factor,
                                                                   shift_factor, cabi.stream_ptr()), "fluxb200_vae_latent_prep"))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Text-encoder ops (second half): T5 / CLIP pieces that are not dense layers
# ---------------------------------------------------------------------------------------------------------------------
def dense(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, residual: Optional[Tensor] = None,
          out: Optional[Tensor] = None) -> Tensor:
    """bf16 F.linear on the tcgen05 implicit-GEMM kernel in its dense form: out[M, N] = x[M, K] weight[N, K]^T (+ bias)
    (+ residual), fp32 accumulate, the roundings of conv2d_nhwc's NHWC mode.  K % 64 == 0, N % 8 == 0."""
    if x.dim() != 2 or x.stride(1) != 1 or weight.dim() != 2 or weight.stride(1) != 1 or weight.shape[1] != x.shape[1]:
        raise ValueError(f"dense: x {tuple(x.shape)} / weight {tuple(weight.shape)} must be row-major [M, K] and [N, K]")
    M, K = x.shape
    N = weight.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=BF16, device=x.device)
    xv = x.as_strided((1, 1, M, K), (M * x.stride(0), M * x.stride(0), x.stride(0), 1))
    conv2d_nhwc(xv, weight, bias, 1, residual=None if residual is None else residual.view(1, 1, M, N), out=out.view(1, 1, M, N))
    return out


def rows_norm(x: Tensor, weight: Tensor, bias: Optional[Tensor], eps: float, out: Optional[Tensor] = None) -> Tensor:
    """bias None: T5LayerNorm (RMS); else nn.LayerNorm with affine.  x bf16 [rows, D] (row stride allowed)."""
    cabi.require_cuda(x, weight)
    _want(x, BF16, "rows_norm: x"), _want(weight, BF16, "rows_norm: weight"), _want(bias, BF16, "rows_norm: bias")
    if x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("rows_norm: x must be [rows, D] with contiguous rows")
    rows, D = x.shape
    if out is None:
        out = torch.empty((rows, D), dtype=BF16, device=x.device)
    _timed("rows_norm", 0.0,
           lambda: cabi.check(cabi.load().fluxb200_rows_norm(x.data_ptr(), x.stride(0), weight.data_ptr(), cabi.ptr(bias), out.data_ptr(),
                                                             out.stride(0), rows, D, eps, cabi.stream_ptr()), "fluxb200_rows_norm"))
    return out


def gated_act(x: Tensor, F: int, mode: int, out: Optional[Tensor] = None) -> Tensor:
    """mode 0: bf16(gelu_new(x[:, :F])) * x[:, F:2F] (T5);  mode 1: quick_gelu(x[:, :F]) (CLIP)."""
    cabi.require_cuda(x)
    _want(x, BF16, "gated_act: x")
    if x.dim() != 2 or x.stride(1) != 1 or x.shape[1] < (2 * F if mode == 0 else F):
        raise ValueError("gated_act: x must be [rows, >= F (2 F for the gated form)] with contiguous rows")
    rows = x.shape[0]
    if out is None:
        out = torch.empty((rows, F), dtype=BF16, device=x.device)
    _timed("gated_act", 0.0,
           lambda: cabi.check(cabi.load().fluxb200_gated_act(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rows, F, mode,
                                                             cabi.stream_ptr()), "fluxb200_gated_act"))
    return out


def attention_d64(qkv: Tensor, B: int, S: int, H: int, bias: Optional[Tensor], scale: float, causal: bool,
                  out: Optional[Tensor] = None) -> Tensor:
    """qkv bf16 [B*S, 3*H*64] (q | k | v, head h at columns [64 h, 64 h + 64) of each third) -> bf16 [B*S, H*64]."""
    cabi.require_cuda(qkv)
    _want(qkv, BF16, "attention_d64: qkv"), _want(bias, BF16, "attention_d64: bias")
    if qkv.dim() != 2 or qkv.shape != (B * S, 3 * H * 64) or qkv.stride(1) != 1:
        raise ValueError(f"attention_d64: qkv must be [{B * S}, {3 * H * 64}], got {tuple(qkv.shape)}")
    if bias is not None and (tuple(bias.shape) != (H, S, S) or not bias.is_contiguous()):
        raise ValueError("attention_d64: bias must be a contiguous [H, S, S]")
    if out is None:
        out = torch.empty((B * S, H * 64), dtype=BF16, device=qkv.device)
    es = qkv.element_size()
    q, k, v = qkv.data_ptr(), qkv.data_ptr() + H * 64 * es, qkv.data_ptr() + 2 * H * 64 * es
    _timed("attention_d64", 4.0 * B * H * S * S * 64,
           lambda: cabi.check(cabi.load().fluxb200_attention_d64(q, k, v, qkv.stride(0), cabi.ptr(bias), out.data_ptr(), out.stride(0),
                                                                 B, H, S, scale, int(causal), cabi.stream_ptr()), "fluxb200_attention_d64"))
    return out
