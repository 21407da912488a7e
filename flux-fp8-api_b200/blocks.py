"""Block layer of the Flux DiT on B200: attention / rope / QKNorm / Modulation / DoubleStreamBlock /
SingleStreamBlock with the reference's signatures (reference: modules/flux_model.py:41-92, 158-485).

Two execution modes per block, both entirely on the GPU through libflux_b200.so:

* fused (steady state: every linear is an F8Linear with a frozen input scale).  A DoubleStreamBlock is
  2 x (LN-modulate-quantise, QKV GEMM with QK-RMSNorm+RoPE epilogue) + one attention kernel (fp8 output)
  + 2 x (proj GEMM with gate-residual epilogue, LN-modulate-quantise, MLP-up GEMM with GELU+quantise
  epilogue, MLP-down GEMM with gate-residual epilogue).  A SingleStreamBlock is 4 launches.  No bf16
  intermediate other than the residual stream, q, k and v ever reaches HBM.
* eager (calibration calls, or un-quantised bf16 linears): the reference's op order, each op one kernel,
  so F8Linear.quantize_input sees exactly the tensors the reference would calibrate on.
"""
from __future__ import annotations

import threading
import weakref
from collections import namedtuple
from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from . import _cabi as cabi
from . import ops
from .f8linear import F8Linear

BF16 = torch.bfloat16
HEAD_DIM = 128

#: diagnostics: when set to a dict, the fused block paths record the fp8 GEMM operands they produce under the name of
#: the consuming layer ("img_attn.qkv", "linear2", ...), so a test can compare them byte for byte with the reference's
#: own quantised inputs (the e5m2/e4m3 "code flip" rate, tests/test_gpu_reference.py).  None on the product path.
TAP = None


def _tap(name: str, t: Tensor) -> None:
    if TAP is not None:
        TAP[name] = t


# ------------------------------------------------------------------------------------------------
# RoPE tables
# ------------------------------------------------------------------------------------------------
def rope(pos: Tensor, dim: int, theta: int) -> Tensor:
    """[..., dim/2, 2, 2] rotation table in fp32 (reference modules/flux_model.py:49-57).  Step-invariant
    and tiny (S x 64 angles), so it stays in torch and is computed once per request by EmbedND."""
    omega = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32, device=pos.device) / dim))
    ang = pos.float().unsqueeze(-1) * omega
    c, s = torch.cos(ang), torch.sin(ang)
    return torch.stack([c, -s, s, c], dim=-1).reshape(*ang.shape, 2, 2)


class EmbedND(nn.Module):
    def __init__(self, dim: int, theta: int, axes_dim: list, dtype: torch.dtype = torch.bfloat16):
        super().__init__()
        self.dim, self.theta, self.axes_dim, self.dtype = dim, theta, axes_dim, dtype

    def forward(self, ids: Tensor) -> Tensor:
        tables = [rope(ids[..., i], self.axes_dim[i], self.theta).type(self.dtype) for i in range(ids.shape[-1])]
        return torch.cat(tables, dim=-3).unsqueeze(1)


class _RopeCache(threading.local):
    """Last (pe -> cos, sin) extraction of this thread.  The entry holds `pe` by WEAK reference and is a hit only
    for the very same tensor object: a new `pe` allocated at a recycled address (same shape, and version 0 under
    inference_mode) can never alias it -- the round-1 key (data_ptr, shape, version) could."""

    def __init__(self):
        self.entry = None


_rope_cache = _RopeCache()


def tensor_version(t: Tensor) -> int:
    """In-place version counter, or 0 for inference tensors (which do not track one)."""
    try:
        return t._version
    except RuntimeError:
        return 0


def extract_cos_sin(pe: Tensor) -> Tuple[Tensor, Tensor]:
    """pe [B,1,S,64,2,2] -> (cos, sin) as contiguous bf16 [B,S,64]: the two table entries the kernels need
    (pe[...,0,0] = cos, pe[...,1,0] = sin; the other two are their negation / copy)."""
    if pe.dtype != BF16:
        raise ValueError(f"pe must be bfloat16 (EmbedND output), got {pe.dtype}")
    return pe[:, 0, :, :, 0, 0].contiguous(), pe[:, 0, :, :, 1, 0].contiguous()


def rope_cos_sin(pe: Tensor) -> Tuple[Tensor, Tensor]:
    """(cos, sin) of a `pe` table for callers that only hand over `pe` (the reference's block signature).  Flux.forward
    computes the pair once next to `pe` and passes it down (`rope=`), so the steady-state path never comes here."""
    hit = _rope_cache.entry
    if hit is not None and hit[0]() is pe and hit[1] == tensor_version(pe):
        return hit[2], hit[3]
    cos, sin = extract_cos_sin(pe)
    _rope_cache.entry = (weakref.ref(pe), tensor_version(pe), cos, sin)
    return cos, sin


def apply_rope(xq: Tensor, xk: Tensor, freqs_cis: Tensor) -> Tuple[Tensor, Tensor]:
    cos, sin = rope_cos_sin(freqs_cis)
    return ops.qknorm_rope(xq, None, cos, sin), ops.qknorm_rope(xk, None, cos, sin)


def attention(q: Tensor, k: Tensor, v: Tensor, pe: Tensor) -> Tensor:
    """RoPE on q,k then softmax(q k^T / sqrt(d)) v, returned as [B, S, H*d]."""
    q, k = apply_rope(q, k, pe)
    return ops.attention(q, k, v.contiguous())


# ------------------------------------------------------------------------------------------------
# norms
# ------------------------------------------------------------------------------------------------
class RMSNorm(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.scale = nn.Parameter(torch.ones(dim))
        self._w32: Optional[Tensor] = None
        self._w32_key = None

    def weight_fp32(self) -> Tensor:
        """The learned scale as the fp32 vector the kernels read.  Cached against the identity of the Parameter
        object, its storage and (where tracked) its in-place version; load_state_dict / .to() / invalidate_derived()
        drop it.  An in-place write into an *inference-mode* parameter is the one change this cannot see: call
        invalidate_derived(model) after such a write (parallel.broadcast_state does)."""
        w = self.scale
        key = (id(w), w.data_ptr(), w.device, tensor_version(w))
        if self._w32 is None or self._w32_key != key:
            self._w32 = w.detach().float().contiguous()
            self._w32_key = key
        return self._w32

    def invalidate_derived(self) -> None:
        self._w32, self._w32_key = None, None

    def _apply(self, fn, *a, **kw):
        self.invalidate_derived()
        return super()._apply(fn, *a, **kw)

    def _load_from_state_dict(self, *a, **kw):
        self.invalidate_derived()
        return super()._load_from_state_dict(*a, **kw)

    def forward(self, x: Tensor):
        return ops.qknorm_rope(x, self.weight_fp32(), None, None)


def invalidate_derived(model: nn.Module) -> None:
    """Drop every value cached from a parameter / buffer (RMSNorm fp32 weights, F8Linear quantising scales, batched
    modulation tables) after buffers were rewritten IN PLACE (e.g. parallel.broadcast_state's copy_)."""
    for m in model.modules():
        if isinstance(m, RMSNorm):
            m.invalidate_derived()
        elif isinstance(m, F8Linear):
            m.__dict__.pop("_qscale_cache", None)
    model.__dict__.pop("_mod_bank", None)


class QKNorm(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.query_norm = RMSNorm(dim)
        self.key_norm = RMSNorm(dim)

    def forward(self, q: Tensor, k: Tensor, v: Tensor) -> Tuple[Tensor, Tensor]:
        return self.query_norm(q), self.key_norm(k)


def _make_linear(in_f: int, out_f: int, bias: bool, quantized: bool) -> nn.Module:
    if quantized:
        return F8Linear(in_features=in_f, out_features=out_f, bias=bias)
    return nn.Linear(in_f, out_f, bias=bias)


def _split_heads(x: Tensor, heads: int) -> Tuple[Tensor, Tensor, Tensor]:
    B, L, D3 = x.shape
    q, k, v = x.reshape(B, L, 3, heads, D3 // (3 * heads)).permute(2, 0, 3, 1, 4)
    return q, k, v


class SelfAttention(nn.Module):
    """Container of qkv / norm / proj (the blocks drive its children directly, as in the reference)."""

    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = False, prequantized: bool = False):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = _make_linear(dim, dim * 3, qkv_bias, prequantized)
        self.norm = QKNorm(dim // num_heads)
        self.proj = _make_linear(dim, dim, True, prequantized)
        self.K, self.H = 3, num_heads
        self.KH = self.K * self.H

    def rearrange_for_norm(self, x: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        return _split_heads(x, self.H)

    def forward(self, x: Tensor, pe: Tensor) -> Tensor:
        q, k, v = self.rearrange_for_norm(self.qkv(x))
        q, k = self.norm(q, k, v)
        return self.proj(attention(q, k, v, pe=pe))


# ------------------------------------------------------------------------------------------------
# modulation
# ------------------------------------------------------------------------------------------------
ModulationOut = namedtuple("ModulationOut", ["shift", "scale", "gate"])


class Modulation(nn.Module):
    def __init__(self, dim: int, double: bool, quantized_modulation: bool = False):
        super().__init__()
        self.is_double = double
        self.multiplier = 6 if double else 3
        self.lin = _make_linear(dim, self.multiplier * dim, True, quantized_modulation)
        self.act = nn.SiLU()

    def project(self, vec: Tensor) -> Tensor:
        """lin(silu(vec)) as [B, multiplier*dim]."""
        lin = self.lin
        if isinstance(lin, F8Linear) and lin.frozen:
            vq, _ = ops.silu_quant(vec, lin.qscale, lin.input_float8_dtype)
            return ops.f8_gemm(vq, lin.float8_data, lin.bias, lin.input_scale_reciprocal, lin.scale_reciprocal)
        _, act = ops.silu_quant(vec, None, None, want_bf16=True)
        return lin(act)

    def forward(self, vec: Tensor):
        out = self.project(vec)[:, None, :].chunk(self.multiplier, dim=-1)
        return ModulationOut(*out[:3]), (ModulationOut(*out[3:]) if self.is_double else None)


class ModulationBank:
    """Every Modulation.lin of a model as ONE batched launch: all of them consume the same `vec`, so a step needs one
    weight-streaming GEMV over the concatenated rows instead of 76 x (silu [, quantise], GEMV) launches.

    * F8Linear modulation (frozen): fluxb200_modulation_batched -- one SiLU+quantise pass per layer scale + the GEMV
      over 3.2 GB of e4m3 per step for Flux-dev.
    * nn.Linear modulation (quantize_modulation=False, BASELINE config c5): fluxb200_modulation_batched_bf16 -- the
      same GEMV over 6.5 GB of bf16 weights, SiLU fused into the staging of `vec`.
    Results are views into one [B, sum(N)] buffer, chunked exactly like Modulation.forward."""

    COLS_PER_BLOCK = 64
    MAX_BATCH = 16

    def __init__(self, mods):
        self.mods = list(mods)
        lins = [m.lin for m in self.mods]
        if not lins:
            raise ValueError("ModulationBank needs at least one Modulation")
        self.f8 = all(isinstance(l, F8Linear) for l in lins)
        self.K = lins[0].in_features
        if self.f8:
            if not _frozen(*lins):
                raise ValueError("ModulationBank needs frozen F8Linear modulation layers")
            self.in_dtype = lins[0].input_float8_dtype
            if any(l.in_features != self.K or l.input_float8_dtype != self.in_dtype or
                   l.float8_dtype != torch.float8_e4m3fn for l in lins):
                raise ValueError("ModulationBank: heterogeneous modulation layers")
            if self.K % 16 or self.K > 4096:
                raise ValueError(f"ModulationBank: K={self.K} not supported by the batched kernel")
        else:
            if any(isinstance(l, F8Linear) or not isinstance(l, nn.Linear) for l in lins):
                raise ValueError("ModulationBank: a mix of F8Linear and nn.Linear modulation layers")
            if any(l.in_features != self.K or l.weight.dtype != BF16 or not l.weight.is_cuda or
                   not l.weight.is_contiguous() or (l.bias is not None and l.bias.dtype != BF16) for l in lins):
                raise ValueError("ModulationBank: bf16 modulation needs contiguous bfloat16 CUDA weights of one width")
            if self.K % 32 or self.K > 4096:
                raise ValueError(f"ModulationBank: K={self.K} not supported by the batched bf16 kernel")
        table = (cabi.GemvLayer * len(lins))()
        self._keep = []
        off = blocks = 0
        self.offsets = []
        for i, l in enumerate(lins):
            if self.f8:
                q = l.qscale
                self._keep += [l.float8_data, l.bias, q, l.input_scale_reciprocal, l.scale_reciprocal]
                table[i].w, table[i].in_qscale = l.float8_data.data_ptr(), q.data_ptr()
                table[i].a_scale_recip = l.input_scale_reciprocal.data_ptr()
                table[i].w_scale_recip = l.scale_reciprocal.data_ptr()
            else:
                self._keep += [l.weight, l.bias]
                table[i].w = l.weight.data_ptr()
            table[i].bias = cabi.ptr(l.bias)
            table[i].N, table[i].out_offset, table[i].block_start = l.out_features, off, blocks
            self.offsets.append(off)
            off += l.out_features
            blocks += (l.out_features + self.COLS_PER_BLOCK - 1) // self.COLS_PER_BLOCK
        self.total_n, self.total_blocks = off, blocks
        dev = (lins[0].float8_data if self.f8 else lins[0].weight).device
        self.table = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).to(dev)
        self.signature = self._signature()

    def _signature(self):
        if self.f8:
            return tuple((m.lin.float8_data.data_ptr(), m.lin.input_scale.data_ptr(), m.lin.scale_reciprocal.data_ptr(),
                          id(m.lin.input_scale)) for m in self.mods)
        return tuple((id(m.lin.weight), m.lin.weight.data_ptr(), cabi.ptr(m.lin.bias)) for m in self.mods)

    def stale(self) -> bool:
        """True when a layer's buffers were replaced (e.g. LoRA fuse through set_weight_tensor)."""
        try:
            return self._signature() != self.signature
        except AttributeError:  # a layer changed type under us
            return True

    def __call__(self, vec: Tensor):
        B = vec.shape[0]
        if B > self.MAX_BATCH:
            raise ValueError(f"ModulationBank: batch {B} > {self.MAX_BATCH} (use the per-block Modulation path)")
        vec = vec.contiguous()
        out = torch.empty((B, self.total_n), dtype=BF16, device=vec.device)
        if self.f8:
            aq = torch.empty((len(self.mods), B, self.K), dtype=torch.uint8, device=vec.device)
            ops._timed("modulation_batched", float(self.total_n) * self.K,
                       lambda: cabi.check(cabi.load().fluxb200_modulation_batched(
                           vec.data_ptr(), self.table.data_ptr(), len(self.mods), self.total_blocks, aq.data_ptr(),
                           out.data_ptr(), out.stride(0), B, self.K, cabi.fp8_fmt(self.in_dtype), cabi.E4M3,
                           cabi.stream_ptr()), "fluxb200_modulation_batched"))
        else:
            ops._timed("modulation_batched", 2.0 * self.total_n * self.K,
                       lambda: cabi.check(cabi.load().fluxb200_modulation_batched_bf16(
                           vec.data_ptr(), self.table.data_ptr(), len(self.mods), self.total_blocks, out.data_ptr(),
                           out.stride(0), B, self.K, cabi.stream_ptr()), "fluxb200_modulation_batched_bf16"))
        res = []
        for m, off in zip(self.mods, self.offsets):
            chunks = out[:, None, off:off + m.lin.out_features].chunk(m.multiplier, dim=-1)
            res.append((ModulationOut(*chunks[:3]), ModulationOut(*chunks[3:]) if m.is_double else None))
        return res


def _frozen(*lins: nn.Module) -> bool:
    return all(isinstance(l, F8Linear) and l.frozen for l in lins)


def _gate2d(gate: Tensor) -> Tensor:
    """[B,1,D] view into the modulation output -> [B,D] view (same storage, sample stride preserved)."""
    return gate.reshape(gate.shape[0], gate.shape[-1])


# ------------------------------------------------------------------------------------------------
# DoubleStreamBlock
# ------------------------------------------------------------------------------------------------
class DoubleStreamBlock(nn.Module):
    #: launch each txt/img GEMM pair as one grouped kernel (False: two launches, for A/B measurements)
    group_streams = True

    def __init__(self, hidden_size: int, num_heads: int, mlp_ratio: float, qkv_bias: bool = False,
                 dtype: torch.dtype = torch.float16, quantized_modulation: bool = False, prequantized: bool = False):
        super().__init__()
        self.dtype = dtype
        mlp_hidden = int(hidden_size * mlp_ratio)
        self.num_heads, self.hidden_size = num_heads, hidden_size
        for s in ("img", "txt"):
            setattr(self, f"{s}_mod", Modulation(hidden_size, double=True, quantized_modulation=quantized_modulation))
            setattr(self, f"{s}_norm1", nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6))
            setattr(self, f"{s}_attn", SelfAttention(dim=hidden_size, num_heads=num_heads, qkv_bias=qkv_bias,
                                                     prequantized=prequantized))
            setattr(self, f"{s}_norm2", nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6))
            setattr(self, f"{s}_mlp", nn.Sequential(_make_linear(hidden_size, mlp_hidden, True, prequantized),
                                                    nn.GELU(approximate="tanh"),
                                                    _make_linear(mlp_hidden, hidden_size, True, prequantized)))
        self.K, self.H = 3, num_heads
        self.KH = self.K * self.H
        self.do_clamp = dtype == torch.float16

    def rearrange_for_norm(self, x: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        return _split_heads(x, self.H)

    # -- helpers ---------------------------------------------------------------------------------
    def _fusable(self, img: Tensor, txt: Tensor) -> bool:
        lins = []
        for s in ("img", "txt"):
            attn, mlp = getattr(self, f"{s}_attn"), getattr(self, f"{s}_mlp")
            lins += [attn.qkv, attn.proj, mlp[0], mlp[2]]
        return (_frozen(*lins) and img.dtype == BF16 and txt.dtype == BF16 and self.hidden_size % 256 == 0
                and self.hidden_size <= 4096 and self.hidden_size // self.num_heads == HEAD_DIM)

    def forward(self, img: Tensor, txt: Tensor, vec: Tensor, pe: Tensor, mods=None, rope=None) -> Tuple[Tensor, Tensor]:
        """Reference signature (img, txt, vec, pe).  `mods` (this block's precomputed modulation outputs) and `rope`
        (the (cos, sin) pair extracted from `pe`) are optional hand-downs from Flux.forward."""
        cabi.require_cuda(img, txt, vec, pe)
        if mods is None:
            mods = (self.img_mod(vec), self.txt_mod(vec))
        if self._fusable(img, txt):
            return self._forward_fused(img, txt, pe, mods, rope)
        return self._forward_eager(img, txt, pe, mods)

    def _forward_fused(self, img, txt, pe, mods, rope=None):
        (img_mod1, img_mod2), (txt_mod1, txt_mod2) = mods
        B, L, D = img.shape
        T = txt.shape[1]
        S, H = L + T, self.num_heads
        img, txt = img.contiguous(), txt.contiguous()
        cos, sin = rope if rope is not None else rope_cos_sin(pe)
        dev = img.device
        q = torch.empty((B, H, S, HEAD_DIM), dtype=BF16, device=dev)
        k, v = torch.empty_like(q), torch.empty_like(q)

        # The txt and img streams use different weights on different row counts; each pair of GEMMs goes out as ONE
        # grouped launch (the txt problem alone, M = 512 per sample, would leave most SMs idle).
        streams = ((txt, txt_mod1, txt_mod2, self.txt_attn, self.txt_mlp, T, 0),
                   (img, img_mod1, img_mod2, self.img_attn, self.img_mlp, L, T))
        def qkv_gemms(a8s):
            group = []
            for (x, mod1, _, attn, _, rows, off), a8 in zip(streams, a8s):
                lin = attn.qkv
                ops.f8_gemm_qkv_rope(a8.view(-1, D), lin.float8_data, lin.bias, lin.input_scale_reciprocal,
                                     lin.scale_reciprocal, q, k, v, attn.norm.query_norm.weight_fp32(),
                                     attn.norm.key_norm.weight_fp32(), cos, sin, rows_per_batch=rows, seq_offset=off,
                                     defer=group)
            return group

        # LN1 -> modulate -> quantise -> QKV GEMM (+ QK-RMSNorm + RoPE epilogue) of both streams: one launch
        a8s = self._ln_then_gemms([(x, mod1, attn.qkv) for x, mod1, _, attn, _, _, _ in streams], qkv_gemms)
        _tap("txt_attn.qkv", a8s[0]), _tap("img_attn.qkv", a8s[1])

        tp, ip = self.txt_attn.proj, self.img_attn.proj
        if tp.input_float8_dtype != ip.input_float8_dtype:
            raise ValueError("txt_attn.proj and img_attn.proj must share an input float8 dtype")
        txt_a8 = torch.empty((B, T, D), dtype=tp.input_float8_dtype, device=dev)
        img_a8 = torch.empty((B, L, D), dtype=ip.input_float8_dtype, device=dev)
        ops.attention(q, k, v, out=txt_a8, out_scale0=tp.qscale, out_scale1=ip.qscale, split_row=T,
                      out1=img_a8)

        _tap("txt_attn.proj", txt_a8), _tap("img_attn.proj", img_a8)
        # x = x + gate1 * proj(attn)
        ys, group = [], []
        for (x, mod1, _, attn, _, rows, _), a8 in zip(streams, (txt_a8, img_a8)):
            proj = attn.proj
            ys.append(ops.f8_gemm_gate_residual(a8.view(-1, D), proj.float8_data, proj.bias, proj.input_scale_reciprocal,
                                                proj.scale_reciprocal, x.view(-1, D), _gate2d(mod1.gate), rows,
                                                defer=group))
        self._launch(group)
        # x = x + gate2 * mlp((1 + scale2) * LN(x) + shift2): LN2 -> modulate -> quantise -> MLP-up GEMM (GELU + quantise
        # epilogue) of both streams in one launch, then the MLP-down GEMMs
        hs = []

        def up_gemms(m8s):
            group = []
            for (x, _, mod2, _, mlp, rows, _), m8 in zip(streams, m8s):
                up, down = mlp[0], mlp[2]
                hs.append(ops.f8_gemm_gelu_quant(m8.view(-1, D), up.float8_data, up.bias, up.input_scale_reciprocal,
                                                 up.scale_reciprocal, down.qscale, down.input_float8_dtype, defer=group))
            return group

        m8s = self._ln_then_gemms([(y.view(B, rows, D), mod2, mlp[0])
                                   for (_, _, mod2, _, mlp, rows, _), y in zip(streams, ys)], up_gemms)
        _tap("txt_mlp.0", m8s[0]), _tap("img_mlp.0", m8s[1]), _tap("txt_mlp.2", hs[0]), _tap("img_mlp.2", hs[1])
        group = []
        for (x, _, mod2, _, mlp, rows, _), y, h8 in zip(streams, ys, hs):
            down = mlp[2]
            ops.f8_gemm_gate_residual(h8, down.float8_data, down.bias, down.input_scale_reciprocal,
                                      down.scale_reciprocal, y, _gate2d(mod2.gate), rows, out=y, defer=group)
        self._launch(group)
        txt_out, img_out = ys[0].view(B, T, D), ys[1].view(B, L, D)
        return img_out, txt_out

    @staticmethod
    def _ln_then_gemms(items, build_gemms):
        """LN -> modulate -> quantise of the txt and img rows followed by the GEMMs that consume them.  When both
        consumers take the same fp8 input format everything goes out as ONE launch (ops.ln_gemm_group: the LayerNorm
        is a prologue phase of the persistent GEMM grid); else one LN launch per stream, then the GEMMs."""
        (x0, m0, l0), (x1, m1, l1) = items
        if l0.input_float8_dtype == l1.input_float8_dtype:
            return ops.ln_gemm_group([(x0, m0.shift, m0.scale, l0.qscale), (x1, m1.shift, m1.scale, l1.qscale)],
                                     l0.input_float8_dtype, build_gemms)
        a8s = [ops.ln_mod_quant(x, m.shift, m.scale, l.qscale, l.input_float8_dtype)[0] for x, m, l in items]
        DoubleStreamBlock._launch(build_gemms(a8s))
        return a8s

    @staticmethod
    def _ln_pair(items):
        """LN -> modulate -> quantise for the txt and img streams: one grouped launch when both consumers take the same
        fp8 input format, else one launch each.  items: [(x, ModulationOut, consuming F8Linear)] * 2."""
        (x0, m0, l0), (x1, m1, l1) = items
        if l0.input_float8_dtype == l1.input_float8_dtype:
            return ops.ln_mod_quant_pair([(x0, m0.shift, m0.scale, l0.qscale), (x1, m1.shift, m1.scale, l1.qscale)],
                                         l0.input_float8_dtype)
        return [ops.ln_mod_quant(x, m.shift, m.scale, l.qscale, l.input_float8_dtype)[0] for x, m, l in items]

    @staticmethod
    def _launch(group):
        """Launch the txt/img pair grouped when the two problems are compatible, else one by one."""
        a, b = group
        same = (a.N == b.N and a.K == b.K and a.a_fmt == b.a_fmt and a.w_fmt == b.w_fmt and a.out_fmt == b.out_fmt)
        if same and DoubleStreamBlock.group_streams:
            ops.run_gemm_group(group)
        else:
            ops.run_gemm(a)
            ops.run_gemm(b)

    def _forward_eager(self, img, txt, pe, mods):
        (img_mod1, img_mod2), (txt_mod1, txt_mod2) = mods

        def qkv_of(x, mod, attn):
            _, xm = ops.ln_mod_quant(x, mod.shift, mod.scale, None, None, want_bf16=True)
            q, k, v = _split_heads(attn.qkv(xm), self.H)
            q, k = attn.norm(q, k, v)
            return q, k, v

        iq, ik, iv = qkv_of(img, img_mod1, self.img_attn)
        tq, tk, tv = qkv_of(txt, txt_mod1, self.txt_attn)
        att = attention(torch.cat((tq, iq), dim=2), torch.cat((tk, ik), dim=2), torch.cat((tv, iv), dim=2), pe=pe)
        T = txt.shape[1]
        txt_attn, img_attn = att[:, :T], att[:, T:]

        def tail(x, a, mod1, mod2, attn, mlp):
            x = x + mod1.gate * attn.proj(a.contiguous())
            _, xm = ops.ln_mod_quant(x, mod2.shift, mod2.scale, None, None, want_bf16=True)
            return x + mod2.gate * mlp(xm)

        img = tail(img, img_attn, img_mod1, img_mod2, self.img_attn, self.img_mlp)
        txt = tail(txt, txt_attn, txt_mod1, txt_mod2, self.txt_attn, self.txt_mlp)
        return img, txt


# ------------------------------------------------------------------------------------------------
# SingleStreamBlock
# ------------------------------------------------------------------------------------------------
class SingleStreamBlock(nn.Module):
    """DiT block with parallel attention / MLP branches sharing linear1 and linear2."""

    def __init__(self, hidden_size: int, num_heads: int, mlp_ratio: float = 4.0, qk_scale: Optional[float] = None,
                 dtype: torch.dtype = torch.float16, quantized_modulation: bool = False, prequantized: bool = False):
        super().__init__()
        self.dtype = dtype
        self.hidden_dim = self.hidden_size = hidden_size
        self.num_heads = num_heads
        head_dim = hidden_size // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.mlp_hidden_dim = int(hidden_size * mlp_ratio)
        self.linear1 = _make_linear(hidden_size, hidden_size * 3 + self.mlp_hidden_dim, True, prequantized)
        self.linear2 = _make_linear(hidden_size + self.mlp_hidden_dim, hidden_size, True, prequantized)
        self.norm = QKNorm(head_dim)
        self.pre_norm = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.mlp_act = nn.GELU(approximate="tanh")
        self.modulation = Modulation(hidden_size, double=False,
                                     quantized_modulation=quantized_modulation and prequantized)
        self.K, self.H = 3, num_heads
        self.KH = self.K * self.H
        self.do_clamp = dtype == torch.float16

    def _fusable(self, x: Tensor) -> bool:
        return (_frozen(self.linear1, self.linear2) and x.dtype == BF16 and self.hidden_size % 256 == 0
                and self.hidden_size <= 4096 and self.hidden_size // self.num_heads == HEAD_DIM
                and self.mlp_hidden_dim % 128 == 0
                and self.linear1.input_float8_dtype == self.linear2.input_float8_dtype)

    def forward(self, x: Tensor, vec: Tensor, pe: Tensor, mod=None, rope=None) -> Tensor:
        cabi.require_cuda(x, vec, pe)
        if mod is None:
            mod = self.modulation(vec)[0]
        if self._fusable(x):
            return self._forward_fused(x, pe, mod, rope)
        return self._forward_eager(x, pe, mod)

    def _forward_fused(self, x, pe, mod, rope=None):
        B, S, D = x.shape
        H, l1, l2 = self.num_heads, self.linear1, self.linear2
        x = x.contiguous()
        cos, sin = rope if rope is not None else rope_cos_sin(pe)
        dev = x.device
        q = torch.empty((B, H, S, HEAD_DIM), dtype=BF16, device=dev)
        k, v = torch.empty_like(q), torch.empty_like(q)
        # linear2's input = [attention | gelu(mlp)] is assembled in fp8 by the two producers
        cat8 = torch.empty((B, S, D + self.mlp_hidden_dim), dtype=l2.input_float8_dtype, device=dev)

        def linear1_gemm(a8s):
            group = []
            ops.f8_gemm_qkv_rope(a8s[0].view(-1, D), l1.float8_data, l1.bias, l1.input_scale_reciprocal,
                                 l1.scale_reciprocal, q, k, v, self.norm.query_norm.weight_fp32(),
                                 self.norm.key_norm.weight_fp32(), cos, sin, rows_per_batch=S, seq_offset=0,
                                 mlp_out=cat8.view(B * S, -1), mlp_scale=l2.qscale, mlp_col_offset=D, defer=group)
            return group

        # pre_norm -> modulate -> quantise -> linear1 (QKV + RMSNorm + RoPE | GELU + quantise epilogues): one launch
        (a8,) = ops.ln_gemm_group([(x, mod.shift, mod.scale, l1.qscale)], l1.input_float8_dtype, linear1_gemm)
        ops.attention(q, k, v, out=cat8[..., :D], out_scale0=l2.qscale, split_row=0)
        _tap("linear1", a8), _tap("linear2", cat8)
        out = ops.f8_gemm_gate_residual(cat8.view(B * S, -1), l2.float8_data, l2.bias, l2.input_scale_reciprocal,
                                        l2.scale_reciprocal, x.view(-1, D), _gate2d(mod.gate), S)
        return out.view(B, S, D)

    def _forward_eager(self, x, pe, mod):
        _, x_mod = ops.ln_mod_quant(x, mod.shift, mod.scale, None, None, want_bf16=True)
        qkv, mlp = torch.split(self.linear1(x_mod), [3 * self.hidden_size, self.mlp_hidden_dim], dim=-1)
        q, k, v = _split_heads(qkv, self.H)
        q, k = self.norm(q, k, v)
        attn = attention(q, k, v, pe=pe)
        output = self.linear2(torch.cat((attn, self.mlp_act(mlp)), 2))
        return x + mod.gate * output
