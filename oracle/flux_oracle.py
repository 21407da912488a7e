"""CPU oracle for the FP8 Flux-DiT denoise hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch restatement of what aredden/flux-fp8-api computes on the path SURVEY.md section 8
names (F8Linear, Modulation, QKNorm, RoPE, attention, Double/SingleStreamBlock, Flux.forward).  It
is written as pure functions over a flat state-dict (the reference's own key names) so that it
shares no structure with the reference's nn.Module code; every function cites the reference lines
it follows (paths relative to /root/reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker or the reported CPU baseline.  Nothing under
flux-fp8-api_b200/ imports it; the product path fails loudly without its CUDA library.

Pinning: the reference ships no tests or golden vectors for this path.  The
oracle is therefore pinned against the reference itself, imported unmodified in the authoring
container: oracle/make_golden.py runs reference and oracle on the same seeded inputs, asserts they
agree, and commits small reference outputs under tests/golden/ which tests/test_oracle_golden.py
re-checks on every run (no /root/reference needed at test time).

Numerics honoured: bf16 rounding after every eager op, fp32 accumulation
inside matmuls, quantisation as fp8(clamp(bf16(x*s))), scale clamp in amax_to_scale.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
E4M3 = torch.float8_e4m3fn
E5M2 = torch.float8_e5m2


# --------------------------------------------------------------------------------------------
# F8Linear arithmetic (float8_quantize.py)
# --------------------------------------------------------------------------------------------
def fp8_max(dtype: torch.dtype) -> float:
    return torch.finfo(dtype).max


def amax_to_scale(amax: Tensor, max_val: float) -> Tensor:
    """float8_quantize.py:214-215 -- note the clamp of the *scale* to max_val."""
    return (max_val / torch.clamp(amax, min=1e-12)).clamp(max=max_val)


def quantize(x: Tensor, scale: Tensor, dtype: torch.dtype) -> Tensor:
    """float8_quantize.py:217-218 + .to(fp8) at :200-202 / :274-276.
    x bf16, scale 0-dim fp32: the product is rounded to bf16 (type promotion), clamped, then cast."""
    m = fp8_max(dtype)
    return (x * scale).clamp(-m, m).to(dtype)


def quantize_weight(w: Tensor, dtype: torch.dtype = E4M3) -> Tuple[Tensor, Tensor, Tensor]:
    """float8_quantize.py:195-207 -> (float8_data, scale, scale_reciprocal)."""
    amax = torch.max(torch.abs(w)).float()
    scale = amax_to_scale(amax, fp8_max(dtype))
    return quantize(w, scale, dtype), scale, scale.reciprocal()


#: "restated": the arithmetic below.  "library": call the same PyTorch entry points the reference calls
#: (torch._scaled_mm with use_fast_accum, F.scaled_dot_product_attention) -- CUDA only; a round-1 cross-check
#: of the restatement against the library kernels.  (Round 2 runs the staged reference modules themselves on the
#: B200: oracle/ref_loader.py, tests/test_gpu_reference.py, bench.py's gpu_reference block.)
BACKEND = "restated"


def scaled_mm(xq: Tensor, wq: Tensor, sa_recip: Tensor, sw_recip: Tensor, bias: Optional[Tensor],
              out_dtype: torch.dtype = torch.bfloat16) -> Tensor:
    """torch._scaled_mm as called at float8_quantize.py:284-292: fp8 x fp8 products are exact in fp32,
    fp32 accumulation, then * scale_a * scale_b + bias, cast to out_dtype."""
    if BACKEND == "library":
        return torch._scaled_mm(xq, wq.T, scale_a=sa_recip, scale_b=sw_recip, bias=bias, out_dtype=out_dtype,
                                use_fast_accum=True)
    acc = xq.float() @ wq.float().t()
    out = acc * (sa_recip.float() * sw_recip.float())
    if bias is not None:
        out = out + bias.float()
    return out.to(out_dtype)


def f8linear(x: Tensor, p: Dict[str, Tensor], prefix: str, in_dtype: torch.dtype = E5M2) -> Tensor:
    """F8Linear.forward with frozen input scale (float8_quantize.py:272-296)."""
    xq = quantize(x, p[prefix + "input_scale"], in_dtype)
    lead = xq.shape[:-1]
    out = scaled_mm(xq.reshape(-1, xq.shape[-1]), p[prefix + "float8_data"], p[prefix + "input_scale_reciprocal"],
                    p[prefix + "scale_reciprocal"], p.get(prefix + "bias"), x.dtype)
    return out.reshape(*lead, -1)


def linear(x: Tensor, p: Dict[str, Tensor], prefix: str, in_dtype: torch.dtype = E5M2) -> Tensor:
    """nn.Linear or F8Linear depending on what the state-dict holds for `prefix`."""
    if prefix + "float8_data" in p:
        return f8linear(x, p, prefix, in_dtype)
    return F.linear(x, p[prefix + "weight"], p.get(prefix + "bias"))


class CalibratingLinear:
    """F8Linear.quantize_input's dynamic->static input-scale calibration (float8_quantize.py:220-246):
    calls 1..num_trials record amax and use the running-max scale; call num_trials+1 freezes."""

    def __init__(self, num_trials: int = 12, in_dtype: torch.dtype = E5M2):
        self.num_trials = num_trials
        self.in_dtype = in_dtype
        self.trials = torch.zeros(num_trials, dtype=torch.float32)
        self.index = 0
        self.initialized = False
        self.input_scale: Optional[Tensor] = None

    def quantize_input(self, x: Tensor) -> Tensor:
        m = fp8_max(self.in_dtype)
        if self.initialized:
            return quantize(x, self.input_scale, self.in_dtype)
        if self.index < self.num_trials:
            self.trials[self.index] = torch.max(torch.abs(x)).float()
            self.index += 1
            self.input_scale = amax_to_scale(self.trials[: self.index].max(), m)
        else:
            self.input_scale = amax_to_scale(self.trials.max(), m)
            self.initialized = True
        return quantize(x, self.input_scale, self.in_dtype)


# --------------------------------------------------------------------------------------------
# elementwise pieces of the blocks (modules/flux_model.py)
# --------------------------------------------------------------------------------------------
def layernorm_modulate(x: Tensor, shift: Tensor, scale: Tensor) -> Tensor:
    """nn.LayerNorm(D, elementwise_affine=False, eps=1e-6) then (1 + scale) * x + shift
    (modules/flux_model.py:367-368, 374-375, 389, 395, 469-470). shift/scale: [B,1,D]."""
    ln = F.layer_norm(x, (x.shape[-1],), eps=1e-6)
    return (1 + scale) * ln + shift


def rms_norm(x: Tensor, weight: Tensor) -> Tensor:
    """RMSNorm.forward (modules/flux_model.py:158-164): fp32 rms_norm, eps 1e-6, cast back."""
    xf = x.float()
    y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6) * weight.float()
    return y.to(x.dtype)


def rope_table(pos: Tensor, dim: int, theta: int) -> Tensor:
    """rope() (modules/flux_model.py:49-57): [..., dim/2, 2, 2] = [[cos,-sin],[sin,cos]] in fp32."""
    scale = torch.arange(0, dim, 2, dtype=torch.float32, device=pos.device) / dim
    omega = 1.0 / (theta**scale)
    ang = pos.float()[..., None] * omega  # einsum("...n,d->...nd") promotes pos to fp32
    out = torch.stack([torch.cos(ang), -torch.sin(ang), torch.sin(ang), torch.cos(ang)], dim=-1)
    return out.reshape(*out.shape[:-1], 2, 2)


def embed_nd(ids: Tensor, axes_dim, theta: int, dtype: torch.dtype) -> Tensor:
    """EmbedND.forward (modules/flux_model.py:82-92): per-axis tables cast to model dtype,
    concatenated on dim -3, unsqueeze(1) -> [B,1,S,64,2,2]."""
    emb = torch.cat([rope_table(ids[..., i], axes_dim[i], theta).to(dtype) for i in range(ids.shape[-1])], dim=-3)
    return emb.unsqueeze(1)


def apply_rope(xq: Tensor, xk: Tensor, pe: Tensor) -> Tuple[Tensor, Tensor]:
    """apply_rope (modules/flux_model.py:60-65): interleaved pairs, eager bf16 products and sum."""

    def rot(x: Tensor) -> Tensor:
        x_ = x.reshape(*x.shape[:-1], -1, 1, 2)
        out = pe[..., 0] * x_[..., 0] + pe[..., 1] * x_[..., 1]
        return out.reshape(*x.shape)

    return rot(xq), rot(xk)


#: "fp32": probabilities kept in fp32 for P.V (math backend).  "bf16": probabilities rounded to the input
#: dtype before P.V, as the fused flash / cuDNN kernels do.  Both are legitimate outcomes of the reference's
#: F.scaled_dot_product_attention call depending on the backend PyTorch dispatches to; the difference between
#: them is the reference's own noise floor, used by the full-width parity tests.
SDPA_P_DTYPE = "fp32"


def sdpa(q: Tensor, k: Tensor, v: Tensor) -> Tensor:
    """F.scaled_dot_product_attention(q,k,v) (modules/flux_model.py:43): scale 1/sqrt(d), no mask.
    Restated with fp32 scores and softmax; output in the input dtype."""
    if BACKEND == "library":
        return F.scaled_dot_product_attention(q, k, v)
    scale = 1.0 / math.sqrt(q.shape[-1])
    s = (q.float() @ k.float().transpose(-1, -2)) * scale
    p = torch.softmax(s, dim=-1)
    if SDPA_P_DTYPE == "bf16":
        p = p.to(q.dtype).float()
    return (p @ v.float()).to(q.dtype)


def attention(q: Tensor, k: Tensor, v: Tensor, pe: Tensor) -> Tensor:
    """attention() (modules/flux_model.py:41-45)."""
    q, k = apply_rope(q, k, pe)
    x = sdpa(q, k, v).transpose(1, 2)
    return x.reshape(*x.shape[:-2], -1)


def split_heads(qkv: Tensor, num_heads: int) -> Tuple[Tensor, Tensor, Tensor]:
    """rearrange_for_norm (modules/flux_model.py:350-354): [B,L,3*H*D] -> 3 x [B,H,L,D]."""
    B, L, D3 = qkv.shape
    q, k, v = qkv.reshape(B, L, 3, num_heads, D3 // (3 * num_heads)).permute(2, 0, 3, 1, 4)
    return q, k, v


def modulation(vec: Tensor, p: Dict[str, Tensor], prefix: str, double: bool, in_dtype: torch.dtype = E5M2):
    """Modulation.forward (modules/flux_model.py:251-257): lin(silu(vec))[:,None,:].chunk(6|3)."""
    out = linear(F.silu(vec), p, prefix + "lin.", in_dtype)[:, None, :].chunk(6 if double else 3, dim=-1)
    return out[:3], (out[3:] if double else None)


# --------------------------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------------------------
def double_block(img: Tensor, txt: Tensor, vec: Tensor, pe: Tensor, p: Dict[str, Tensor], prefix: str,
                 num_heads: int, in_dtype: torch.dtype = E5M2) -> Tuple[Tensor, Tensor]:
    """DoubleStreamBlock.forward (modules/flux_model.py:356-400), bf16 (no fp16 clamp)."""
    (i_shift1, i_scale1, i_gate1), (i_shift2, i_scale2, i_gate2) = modulation(vec, p, prefix + "img_mod.", True, in_dtype)
    (t_shift1, t_scale1, t_gate1), (t_shift2, t_scale2, t_gate2) = modulation(vec, p, prefix + "txt_mod.", True, in_dtype)

    def qkv_of(x, shift, scale, stream):
        xm = layernorm_modulate(x, shift, scale)
        q, k, v = split_heads(linear(xm, p, f"{prefix}{stream}_attn.qkv.", in_dtype), num_heads)
        q = rms_norm(q, p[f"{prefix}{stream}_attn.norm.query_norm.scale"])
        k = rms_norm(k, p[f"{prefix}{stream}_attn.norm.key_norm.scale"])
        return q, k, v

    iq, ik, iv = qkv_of(img, i_shift1, i_scale1, "img")
    tq, tk, tv = qkv_of(txt, t_shift1, t_scale1, "txt")
    q = torch.cat((tq, iq), dim=2)
    k = torch.cat((tk, ik), dim=2)
    v = torch.cat((tv, iv), dim=2)
    attn = attention(q, k, v, pe)
    t_attn, i_attn = attn[:, : txt.shape[1]], attn[:, txt.shape[1]:]

    def mlp(x, stream):
        h = linear(x, p, f"{prefix}{stream}_mlp.0.", in_dtype)
        return linear(F.gelu(h, approximate="tanh"), p, f"{prefix}{stream}_mlp.2.", in_dtype)

    img = img + i_gate1 * linear(i_attn, p, prefix + "img_attn.proj.", in_dtype)
    img = img + i_gate2 * mlp(layernorm_modulate(img, i_shift2, i_scale2), "img")
    txt = txt + t_gate1 * linear(t_attn, p, prefix + "txt_attn.proj.", in_dtype)
    txt = txt + t_gate2 * mlp(layernorm_modulate(txt, t_shift2, t_scale2), "txt")
    return img, txt


def single_block(x: Tensor, vec: Tensor, pe: Tensor, p: Dict[str, Tensor], prefix: str, num_heads: int,
                 in_dtype: torch.dtype = E5M2) -> Tensor:
    """SingleStreamBlock.forward (modules/flux_model.py:467-485), bf16."""
    (shift, scale, gate), _ = modulation(vec, p, prefix + "modulation.", False, in_dtype)
    hidden = x.shape[-1]
    x_mod = layernorm_modulate(x, shift, scale)
    lin1 = linear(x_mod, p, prefix + "linear1.", in_dtype)
    qkv, mlp = lin1[..., : 3 * hidden], lin1[..., 3 * hidden:]
    q, k, v = split_heads(qkv, num_heads)
    q = rms_norm(q, p[prefix + "norm.query_norm.scale"])
    k = rms_norm(k, p[prefix + "norm.key_norm.scale"])
    attn = attention(q, k, v, pe)
    out = linear(torch.cat((attn, F.gelu(mlp, approximate="tanh")), 2), p, prefix + "linear2.", in_dtype)
    return x + gate * out


# --------------------------------------------------------------------------------------------
# model container (caller of the hot path; modules/flux_model.py:95-155, 488-503, 672-716)
# --------------------------------------------------------------------------------------------
def timestep_embedding(t: Tensor, dim: int, max_period: int = 10000, time_factor: float = 1000.0) -> Tensor:
    """modules/flux_model.py:95-116 (t stays in its own dtype for `time_factor * t`)."""
    t = time_factor * t
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def mlp_embedder(x: Tensor, p: Dict[str, Tensor], prefix: str, in_dtype: torch.dtype = E5M2) -> Tensor:
    """MLPEmbedder.forward (modules/flux_model.py:154-155)."""
    return linear(F.silu(linear(x, p, prefix + "in_layer.", in_dtype)), p, prefix + "out_layer.", in_dtype)


def last_layer(x: Tensor, vec: Tensor, p: Dict[str, Tensor]) -> Tensor:
    """LastLayer.forward (modules/flux_model.py:499-503)."""
    shift, scale = linear(F.silu(vec), p, "final_layer.adaLN_modulation.1.").chunk(2, dim=1)
    x = (1 + scale[:, None, :]) * F.layer_norm(x, (x.shape[-1],), eps=1e-6) + shift[:, None, :]
    return linear(x, p, "final_layer.linear.")


def flux_forward(p: Dict[str, Tensor], cfg: dict, img: Tensor, img_ids: Tensor, txt: Tensor, txt_ids: Tensor,
                 timesteps: Tensor, y: Tensor, guidance: Optional[Tensor] = None,
                 in_dtype: torch.dtype = E5M2) -> Tensor:
    """Flux.forward (modules/flux_model.py:672-716).  cfg: dict(num_heads, depth, depth_single_blocks,
    axes_dim, theta, guidance_embed)."""
    dtype = img.dtype
    img = linear(img, p, "img_in.", in_dtype)
    vec = mlp_embedder(timestep_embedding(timesteps, 256).to(dtype), p, "time_in.", in_dtype)
    if cfg["guidance_embed"]:
        vec = vec + mlp_embedder(timestep_embedding(guidance, 256).to(dtype), p, "guidance_in.", in_dtype)
    vec = vec + mlp_embedder(y, p, "vector_in.", in_dtype)
    txt = linear(txt, p, "txt_in.", in_dtype)
    pe = embed_nd(torch.cat((txt_ids, img_ids), dim=1), cfg["axes_dim"], cfg["theta"], dtype)
    for i in range(cfg["depth"]):
        img, txt = double_block(img, txt, vec, pe, p, f"double_blocks.{i}.", cfg["num_heads"], in_dtype)
    x = torch.cat((txt, img), 1)
    for i in range(cfg["depth_single_blocks"]):
        x = single_block(x, vec, pe, p, f"single_blocks.{i}.", cfg["num_heads"], in_dtype)
    x = x[:, txt.shape[1]:, ...]
    return last_layer(x, vec, p)


# --------------------------------------------------------------------------------------------
# LoRA fuse / unfuse into an F8Linear (lora_loading.py)
# --------------------------------------------------------------------------------------------
def lora_delta(lora_A: Tensor, lora_B: Tensor, alpha, rank: Optional[int] = None, lora_scale: float = 1.0) -> Tensor:
    """calculate_lora_weight (lora_loading.py:509-547): fp32 `lora_scale * (lora_B @ lora_A)`, with lora_A scaled by
    alpha / rank when they differ, and an "uneven rank" lora_A ([c*r, K] against lora_B [N, r]) fused chunk by chunk."""
    uneven = lora_B.shape[1] != lora_A.shape[0]
    rank_diff = lora_A.shape[0] / lora_B.shape[1]
    if rank is None:
        rank = lora_B.shape[1]
    if alpha is None:
        alpha = rank
    up = lora_A.to(torch.float32)
    down = lora_B.to(torch.float32)
    if alpha != rank:
        up = up * alpha / rank
    if uneven:
        fused = torch.zeros((lora_B.shape[0], lora_A.shape[1]), dtype=torch.float32, device=up.device)
        for chunk in up.chunk(int(rank_diff), dim=0):
            fused = fused + (lora_scale * torch.mm(down, chunk))
        return fused
    return lora_scale * torch.mm(down, up)


def lora_fuse_f8(float8_data: Tensor, scale_reciprocal: Tensor, lora_A: Tensor, lora_B: Tensor, alpha,
                 lora_scale: float = 1.0, unfuse: bool = False, w_dtype: torch.dtype = torch.bfloat16,
                 f8_dtype: torch.dtype = E4M3) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """One F8Linear through apply_lora_to_model / remove_lora_from_module (lora_loading.py:679-687, 742-749):
    dequantise `float8_data.float() * scale_reciprocal` (:615-626), add (or subtract) the fp32 delta (:566-577,
    :549-563), cast to the layer's weight dtype, then set_weight_tensor -> quantize_weight
    (float8_quantize.py:209-212, 195-207).  Returns (fused weight in w_dtype, float8_data, scale, scale_reciprocal)."""
    w = float8_data.float().mul(scale_reciprocal)
    delta = lora_delta(lora_A, lora_B, alpha, None, lora_scale)
    fused = (w - delta) if unfuse else (w + delta)
    fused = fused.to(w_dtype)
    q, s, sr = quantize_weight(fused, f8_dtype)
    return fused, q, s, sr


# --------------------------------------------------------------------------------------------
# denoise-loop glue used by tests and the CPU baseline (flux_pipeline.py:315-344, 627-651)
# --------------------------------------------------------------------------------------------
def time_shift(mu: float, sigma: float, t: Tensor) -> Tensor:
    """flux_pipeline.py:315-316."""
    return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)


def get_schedule(num_steps: int, image_seq_len: int, base_shift: float = 0.5, max_shift: float = 1.15,
                 shift: bool = True) -> list:
    """flux_pipeline.py:318-344."""
    timesteps = torch.linspace(1, 0, num_steps + 1)
    if shift:
        m = (max_shift - base_shift) / (4096 - 256)
        b = base_shift - m * 256
        mu = m * image_seq_len + b
        timesteps = time_shift(mu, 1.0, timesteps)
    return timesteps.tolist()


def make_img_ids(batch: int, h2: int, w2: int, dtype: torch.dtype, device="cpu") -> Tensor:
    """flux_pipeline.py:280-292: ids[...,1] = row, ids[...,2] = col over the (H/16, W/16) grid."""
    ids = torch.zeros(h2, w2, 3, device=device, dtype=dtype)
    ids[..., 1] = ids[..., 1] + torch.arange(h2, device=device, dtype=dtype)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(w2, device=device, dtype=dtype)[None, :]
    return ids.reshape(1, h2 * w2, 3).repeat(batch, 1, 1)


def euler_step(img: Tensor, pred: Tensor, t_curr: float, t_prev: float) -> Tensor:
    """flux_pipeline.py:651."""
    return img + (t_prev - t_curr) * pred
