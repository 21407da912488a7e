"""CPU restatement of the text encoders behind the reference's HFEmbedder -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this; the product path
(flux-fp8-api_b200/conditioner.py) never does and has no CPU fallback.

The reference (modules/conditioner.py:37-117) delegates the arithmetic to the third-party `transformers` package, which is
NOT under /root/reference and which the reference does not pin (requirements.txt lists it without a version; this image has
5.5.0): `T5EncoderModel` (models/t5/modeling_t5.py: T5Stack, T5Block, T5LayerSelfAttention, T5Attention incl.
`_relative_position_bucket` / `compute_bias`, T5LayerFF, T5DenseGatedActDense, T5LayerNorm) and `CLIPTextModel`
(models/clip/modeling_clip.py: CLIPTextTransformer, CLIPTextEmbeddings, CLIPEncoderLayer, CLIPAttention, CLIPMLP,
quick_gelu).  Below is a functional restatement of those published algorithms over the modules' state dicts, in fp32;
oracle/make_golden.py pins it against the installed package (tiny configurations, seeded weights) and commits that
package's outputs as tests/golden/text_tiny.pt, so tests/test_oracle_golden.py re-checks the restatement without
transformers.  Parity of the CUDA path is anchored on the reference's call site: `hf_module(input_ids=ids,
attention_mask=None, output_hidden_states=False)[output_key]` (conditioner.py:109-114).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F
from torch import Tensor


def t5_relative_position_bucket(rel: Tensor, num_buckets: int = 32, max_distance: int = 128) -> Tensor:
    """T5Attention._relative_position_bucket (bidirectional)."""
    num_buckets //= 2
    out = (rel > 0).to(torch.long) * num_buckets
    rp = rel.abs()
    max_exact = num_buckets // 2
    large = max_exact + (torch.log(rp.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return out + torch.where(rp < max_exact, rp, large)


def t5_layer_norm(x: Tensor, w: Tensor, eps: float) -> Tensor:
    """T5LayerNorm.forward: no mean subtraction, no bias."""
    return w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))


def gelu_new(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def t5_encoder(sd: Dict[str, Tensor], cfg: dict, input_ids: Tensor) -> Tensor:
    """T5EncoderModel(input_ids, attention_mask=None).last_hidden_state in fp32."""
    sd = {k: v.float() for k, v in sd.items()}
    H, dk, eps = cfg["num_heads"], cfg["d_kv"], cfg["layer_norm_epsilon"]
    B, S = input_ids.shape
    x = sd["encoder.embed_tokens.weight"][input_ids]
    pos = torch.arange(S)
    bucket = t5_relative_position_bucket(pos[None, :] - pos[:, None], cfg["relative_attention_num_buckets"],
                                         cfg["relative_attention_max_distance"])
    bias = sd["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"][bucket].permute(2, 0, 1)[None]
    for i in range(cfg["num_layers"]):
        p = f"encoder.block.{i}.layer."
        h = t5_layer_norm(x, sd[p + "0.layer_norm.weight"], eps)
        q, k, v = (F.linear(h, sd[p + f"0.SelfAttention.{n}.weight"]).view(B, S, H, dk).transpose(1, 2) for n in "qkv")
        a = torch.softmax(q @ k.transpose(-1, -2) + bias, -1) @ v  # no 1/sqrt(d) in T5
        x = x + F.linear(a.transpose(1, 2).reshape(B, S, H * dk), sd[p + "0.SelfAttention.o.weight"])
        h = t5_layer_norm(x, sd[p + "1.layer_norm.weight"], eps)
        u = gelu_new(F.linear(h, sd[p + "1.DenseReluDense.wi_0.weight"])) * F.linear(h, sd[p + "1.DenseReluDense.wi_1.weight"])
        x = x + F.linear(u, sd[p + "1.DenseReluDense.wo.weight"])
    return t5_layer_norm(x, sd["encoder.final_layer_norm.weight"], eps)


def clip_text(sd: Dict[str, Tensor], cfg: dict, input_ids: Tensor):
    """CLIPTextModel(input_ids, attention_mask=None) -> (last_hidden_state, pooler_output) in fp32."""
    sd = {k: v.float() for k, v in sd.items()}
    H, D, eps = cfg["num_attention_heads"], cfg["hidden_size"], cfg["layer_norm_eps"]
    B, S = input_ids.shape
    p = "text_model."
    x = sd[p + "embeddings.token_embedding.weight"][input_ids] + sd[p + "embeddings.position_embedding.weight"][:S][None]
    mask = torch.full((S, S), float("-inf")).triu(1)
    for i in range(cfg["num_hidden_layers"]):
        q = f"{p}encoder.layers.{i}."
        h = F.layer_norm(x, (D,), sd[q + "layer_norm1.weight"], sd[q + "layer_norm1.bias"], eps)
        qq, kk, vv = (F.linear(h, sd[q + f"self_attn.{n}_proj.weight"], sd[q + f"self_attn.{n}_proj.bias"])
                      .view(B, S, H, D // H).transpose(1, 2) for n in "qkv")
        a = torch.softmax(qq @ kk.transpose(-1, -2) * (D // H) ** -0.5 + mask, -1) @ vv
        x = x + F.linear(a.transpose(1, 2).reshape(B, S, D), sd[q + "self_attn.out_proj.weight"], sd[q + "self_attn.out_proj.bias"])
        h = F.layer_norm(x, (D,), sd[q + "layer_norm2.weight"], sd[q + "layer_norm2.bias"], eps)
        u = F.linear(h, sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"])
        x = x + F.linear(u * torch.sigmoid(1.702 * u), sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"])
    x = F.layer_norm(x, (D,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], eps)
    ids = input_ids.to(torch.int)
    idx = ids.argmax(-1) if cfg["eos_token_id"] == 2 else (ids == cfg["eos_token_id"]).int().argmax(-1)
    return x, x[torch.arange(B), idx]


T5_TINY = dict(vocab_size=512, d_model=256, d_kv=64, d_ff=512, num_layers=2, num_heads=4, relative_attention_num_buckets=32,
               relative_attention_max_distance=128, dropout_rate=0.0, layer_norm_epsilon=1e-6, feed_forward_proj="gated-gelu",
               is_encoder_decoder=False, use_cache=False, tie_word_embeddings=False)
CLIP_TINY = dict(vocab_size=512, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                 max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5, attention_dropout=0.0, eos_token_id=2,
                 bos_token_id=0, pad_token_id=1)


def seeded_state_from_shapes(shapes: Dict[str, tuple], seed: int, scale: float = 1.0) -> Dict[str, Tensor]:
    """Seeded parameters for a Hugging Face text encoder given {key: (shape, dtype name)} (no checkpoints to download):
    N(0, 1/fan_in) linears, 0.5 N(0, 1) embeddings, norms 1 + 0.1 N (biases 0.05 N); keys visited in sorted order."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(shapes):
        shape, dtype = shapes[k]
        dtype = getattr(torch, dtype.replace("torch.", "")) if isinstance(dtype, str) else dtype
        if not dtype.is_floating_point:
            out[k] = torch.arange(shape[-1], dtype=dtype).expand(shape).clone()  # CLIP's position_ids buffer
            continue
        if "norm" in k and k.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif k.endswith("bias") and len(shape) == 1:
            t = 0.05 * torch.randn(shape, generator=g)
        elif "embed" in k or "shared" in k or "relative_attention_bias" in k:
            t = 0.5 * torch.randn(shape, generator=g)
        else:
            t = torch.randn(shape, generator=g) * (scale / math.sqrt(shape[-1]))
        out[k] = t.to(dtype)
    if "shared.weight" in out and "encoder.embed_tokens.weight" in out:
        out["encoder.embed_tokens.weight"] = out["shared.weight"]
    return out


def seeded_state(module, seed: int, scale: float = 1.0) -> Dict[str, Tensor]:
    return seeded_state_from_shapes({k: (tuple(v.shape), str(v.dtype)) for k, v in module.state_dict().items()}, seed, scale)
