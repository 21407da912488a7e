"""Text encoders on the B200 kernels (second half): what `modules/conditioner.py:HFEmbedder` runs.

The reference's `HFEmbedder` (conditioner.py:37-117) owns a Hugging Face `T5EncoderModel` (t5-v1_1-xxl: 24 blocks, d 4096,
64 heads x 64, gated-gelu FF 10240, relative position bias) or `CLIPTextModel` (clip-vit-large-patch14: 12 layers, d 768,
12 heads x 64, causal, quick_gelu) as `self.hf_module` and calls it as

    outputs = self.hf_module(input_ids=..., attention_mask=None, output_hidden_states=False)       (:109-113)
    return outputs[self.output_key]            # "last_hidden_state" (T5) / "pooler_output" (CLIP)   (:114)

`accelerate(hf_module)` returns an object with that call signature and those output keys whose forward runs through
libflux_b200.so: dense layers on the tcgen05 implicit-GEMM kernel (`ops.dense`: q | k | v and wi_0 | wi_1 fused along N,
residual adds in the epilogue), T5LayerNorm / LayerNorm, the gated activation and the head-dim-64 attention on their own
kernels.  The wrapper READS the Hugging Face module's parameters (fused copies are cached on parameter storage / version);
it does not re-implement tokenisation, loading or quantised variants (`quantization_dtype` qfloat8 / qint4 / ... stay on the
Hugging Face path: `accelerate` refuses modules whose linears are not plain bf16 `nn.Linear`).  A maintainer adds, after
`HFEmbedder.__init__`:

    from flux_fp8_api_b200 import conditioner
    embedder.hf_module = conditioner.accelerate(embedder.hf_module)

Numerics follow the eager Hugging Face modules op by op where an op rounds to bf16 (matmul outputs, the score scale /
bias adds, T5LayerNorm's two roundings, residual sums); gelu_new / quick_gelu and LayerNorm are evaluated in fp32 and
rounded once (the eager modules round inside them) -- tests compare against the Hugging Face modules on the same GPU with
the modules' own bf16-vs-fp32 distance as the floor.
"""
from __future__ import annotations

import gc
import math
from typing import Optional

import torch
from torch import Tensor, nn

from . import ops
from .blocks import tensor_version

BF16 = torch.bfloat16


class _Output(dict):
    """Dict with attribute access: `outputs["last_hidden_state"]` and `outputs.last_hidden_state` (as ModelOutput allows)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None


def _plain_bf16_linear(lin) -> bool:
    return isinstance(lin, nn.Linear) and lin.weight.dtype == BF16 and lin.weight.is_cuda


class _Fused:
    """torch.cat of several parameters, rebuilt when any of them changes storage or version."""

    def __init__(self):
        self.key = None
        self.t = None

    def get(self, *params: Tensor) -> Tensor:
        key = tuple((p.data_ptr(), tensor_version(p)) for p in params)
        if key != self.key:
            self.t = torch.cat([p.detach() for p in params], 0).contiguous()
            self.key = key
        return self.t


class _GraphedForward:
    """Replays a module's `_forward(input_ids)` as one CUDA graph per (batch, tokens) shape.  The text encoders are ~200
    (T5) / ~100 (CLIP) small launches of a few-hundred-row problem: issued one by one from Python they are bound by the
    host, not the GPU.  The graph is re-captured when any parameter of the wrapped module changes storage or version.
    Holds no reference to its owner (a reference cycle would leave the destruction of a captured graph to the cyclic
    garbage collector, i.e. possibly to the middle of somebody else's stream capture, which it invalidates)."""

    def __init__(self):
        self.graphs = {}

    def __call__(self, owner: nn.Module, input_ids: Tensor):
        if not owner.use_graph:
            return owner._forward(input_ids)
        key = (tuple(input_ids.shape), input_ids.device)
        wkey = tuple((p.data_ptr(), tensor_version(p)) for p in owner.hf_module.parameters())
        hit = self.graphs.get(key)
        if hit is None or hit[3] != wkey:
            self.graphs.pop(key, None)
            static_ids = input_ids.clone()
            owner._forward(static_ids)  # warm-up: fused weights, position bias, allocator pools
            gc.collect()                # nothing may be torn down while the stream is capturing
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = owner._forward(static_ids)
            self.graphs[key] = hit = (g, static_ids, out, wkey)
        g, static_ids, out, _ = hit
        static_ids.copy_(input_ids)
        g.replay()
        return _Output({k: v.clone() for k, v in out.items()})


def t5_relative_position_bucket(relative_position: Tensor, num_buckets: int = 32, max_distance: int = 128) -> Tensor:
    """T5Attention._relative_position_bucket, bidirectional (encoder) form."""
    num_buckets //= 2
    buckets = (relative_position > 0).to(torch.long) * num_buckets
    rp = torch.abs(relative_position)
    max_exact = num_buckets // 2
    is_small = rp < max_exact
    large = max_exact + (torch.log(rp.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return buckets + torch.where(is_small, rp, large)


class T5EncoderB200(nn.Module):
    """`T5EncoderModel.forward(input_ids, attention_mask=None)` -> {"last_hidden_state": bf16 [B, S, d_model]}."""

    def __init__(self, hf_module: nn.Module):
        super().__init__()
        self.hf_module = hf_module
        cfg = hf_module.config
        if cfg.d_kv != 64:
            raise ValueError(f"T5EncoderB200: the attention kernel is built for d_kv = 64, config has {cfg.d_kv}")
        if not getattr(cfg, "is_gated_act", False) or "gelu" not in cfg.dense_act_fn:
            raise ValueError("T5EncoderB200: only the gated-gelu feed-forward of t5-v1_1 is implemented")
        enc = hf_module.encoder
        for blk in enc.block:
            att, ff = blk.layer[0].SelfAttention, blk.layer[1].DenseReluDense
            for lin in (att.q, att.k, att.v, att.o, ff.wi_0, ff.wi_1, ff.wo):
                if not _plain_bf16_linear(lin) or lin.bias is not None:
                    raise ValueError("T5EncoderB200 needs plain bf16 nn.Linear layers on a CUDA device (quantised encoders stay "
                                     "on the Hugging Face path)")
        self.cfg = cfg
        self._qkv = [_Fused() for _ in enc.block]
        self._wi = [_Fused() for _ in enc.block]
        self._bias_cache = {}
        self.use_graph = True
        self._graphed = _GraphedForward()

    @property
    def device(self):
        return self.hf_module.device

    def position_bias(self, S: int) -> Tensor:
        """compute_bias of block 0 (shared by all blocks): bf16 [H, S, S]."""
        emb = self.hf_module.encoder.block[0].layer[0].SelfAttention.relative_attention_bias.weight
        key = (S, emb.data_ptr(), tensor_version(emb))
        hit = self._bias_cache.get(S)  # one entry per length: captured graphs keep reading the tensor they were built with
        if hit is None or hit[0] != key:
            pos = torch.arange(S, dtype=torch.long, device=emb.device)
            bucket = t5_relative_position_bucket(pos[None, :] - pos[:, None], self.cfg.relative_attention_num_buckets,
                                                 self.cfg.relative_attention_max_distance)
            bias = emb.detach()[bucket].permute(2, 0, 1).contiguous().to(BF16)  # [H, S, S]
            self._bias_cache[S] = hit = (key, bias)
        return hit[1]

    @torch.inference_mode()
    def forward(self, input_ids: Tensor, attention_mask: Optional[Tensor] = None, output_hidden_states: bool = False, **kwargs):
        if attention_mask is not None:
            raise NotImplementedError("T5EncoderB200: the reference always passes attention_mask=None (conditioner.py:111)")
        return self._graphed(self, input_ids)

    def _forward(self, input_ids: Tensor):
        enc, cfg = self.hf_module.encoder, self.cfg
        B, S = input_ids.shape
        H, D, F = cfg.num_heads, cfg.d_model, cfg.d_ff
        eps = cfg.layer_norm_epsilon
        x = enc.embed_tokens.weight.detach()[input_ids.reshape(-1)].contiguous()  # [B*S, D] (embedding gather)
        if x.dtype != BF16:
            raise ValueError("T5EncoderB200: bf16 encoder expected")
        bias = self.position_bias(S)
        for i, blk in enumerate(enc.block):
            att, ff = blk.layer[0].SelfAttention, blk.layer[1].DenseReluDense
            h = ops.rows_norm(x, blk.layer[0].layer_norm.weight.detach(), None, eps)
            qkv = ops.dense(h, self._qkv[i].get(att.q.weight, att.k.weight, att.v.weight))
            a = ops.attention_d64(qkv, B, S, H, bias, 1.0, False)
            x = ops.dense(a, att.o.weight.detach(), residual=x)
            h = ops.rows_norm(x, blk.layer[1].layer_norm.weight.detach(), None, eps)
            u = ops.dense(h, self._wi[i].get(ff.wi_0.weight, ff.wi_1.weight))
            g = ops.gated_act(u, F, 0)
            x = ops.dense(g, ff.wo.weight.detach(), residual=x)
        x = ops.rows_norm(x, enc.final_layer_norm.weight.detach(), None, eps)
        return _Output(last_hidden_state=x.view(B, S, D))


class CLIPTextB200(nn.Module):
    """`CLIPTextModel.forward(input_ids, attention_mask=None)` -> {"last_hidden_state", "pooler_output"}."""

    def __init__(self, hf_module: nn.Module):
        super().__init__()
        self.hf_module = hf_module
        tm = hf_module.text_model
        cfg = hf_module.config
        if cfg.hidden_size // cfg.num_attention_heads != 64:
            raise ValueError("CLIPTextB200: the attention kernel is built for head dim 64")
        if cfg.hidden_act != "quick_gelu":
            raise ValueError(f"CLIPTextB200: activation {cfg.hidden_act!r} is not implemented (clip-vit-large-patch14 uses quick_gelu)")
        for layer in tm.encoder.layers:
            for lin in (layer.self_attn.q_proj, layer.self_attn.k_proj, layer.self_attn.v_proj, layer.self_attn.out_proj,
                        layer.mlp.fc1, layer.mlp.fc2):
                if not _plain_bf16_linear(lin):
                    raise ValueError("CLIPTextB200 needs plain bf16 nn.Linear layers on a CUDA device")
        self.cfg = cfg
        self._qkv_w = [_Fused() for _ in tm.encoder.layers]
        self._qkv_b = [_Fused() for _ in tm.encoder.layers]
        self.use_graph = True
        self._graphed = _GraphedForward()

    @property
    def device(self):
        return self.hf_module.device

    @torch.inference_mode()
    def forward(self, input_ids: Tensor, attention_mask: Optional[Tensor] = None, output_hidden_states: bool = False, **kwargs):
        if attention_mask is not None:
            raise NotImplementedError("CLIPTextB200: the reference always passes attention_mask=None (conditioner.py:111)")
        return self._graphed(self, input_ids)

    def _forward(self, input_ids: Tensor):
        tm, cfg = self.hf_module.text_model, self.cfg
        B, S = input_ids.shape
        H, D, F = cfg.num_attention_heads, cfg.hidden_size, cfg.intermediate_size
        eps = cfg.layer_norm_eps
        emb = tm.embeddings
        x = (emb.token_embedding.weight.detach()[input_ids] + emb.position_embedding.weight.detach()[:S][None]).reshape(B * S, D)
        x = x.contiguous()
        for i, layer in enumerate(tm.encoder.layers):
            sa, mlp = layer.self_attn, layer.mlp
            h = ops.rows_norm(x, layer.layer_norm1.weight.detach(), layer.layer_norm1.bias.detach(), eps)
            qkv = ops.dense(h, self._qkv_w[i].get(sa.q_proj.weight, sa.k_proj.weight, sa.v_proj.weight),
                            self._qkv_b[i].get(sa.q_proj.bias, sa.k_proj.bias, sa.v_proj.bias))
            a = ops.attention_d64(qkv, B, S, H, None, 0.125, True)
            x = ops.dense(a, sa.out_proj.weight.detach(), sa.out_proj.bias.detach(), residual=x)
            h = ops.rows_norm(x, layer.layer_norm2.weight.detach(), layer.layer_norm2.bias.detach(), eps)
            u = ops.dense(h, mlp.fc1.weight.detach(), mlp.fc1.bias.detach())
            g = ops.gated_act(u, F, 1)
            x = ops.dense(g, mlp.fc2.weight.detach(), mlp.fc2.bias.detach(), residual=x)
        x = ops.rows_norm(x, tm.final_layer_norm.weight.detach(), tm.final_layer_norm.bias.detach(), eps).view(B, S, D)
        ids = input_ids.to(dtype=torch.int, device=x.device)
        eos = getattr(tm, "eos_token_id", cfg.eos_token_id)
        idx = ids.argmax(dim=-1) if eos == 2 else (ids == eos).int().argmax(dim=-1)
        pooled = x[torch.arange(B, device=x.device), idx]
        return _Output(last_hidden_state=x, pooler_output=pooled)


def accelerate(hf_module: nn.Module) -> nn.Module:
    """Wrap the `hf_module` of a reference `HFEmbedder` (conditioner.py:80-92): T5EncoderModel -> T5EncoderB200,
    CLIPTextModel -> CLIPTextB200.  Raises for anything else (quantised encoders, other architectures)."""
    name = type(hf_module).__name__
    if name == "T5EncoderModel":
        return T5EncoderB200(hf_module)
    if name == "CLIPTextModel":
        return CLIPTextB200(hf_module)
    raise ValueError(f"conditioner.accelerate: unsupported text encoder {name}")
