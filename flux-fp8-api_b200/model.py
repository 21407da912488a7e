"""Flux container: the caller of the hot path (reference: modules/flux_model.py:24-36, 95-155, 488-734;
SURVEY.md row a11 / N1).  Same constructor, attribute names and forward signature as the reference's
`Flux`, so reference checkpoints (BFL key layout, optionally prequantised) load with load_state_dict.

The embedders and the final layer are < 0.03 % of the step's FLOPs and stay bf16 torch (cuBLAS) modules
unless quantize_flow_embedder_layers is set (then they are F8Linear and run on our kernels).  What this
container adds over the reference's per-step recomputation:

`txt_in(txt)`, `vector_in(y)`, `guidance_in(...)` and the RoPE table are step-invariant: they are computed
once per request and reused for every denoise step (the reference recomputes them each step,
modules/flux_model.py:694-702).
"""
from __future__ import annotations

import math
import weakref
from dataclasses import dataclass, field
from typing import List, Optional

import torch
from torch import Tensor, nn

from . import _cabi as cabi
from . import ops
from .blocks import (DoubleStreamBlock, EmbedND, ModulationBank, SingleStreamBlock, extract_cos_sin,
                     tensor_version)
from .f8linear import F8Linear

BF16 = torch.bfloat16


@dataclass
class FluxParams:
    in_channels: int = 64
    vec_in_dim: int = 768
    context_in_dim: int = 4096
    hidden_size: int = 3072
    mlp_ratio: float = 4.0
    num_heads: int = 24
    depth: int = 19
    depth_single_blocks: int = 38
    axes_dim: List[int] = field(default_factory=lambda: [16, 56, 56])
    theta: int = 10_000
    qkv_bias: bool = True
    guidance_embed: bool = True


@dataclass
class FluxSpec:
    """The subset of the reference's ModelSpec (util.py:38-79) that Flux.__init__ reads."""

    params: FluxParams = field(default_factory=FluxParams)
    prequantized_flow: bool = False
    quantize_modulation: bool = True
    quantize_flow_embedder_layers: bool = False


def flux_dev_spec(**kw) -> FluxSpec:
    """Flux.1-dev hyper-parameters (reference util.py:163-176)."""
    return FluxSpec(params=FluxParams(guidance_embed=True), **kw)


def flux_schnell_spec(**kw) -> FluxSpec:
    return FluxSpec(params=FluxParams(guidance_embed=False), **kw)


def timestep_embedding(t: Tensor, dim, max_period=10000, time_factor: float = 1000.0):
    """Sinusoidal embedding, fp32 (reference modules/flux_model.py:95-116).  `t` keeps its own dtype for
    the `time_factor * t` product (bf16 in the pipeline), as in the reference."""
    t = time_factor * t
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _linear(in_f: int, out_f: int, f8: bool) -> nn.Module:
    return F8Linear(in_features=in_f, out_features=out_f, bias=True) if f8 else nn.Linear(in_f, out_f, bias=True)


def _skinny_bf16(x: Tensor, *lins: nn.Module) -> bool:
    """True when `lins` are plain bf16 nn.Linear layers on the GPU and x is a skinny bf16 matrix our GEMV kernel takes
    (fluxb200_bf16_gemv: B <= 16, K % 32 == 0, K <= 4096).  Otherwise torch's library GEMM runs (larger batches)."""
    if x.dim() != 2 or x.dtype != BF16 or not x.is_cuda or x.shape[0] > 16:
        return False
    for l in lins:
        if type(l) is not nn.Linear or l.weight.dtype != BF16 or not l.weight.is_cuda or not l.weight.is_contiguous():
            return False
        if l.in_features % 32 or l.in_features > 4096 or (l.bias is not None and l.bias.dtype != BF16):
            return False
    return True


def _small_bf16_linear(x: Tensor, lin: nn.Module) -> bool:
    """True when `lin` is a plain bf16 nn.Linear our small mma.sync GEMM takes (fluxb200_bf16_gemm_small)."""
    return (type(lin) is nn.Linear and x.dtype == BF16 and x.is_cuda and lin.weight.dtype == BF16 and lin.weight.is_cuda
            and lin.weight.is_contiguous() and lin.in_features % 32 == 0 and lin.out_features % 2 == 0
            and (lin.bias is None or lin.bias.dtype == BF16) and x.shape[-1] == lin.in_features
            and (x.stride(-1) == 1 and (x.dim() < 2 or x.stride(-2) % 8 == 0)))


class MLPEmbedder(nn.Module):
    def __init__(self, in_dim: int, hidden_dim: int, prequantized: bool = False, quantized=False):
        super().__init__()
        f8 = prequantized and quantized
        self.in_layer = _linear(in_dim, hidden_dim, f8)
        self.silu = nn.SiLU()
        self.out_layer = _linear(hidden_dim, hidden_dim, f8)

    def forward(self, x: Tensor, add0: Optional[Tensor] = None, add1: Optional[Tensor] = None) -> Tensor:
        """out_layer(silu(in_layer(x))) [+ add0] [+ add1] (reference modules/flux_model.py:154-155; the optional addends
        are Flux.forward's `vec = vec + ...` sums, :694-697, carried by the out_layer launch)."""
        if _skinny_bf16(x, self.in_layer, self.out_layer):
            h = ops.bf16_gemv(x, self.in_layer.weight, self.in_layer.bias)
            return ops.bf16_gemv(h, self.out_layer.weight, self.out_layer.bias, silu_input=True, add0=add0, add1=add1)
        out = self.out_layer(self.silu(self.in_layer(x)))
        for a in (add0, add1):
            if a is not None:
                out = out + a
        return out


class LastLayer(nn.Module):
    def __init__(self, hidden_size: int, patch_size: int, out_channels: int):
        super().__init__()
        self.norm_final = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.linear = nn.Linear(hidden_size, patch_size * patch_size * out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 2 * hidden_size, bias=True))

    def forward(self, x: Tensor, vec: Tensor, euler=None) -> Tensor:
        """Reference signature (x, vec).  `euler = (img, dt, out)` (pipeline.GraphedStep) additionally applies the Euler
        update of flux_pipeline.py:651 to the prediction inside the projection's launch: out = img + dt * linear(...)."""
        lin = self.adaLN_modulation[1]
        if _skinny_bf16(vec, lin) and x.dtype == BF16 and x.is_cuda and x.shape[-1] % 256 == 0 and x.shape[-1] <= 4096:
            # SiLU + adaLN linear as one weight-streaming launch, LayerNorm + modulate as one launch (the same kernel
            # the blocks use, bf16 output), the [L, 3072] x [3072, 64] projection (+ Euler update) as one mma.sync launch
            shift, scale = ops.bf16_gemv(vec, lin.weight, lin.bias, silu_input=True).chunk(2, dim=1)
            _, xm = ops.ln_mod_quant(x, shift, scale, None, None, want_bf16=True, eps=self.norm_final.eps)
            if _small_bf16_linear(xm, self.linear):
                if euler is not None:
                    img, dt, out = euler
                    return ops.bf16_gemm_small(xm, self.linear.weight, self.linear.bias, euler_img=img, euler_dt=dt, out=out)
                return ops.bf16_gemm_small(xm, self.linear.weight, self.linear.bias)
            pred = self.linear(xm)
        else:
            shift, scale = self.adaLN_modulation(vec).chunk(2, dim=1)
            x = (1 + scale[:, None, :]) * self.norm_final(x) + shift[:, None, :]
            pred = self.linear(x)
        if euler is not None:
            img, dt, out = euler
            if pred.dtype == BF16:
                return ops.euler_update(img, pred, dt, out=out)
            out.copy_(img + (dt * pred.float()).to(pred.dtype))
            return out
        return pred


class _StepInvariantCache:
    """Results that depend only on per-request inputs.  A hit needs the very same tensor objects (held by
    weak reference, so a recycled address can never alias) with unchanged in-place version counters."""

    def __init__(self):
        self._store = {}

    def get(self, name: str, tensors, compute):
        hit = self._store.get(name)
        if hit is not None:
            refs, versions, val = hit
            if all(r() is t for r, t in zip(refs, tensors)) and versions == tuple(tensor_version(t) for t in tensors):
                return val
        val = compute()
        self._store[name] = (tuple(weakref.ref(t) for t in tensors), tuple(tensor_version(t) for t in tensors), val)
        return val

    def clear(self):
        self._store.clear()


class Flux(nn.Module):
    """Transformer model for flow matching on sequences (B200 hot path inside)."""

    def __init__(self, config, dtype: torch.dtype = torch.float16):
        super().__init__()
        p = config.params
        self.dtype = dtype
        self.params = p
        self.in_channels = self.out_channels = p.in_channels
        self.loras: list = []
        preq = config.prequantized_flow
        q_embed = config.quantize_flow_embedder_layers and preq
        q_mod = config.quantize_modulation and preq
        if p.hidden_size % p.num_heads != 0:
            raise ValueError(f"Hidden size {p.hidden_size} must be divisible by num_heads {p.num_heads}")
        pe_dim = p.hidden_size // p.num_heads
        if sum(p.axes_dim) != pe_dim:
            raise ValueError(f"Got {p.axes_dim} but expected positional dim {pe_dim}")
        self.hidden_size, self.num_heads = p.hidden_size, p.num_heads
        self.pe_embedder = EmbedND(dim=pe_dim, theta=p.theta, axes_dim=p.axes_dim, dtype=self.dtype)
        self.img_in = _linear(self.in_channels, self.hidden_size, q_embed)
        self.time_in = MLPEmbedder(256, self.hidden_size, prequantized=preq, quantized=q_embed)
        self.vector_in = MLPEmbedder(p.vec_in_dim, self.hidden_size, prequantized=preq, quantized=q_embed)
        self.guidance_in = (MLPEmbedder(256, self.hidden_size, prequantized=preq, quantized=q_embed)
                            if p.guidance_embed else nn.Identity())
        self.txt_in = _linear(p.context_in_dim, self.hidden_size, q_embed)
        self.double_blocks = nn.ModuleList([
            DoubleStreamBlock(self.hidden_size, self.num_heads, mlp_ratio=p.mlp_ratio, qkv_bias=p.qkv_bias,
                              dtype=self.dtype, quantized_modulation=q_mod, prequantized=preq)
            for _ in range(p.depth)])
        self.single_blocks = nn.ModuleList([
            SingleStreamBlock(self.hidden_size, self.num_heads, mlp_ratio=p.mlp_ratio, dtype=self.dtype,
                              quantized_modulation=q_mod, prequantized=preq)
            for _ in range(p.depth_single_blocks)])
        self.final_layer = LastLayer(self.hidden_size, 1, self.out_channels)
        self._cache = _StepInvariantCache()
        #: bumped whenever weights under the step-invariant embeddings change (LoRA on txt_in / vector_in /
        #: guidance_in): holders of derived state (pipeline.GraphedStep) compare it and re-capture
        self._invariant_epoch = 0
        #: set False to recompute txt_in / vector_in / pe every step exactly as the reference does
        self.cache_step_invariants = True
        #: one batched launch for all Modulation.lin of a step (False: each block runs its own, as the reference)
        self.batch_modulation = True

    # ---- request cache ---------------------------------------------------------------------------------------
    def reset_request_cache(self) -> None:
        self._cache.clear()

    def invalidate_step_invariants(self) -> None:
        """Weights that the cached per-request embeddings were computed from have changed."""
        self._cache.clear()
        self._invariant_epoch += 1

    def use_request_cache(self, cache: "_StepInvariantCache"):
        """Context manager: run forwards against `cache` instead of the model's own.  pipeline.GraphedStep gives every
        captured graph a private cache, so the tensors its kernels read (txt_in(txt), the vector / guidance embeddings,
        pe, cos/sin) stay alive for as long as the graph does, whatever other requests or sessions do to the model's
        shared cache."""
        model = self

        class _Swap:
            def __enter__(self_inner):
                self_inner.prev, model._cache = model._cache, cache

            def __exit__(self_inner, *exc):
                model._cache = self_inner.prev

        return _Swap()

    # ---- LoRA management (reference modules/flux_model.py:621-670) ----------------------------------------
    def get_lora(self, identifier: str):
        for lora in self.loras:
            if lora.path == identifier or lora.name == identifier:
                return lora

    def has_lora(self, identifier: str):
        for lora in self.loras:
            if lora.path == identifier or lora.name == identifier:
                return True

    def load_lora(self, path, scale: float, name: str = None):
        """Fuse a LoRA into the quantised weights on the device (lora.apply_lora_to_model).  `path` is a .safetensors
        file in the BFL key layout or an already-loaded state dict (then `name` identifies it)."""
        from . import lora as L

        ident = path if isinstance(path, str) else (name or f"<state-dict {id(path):x}>")
        if self.has_lora(ident):
            lora = self.get_lora(ident)
            if lora.scale == scale:
                import warnings

                warnings.warn(f"Lora {lora.name} already loaded with same scale - ignoring!")
            else:
                L.remove_lora_from_module(self, lora, lora.scale)
                L.apply_lora_to_model(self, lora, scale)
                for idx, lora_ in enumerate(self.loras):
                    if lora_.path == lora.path:
                        self.loras[idx].scale = scale
                        break
        else:
            _, weights = L.apply_lora_to_model(self, path, scale, return_lora_resolved=True)
            self.loras.append(L.LoraWeights(weights, ident, name, scale))

    def unload_lora(self, path_or_identifier: str):
        from . import lora as L

        for idx, lora_ in enumerate(list(self.loras)):
            if lora_.path == path_or_identifier or lora_.name == path_or_identifier:
                L.remove_lora_from_module(self, lora_.weights, lora_.scale)
                self.loras.pop(idx)
                return True
        import warnings

        warnings.warn(f"Couldn't remove lora {path_or_identifier} as it wasn't found fused to the model!")
        return False

    # ---- modulation ------------------------------------------------------------------------------------------
    def _modulation_bank(self, vec: Tensor) -> Optional[ModulationBank]:
        """The batched modulation launch, built lazily once every Modulation.lin is a frozen F8Linear -- or when all
        of them are bf16 nn.Linear (quantize_modulation=False).  None while calibrating, for mixed layer types, and
        for batches beyond the kernel's 16 rows (then every block runs its own Modulation, as the reference does)."""
        if not self.batch_modulation or vec.dtype != BF16 or vec.shape[0] > ModulationBank.MAX_BATCH:
            return None
        bank = self.__dict__.get("_mod_bank")
        if bank is not None and not bank.stale():
            return bank
        mods = []
        for b in self.double_blocks:
            mods += [b.img_mod, b.txt_mod]
        mods += [b.modulation for b in self.single_blocks]
        try:
            bank = ModulationBank(mods)
        except ValueError:
            return None
        self.__dict__["_mod_bank"] = bank
        return bank

    def _invariants_cacheable(self) -> bool:
        """While an embedder F8Linear is still calibrating it must see one call per step, like the reference."""
        if not self.cache_step_invariants:
            return False
        for mod in (self.txt_in, self.vector_in, self.guidance_in):
            for m in mod.modules():
                if isinstance(m, F8Linear) and not m.frozen:
                    return False
        return True

    def forward(self, img: Tensor, img_ids: Tensor, txt: Tensor, txt_ids: Tensor, timesteps: Tensor, y: Tensor,
                guidance: Optional[Tensor] = None) -> Tensor:
        return self._forward(img, img_ids, txt, txt_ids, timesteps, y, guidance)

    def denoise_step(self, img: Tensor, img_ids: Tensor, txt: Tensor, txt_ids: Tensor, timesteps: Tensor, y: Tensor,
                     guidance: Optional[Tensor], dt: Tensor, out: Tensor) -> Tensor:
        """One Euler step of the flow (flux_pipeline.py:641-651): out = img + dt * forward(img, ...), with the update
        fused into the final projection's launch.  `dt` is a 0-dim fp32 device tensor (t_prev - t_curr)."""
        return self._forward(img, img_ids, txt, txt_ids, timesteps, y, guidance, euler=(img, dt, out))

    def _forward(self, img, img_ids, txt, txt_ids, timesteps, y, guidance=None, euler=None) -> Tensor:
        if img.ndim != 3 or txt.ndim != 3:
            raise ValueError("Input img and txt tensors must have 3 dimensions.")
        cabi.require_cuda(img, txt)
        cache = self._cache if self._invariants_cacheable() else _StepInvariantCache()
        latent = img

        img = ops.bf16_gemm_small(img, self.img_in.weight, self.img_in.bias) if _small_bf16_linear(img, self.img_in) \
            else self.img_in(img)

        def t_emb(t: Tensor) -> Tensor:
            if t.dtype == BF16 and self.dtype == BF16 and t.is_cuda:
                return ops.timestep_embedding(t, 256)  # one kernel instead of ~10 eager elementwise launches
            return timestep_embedding(t, 256).type(self.dtype)

        g_emb = None
        if self.params.guidance_embed:
            if guidance is None:
                raise ValueError("Didn't get guidance strength for guidance distilled model.")
            g_emb = cache.get("guidance", (guidance,), lambda: self.guidance_in(t_emb(guidance)))
        y_emb = cache.get("vector", (y,), lambda: self.vector_in(y))
        # vec = time_in(..) [+ guidance_in(..)] + vector_in(y): the additions ride on time_in's out_layer launch
        vec = self.time_in(t_emb(timesteps), g_emb if g_emb is not None else y_emb, y_emb if g_emb is not None else None)
        txt = cache.get("txt", (txt,), lambda: self.txt_in(txt))

        def _pe():
            pe_ = self.pe_embedder(torch.cat((txt_ids, img_ids), dim=1))
            # the (cos, sin) pair the kernels read is extracted HERE, next to the table it comes from, and handed
            # down to every block: no cache keyed on addresses anywhere on the steady-state path
            return (pe_, extract_cos_sin(pe_) if pe_.dtype == BF16 else None)

        pe, rope = cache.get("pe", (txt_ids, img_ids), _pe)

        T = txt.shape[1]
        bank = self._modulation_bank(vec)
        if bank is None:
            for block in self.double_blocks:
                img, txt = block(img=img, txt=txt, vec=vec, pe=pe, rope=rope)
            x = torch.cat((txt, img), 1)
            for block in self.single_blocks:
                x = block(x, vec=vec, pe=pe, rope=rope)
        else:
            mods = iter(bank(vec))  # every block's shift/scale/gate from one batched launch
            for block in self.double_blocks:
                img, txt = block(img=img, txt=txt, vec=vec, pe=pe, mods=(next(mods), next(mods)), rope=rope)
            x = torch.cat((txt, img), 1)
            for block in self.single_blocks:
                x = block(x, vec=vec, pe=pe, mod=next(mods)[0], rope=rope)
        x = x[:, T:, ...]
        if euler is not None:
            return self.final_layer(x, vec, euler=(latent, euler[1], euler[2]))
        return self.final_layer(x, vec)

    @classmethod
    def from_pretrained(cls, path: str, dtype: torch.dtype = torch.float16) -> "Flux":
        """Reference modules/flux_model.py:718-734: `path` is a JSON model spec (the reference's ModelSpec fields that
        Flux reads: params, prequantized_flow, quantize_modulation, quantize_flow_embedder_layers, ckpt_path) whose
        `ckpt_path` names a .safetensors state dict in the BFL key layout.  Returns the model on the CPU, like the
        reference; move it with .to("cuda") before use (there is no CPU compute path)."""
        import json
        from pathlib import Path

        from safetensors.torch import load_file

        p = Path(path)
        if not p.exists():
            raise ValueError(f"Path {path} does not exist")
        if not p.is_file():
            raise ValueError(f"Path {path} is not a file")
        cfg = json.loads(p.read_text())
        known = FluxParams.__dataclass_fields__.keys()
        params = FluxParams(**{k: v for k, v in cfg["params"].items() if k in known})
        spec = FluxSpec(params=params, prequantized_flow=cfg.get("prequantized_flow", False),
                        quantize_modulation=cfg.get("quantize_modulation", True),
                        quantize_flow_embedder_layers=cfg.get("quantize_flow_embedder_layers", False))
        with torch.device("meta"):
            klass = cls(spec, dtype=dtype)
            if not spec.prequantized_flow:
                klass.type(dtype)
        ckpt = load_file(cfg["ckpt_path"], device="cpu")
        klass.load_state_dict(ckpt, assign=True)
        return klass.to("cpu")
