"""LoRA hot-swap for the quantised flow -- host mirror of the reference's lora_loading.py fuse / unfuse entry
points, with the per-layer arithmetic on the device (fluxb200_lora_fuse + fluxb200_quantize).

Mirrors (same names, argument meaning and order):
    get_module_for_key            lora_loading.py:466-473
    get_lora_for_key              lora_loading.py:476-497
    apply_lora_to_model           lora_loading.py:634-692
    remove_lora_from_module       lora_loading.py:695-754
    LoraWeights                   lora_loading.py:22-33
The LoRA state dict must already use the BFL key layout (`double_blocks.0.img_attn.qkv.lora_A.weight` ...): the
diffusers -> BFL key conversion (lora_loading.py:36-461) is string manipulation off the hot path and stays with the
reference (out of scope).

Differences by design: the quantised layer's buffers (`float8_data`, `scale`, `scale_reciprocal`) are updated IN
PLACE, so a captured CUDA graph / ModulationBank over the model keeps working after a swap (pipeline.GraphedStep owns
the per-request tensors its kernels read; a LoRA on txt_in / vector_in / guidance_in bumps Flux._invariant_epoch and
the graph re-captures on its next call); no full-size fp32 temporaries are made (the reference makes five per layer).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor, nn

from . import ops
from .f8linear import F8Linear, mul_scale


class LoraWeights:
    def __init__(self, weights: Dict[str, Tensor], path: str, name: Optional[str] = None, scale: float = 1.0) -> None:
        self.path = path
        self.weights = weights
        self.name = name if name else path.replace("\\", "/").split("/")[-1]
        self.scale = scale


def get_module_for_key(key: str, model: nn.Module) -> nn.Module:
    module = model
    for part in key.split("."):
        module = getattr(module, part)
    return module


def get_lora_for_key(key: str, lora_weights: dict) -> Optional[Tuple[Tensor, Tensor, Optional[float]]]:
    prefix = key.split(".lora")[0]
    lora_A = lora_weights.get(f"{prefix}.lora_A.weight")
    lora_B = lora_weights.get(f"{prefix}.lora_B.weight")
    alpha = lora_weights.get(f"{prefix}.alpha")
    if lora_A is None or lora_B is None:
        return None
    return lora_A, lora_B, alpha


def _keys_without_ab(lora_weights: dict):
    return list({key.replace(".lora_A.weight", "").replace(".lora_B.weight", "").replace(".lora_A", "")
                 .replace(".lora_B", "").replace(".alpha", "") for key in lora_weights.keys()})


def lora_operands(lora_sd, rank: Optional[int], device) -> Tuple[Tensor, Tensor, int]:
    """The fp32 factors of calculate_lora_weight (lora_loading.py:509-541): (lora_B [N,r], lora_A [c*r,K], chunks),
    lora_A pre-multiplied by alpha / rank when they differ -- two fp32 roundings, as the reference does it."""
    lora_A, lora_B, alpha = lora_sd
    uneven = lora_B.shape[1] != lora_A.shape[0]
    rank_diff = lora_A.shape[0] / lora_B.shape[1]
    if rank is None:
        rank = lora_B.shape[1]
    if alpha is None:
        alpha = rank
    if isinstance(alpha, Tensor):
        alpha = alpha.item()
    up = lora_A.to(dtype=torch.float32, device=device)
    down = lora_B.to(dtype=torch.float32, device=device)
    if alpha != rank:
        up = up * alpha / rank
    chunks = int(rank_diff) if uneven else 1
    if chunks * lora_B.shape[1] != lora_A.shape[0]:
        raise ValueError(f"LoRA factors do not chain: lora_B {tuple(lora_B.shape)}, lora_A {tuple(lora_A.shape)}")
    return down, up, chunks


@torch.inference_mode()
def fuse_into_f8linear(module: F8Linear, lora_sd, lora_scale: float = 1.0, unfuse: bool = False,
                       rank: Optional[int] = None) -> None:
    """One F8Linear of apply_lora_to_model / remove_lora_from_module, in place on the device."""
    if not module.weight_initialized:
        raise RuntimeError("fuse_into_f8linear: the layer's weight is not quantised yet")
    dev = module.float8_data.device
    down, up, chunks = lora_operands(lora_sd, rank, dev)
    w_new, amax = ops.lora_fuse(module.float8_data, module.scale_reciprocal, down, up, lora_scale, unfuse, chunks)
    # set_weight_tensor -> quantize_weight (float8_quantize.py:195-207) on the fused bf16 weight, written over the
    # existing buffers so that device pointers captured elsewhere stay valid
    scale = module.amax_to_scale(amax, module.max_value)
    ops.quantize(w_new, mul_scale(scale), module.float8_dtype, out=module.float8_data)
    module.scale.copy_(scale)
    module.scale_reciprocal.copy_(scale.reciprocal())


@torch.inference_mode()
def _fuse_into_linear(module: nn.Linear, lora_sd, lora_scale: float, unfuse: bool) -> None:
    """Un-quantised layers (embedders, final layer): the reference arithmetic in torch on the device, in place."""
    down, up, chunks = lora_operands(lora_sd, None, module.weight.device)
    delta = torch.zeros((down.shape[0], up.shape[1]), dtype=torch.float32, device=down.device)
    for c in up.chunk(chunks, dim=0):
        delta = delta + (lora_scale * torch.mm(down, c))
    w = module.weight.data.float()
    module.weight.data.copy_(((w - delta) if unfuse else (w + delta)).to(module.weight.dtype))


#: sub-modules whose outputs Flux caches per request (model._StepInvariantCache): a LoRA touching them changes values a
#: captured CUDA graph has baked in, so holders of such state must re-derive it (Flux.invalidate_step_invariants bumps
#: the epoch pipeline.GraphedStep checks).  Every other layer is updated in place and needs nothing.
_STEP_INVARIANT_PREFIXES = ("txt_in", "vector_in", "guidance_in")


def _walk(model: nn.Module, lora_weights, lora_scale: float, unfuse: bool):
    if isinstance(lora_weights, LoraWeights):
        if unfuse:
            lora_scale = lora_weights.scale
        lora_weights = lora_weights.weights
    touched_invariants = False
    for key in _keys_without_ab(lora_weights):
        module = get_module_for_key(key, model)
        lora_sd = get_lora_for_key(key, lora_weights)
        if lora_sd is None:
            continue
        if isinstance(module, F8Linear):
            fuse_into_f8linear(module, lora_sd, lora_scale, unfuse)
        elif isinstance(module, nn.Linear):
            _fuse_into_linear(module, lora_sd, lora_scale, unfuse)
        else:
            raise TypeError(f"{key}: cannot fuse a LoRA into {type(module).__name__}")
        touched_invariants |= key.split(".")[0] in _STEP_INVARIANT_PREFIXES
    if touched_invariants and hasattr(model, "invalidate_step_invariants"):
        model.invalidate_step_invariants()
    return model


def _resolve(lora_path):
    """A loaded BFL-layout state dict, a LoraWeights, or the path of a .safetensors file holding one.  (The diffusers
    / kohya -> BFL key conversion of lora_loading.py:36-461 is string manipulation off the hot path and stays with the
    reference: convert once with it, save, load here.)"""
    if isinstance(lora_path, str):
        from safetensors.torch import load_file

        sd = load_file(lora_path, device="cpu")
        bad = [k for k in sd if not (k.endswith(".lora_A.weight") or k.endswith(".lora_B.weight") or k.endswith(".alpha"))
               or k.startswith(("transformer.", "lora_unet_"))]
        if bad:
            raise ValueError(f"{lora_path}: not in the BFL key layout (e.g. {bad[0]!r}); convert it with the reference's "
                             "lora_loading.convert_* helpers first")
        return sd
    return lora_path


def apply_lora_to_model(model: nn.Module, lora_path, lora_scale: float = 1.0, return_lora_resolved: bool = False):
    """lora_loading.py:634-692.  `lora_path`: BFL-layout state dict, LoraWeights, or a .safetensors path."""
    lora = _resolve(lora_path)
    model = _walk(model, lora, lora_scale, unfuse=False)
    if return_lora_resolved:
        return model, (lora.weights if isinstance(lora, LoraWeights) else lora)
    return model


def remove_lora_from_module(model: nn.Module, lora_path, lora_scale: float = 1.0):
    """lora_loading.py:695-754."""
    return _walk(model, _resolve(lora_path), lora_scale, unfuse=True)
