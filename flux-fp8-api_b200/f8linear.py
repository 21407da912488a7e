"""F8Linear and the model-quantisation helpers, B200-native.

Mirror of the reference's operator surface for this path (reference: float8_quantize.py) -- same
class / function names, constructor arguments, buffer names and state-dict keys, same exceptions --
so that reference-side code (Flux container, lora_loading, checkpoint loaders) can drive it
unchanged.  All device work goes through libflux_b200.so:

  forward            -> fluxb200_quantize (input) + fluxb200_f8_gemm / fluxb200_f8_gemv   (reference :272-296)
  quantize_weight    -> fluxb200_amax + fluxb200_quantize                                  (reference :195-207)
  quantize_input     -> fluxb200_amax (calibration calls) + fluxb200_quantize             (reference :220-246)

There is no torch._scaled_mm / CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn

from . import _cabi as cabi
from . import ops

__all__ = ["F8Linear", "recursive_swap_linears", "quantize_flow_transformer_and_dispatch_float8",
           "swap_to_cublaslinear", "CublasLinear"]

#: the reference optionally swaps leftover fp16 linears to aredden/torch-cublas-hgemm's CublasLinear
#: (float8_quantize.py:24-27, 372-392).  That extension is fp16-only and inactive for bf16 flows; it is
#: out of scope here and represented by the same "absent" sentinel the reference uses.
CublasLinear = type(None)

_SCALE_KEYS = ("scale", "input_scale", "scale_reciprocal", "input_scale_reciprocal")

#: How `x * scale` treats the 0-dim fp32 scale (reference float8_quantize.py:217-218).  The reference runs
#: on CUDA, where ATen's binary-op kernels cast a 0-dim *CUDA* tensor operand to the common dtype (bf16)
#: before the fp32 multiply -- i.e. the reference on a GPU quantises with bf16(scale) and de-quantises with
#: the exact fp32 reciprocal.  On CPU the 0-dim tensor is unwrapped as an fp32 scalar instead.  "cuda" (the
#: default) reproduces the GPU reference; "cpu" reproduces the CPU run that minted tests/golden/.
SCALE_SEMANTICS = "cuda"


def _version(t: torch.Tensor) -> int:
    try:
        return t._version
    except RuntimeError:  # inference tensors do not track versions
        return 0


def mul_scale(scale: torch.Tensor) -> torch.Tensor:
    """The value the quantisation kernels multiply by for a given F8Linear scale buffer."""
    if SCALE_SEMANTICS == "cuda":
        return scale.to(torch.bfloat16).to(torch.float32)
    return scale


class F8Linear(nn.Module):
    """Linear layer with an fp8 (e4m3) weight and per-tensor scaled fp8 (e5m2 by default) activations."""

    def __init__(
        self,
        in_features: int,
        out_features: int,
        bias: bool = True,
        device=None,
        dtype=torch.float16,
        float8_dtype=torch.float8_e4m3fn,
        float_weight: Optional[torch.Tensor] = None,
        float_bias: Optional[torch.Tensor] = None,
        num_scale_trials: int = 12,
        input_float8_dtype=torch.float8_e5m2,
    ) -> None:
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.float8_dtype, self.input_float8_dtype = float8_dtype, input_float8_dtype
        self.max_value = torch.finfo(float8_dtype).max
        self.input_max_value = torch.finfo(input_float8_dtype).max
        self.num_scale_trials = num_scale_trials
        self.weight_initialized = False
        self.input_scale_initialized = False
        self.trial_index = 0

        if float_weight is not None:
            self.weight = nn.Parameter(float_weight, requires_grad=float_weight.requires_grad)
        else:
            self.weight = nn.Parameter(torch.empty((out_features, in_features), dtype=dtype, device=device))
        if float_bias is not None:
            self.bias = nn.Parameter(float_bias, requires_grad=float_bias.requires_grad)
        elif bias:
            self.bias = nn.Parameter(torch.empty(out_features, dtype=dtype, device=device))
        else:
            self.register_parameter("bias", None)
        self.input_amax_trials = torch.zeros(num_scale_trials, dtype=torch.float32, device=device)
        for name in ("scale", "input_scale", "float8_data", "scale_reciprocal", "input_scale_reciprocal"):
            self.register_buffer(name, None)

    # ---- scale arithmetic (0-dim fp32 tensors; reference :214-218) -----------------------------------
    def amax_to_scale(self, amax, max_val):
        return (max_val / torch.clamp(amax, min=1e-12)).clamp(max=max_val)

    def to_fp8_saturated(self, x, scale, max_val):
        """Kept for surface compatibility: the un-cast saturated product.  The hot path never calls
        this -- fluxb200_quantize does product, clamp and cast in one pass."""
        return (x * scale).clamp(-max_val, max_val)

    # ---- weights ------------------------------------------------------------------------------------
    def quantize_weight(self):
        """amax -> scale (clamped to fp8 max) -> e4m3 bytes; `weight` becomes a zeros[1] placeholder whose
        dtype still defines the output dtype (reference :195-207)."""
        if self.weight_initialized:
            return
        w = self.weight.data
        cabi.require_cuda(w)
        w = w.contiguous()
        if w.dtype != torch.bfloat16:
            # fp16/fp32 masters: the kernels take bf16; the reference multiplies in the weight's own dtype.
            raise cabi.FluxB200Error(f"F8Linear on B200 expects bfloat16 master weights, got {w.dtype}")
        self.scale = self.amax_to_scale(ops.amax(w), self.max_value)
        self.float8_data = ops.quantize(w, mul_scale(self.scale), self.float8_dtype)
        self.scale_reciprocal = self.scale.reciprocal()
        self.weight.data = torch.zeros(1, dtype=w.dtype, device=w.device)
        self.weight_initialized = True

    def set_weight_tensor(self, tensor: torch.Tensor):
        self.weight.data = tensor
        self.weight_initialized = False
        self.quantize_weight()

    # ---- activations ---------------------------------------------------------------------------------
    def _freeze_or_track(self, x: torch.Tensor) -> None:
        """Dynamic -> static input-scale calibration (reference :220-246): the first `num_scale_trials`
        calls record max|x| and use the running max; the next call freezes the scale."""
        if self.trial_index < self.num_scale_trials:
            if self.input_amax_trials.device != x.device:
                self.input_amax_trials = self.input_amax_trials.to(x.device)
            self.input_amax_trials[self.trial_index] = ops.amax(x)
            self.trial_index += 1
            running = self.input_amax_trials[: self.trial_index].max()
        else:
            running = self.input_amax_trials.max()
            self.input_scale_initialized = True
        self.input_scale = self.amax_to_scale(running, self.input_max_value)
        self.input_scale_reciprocal = self.input_scale.reciprocal()

    @property
    def qscale(self) -> torch.Tensor:
        """input_scale as the quantising kernels consume it (see SCALE_SEMANTICS); cached once frozen.  The cache is
        keyed on the identity of the `input_scale` buffer and, where the tensor tracks one, its in-place version; code
        that rewrites the buffer in place under inference_mode calls blocks.invalidate_derived (broadcast_state does)."""
        if self.input_scale_initialized:
            c = self.__dict__.get("_qscale_cache")
            key = (id(self.input_scale), self.input_scale.data_ptr(), _version(self.input_scale), SCALE_SEMANTICS)
            if c is None or c[0] != key:
                c = (key, mul_scale(self.input_scale))
                self.__dict__["_qscale_cache"] = c
            return c[1]
        return mul_scale(self.input_scale)

    def quantize_input(self, x: torch.Tensor):
        if not self.input_scale_initialized:
            self._freeze_or_track(x)
        return ops.quantize(x, self.qscale, self.input_float8_dtype)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        cabi.require_cuda(x)
        xq = self.quantize_input(x)
        lead = xq.shape[:-1]
        out = ops.f8_gemm(xq.view(-1, self.in_features), self.float8_data, self.bias, self.input_scale_reciprocal,
                          self.scale_reciprocal)
        if out.dtype != self.weight.dtype:
            out = out.to(self.weight.dtype)
        return out.view(*lead, self.out_features)

    @property
    def frozen(self) -> bool:
        return self.weight_initialized and self.input_scale_initialized

    # ---- (de)serialisation: the "prequantised flow" checkpoint format (reference :91-193) -------------
    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        sd = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
        if "weight" not in sd:
            raise RuntimeError("Weight tensor not found or has incorrect shape in state dict")
        w = sd["weight"]
        full_shape = (self.out_features, self.in_features)
        f8 = sd.get("float8_data")
        if f8 is None:
            if tuple(w.shape) != full_shape:
                raise RuntimeError(f"Weight tensor not found or has incorrect shape in state dict: {sd.keys()}")
            # an un-quantised checkpoint: take the float weight and quantise it now
            self._parameters["weight"] = nn.Parameter(w, requires_grad=False)
            if "bias" in sd:
                self._parameters["bias"] = nn.Parameter(sd["bias"], requires_grad=False)
            self.weight_initialized = False
            self.quantize_weight()
            return
        if tuple(f8.shape) != full_shape or bool((w != 0).any()):
            raise RuntimeError(f"Weight tensor not found or has incorrect shape in state dict: {sd.keys()}")
        if f8.dtype not in (torch.float8_e4m3fn, torch.float8_e5m2):
            raise RuntimeError(f"float8_data has dtype {f8.dtype}, expected a float8 tensor: {sd.keys()}")
        if w.dtype != torch.bfloat16 or ("bias" in sd and sd["bias"].dtype != torch.bfloat16):
            # the kernels read bias as bf16 and write bf16; the placeholder's dtype defines out_dtype (reference
            # float8_quantize.py:204-206, 290).  A float16 flow would be silently mis-read: refuse it.
            raise cabi.FluxB200Error(
                f"prequantised F8Linear state with weight placeholder {w.dtype} / bias "
                f"{sd['bias'].dtype if 'bias' in sd else None}: the B200 path supports bfloat16 flows only "
                "(flow_dtype=bfloat16); re-save the checkpoint from a bfloat16 flow")
        self._buffers["float8_data"] = f8
        self._parameters["weight"] = nn.Parameter(torch.zeros(1, dtype=w.dtype, device=w.device), requires_grad=False)
        if "bias" in sd:
            self._parameters["bias"] = nn.Parameter(sd["bias"], requires_grad=False)
        self.weight_initialized = True
        if all(k in sd for k in _SCALE_KEYS):
            for k in _SCALE_KEYS:
                self._buffers[k] = sd[k].float()
            self.input_scale_initialized = True
            self.trial_index = self.num_scale_trials
            return
        # weight scales only (or nothing): input scale has to be re-calibrated
        if "scale" in sd and "scale_reciprocal" in sd:
            self._buffers["scale"] = sd["scale"].float()
            self._buffers["scale_reciprocal"] = sd["scale_reciprocal"].float()
            self._buffers["input_scale"] = sd["input_scale"].float() if "input_scale" in sd else None
            self._buffers["input_scale_reciprocal"] = (
                sd["input_scale_reciprocal"].float() if "input_scale_reciprocal" in sd else None)
        self.input_scale_initialized = False
        self.trial_index = 0
        self.input_amax_trials = torch.zeros(self.num_scale_trials, dtype=torch.float32, device=f8.device)

    def reset_parameters(self) -> None:
        if self.weight_initialized:
            self.weight = nn.Parameter(torch.empty((self.out_features, self.in_features), dtype=self.weight.dtype,
                                                   device=self.weight.device))
            self.weight_initialized = False
            self.input_scale_initialized = False
            self.trial_index = 0
            self.input_amax_trials.zero_()
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.in_features
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)
        self.quantize_weight()
        self.max_value = torch.finfo(self.float8_dtype).max
        self.input_max_value = torch.finfo(self.input_float8_dtype).max

    @classmethod
    def from_linear(cls, linear: nn.Linear, float8_dtype=torch.float8_e4m3fn,
                    input_float8_dtype=torch.float8_e5m2) -> "F8Linear":
        out = cls(
            in_features=linear.in_features,
            out_features=linear.out_features,
            bias=linear.bias is not None,
            device=linear.weight.device,
            dtype=linear.weight.dtype,
            float8_dtype=float8_dtype,
            float_weight=linear.weight.data,
            float_bias=None if linear.bias is None else linear.bias.data,
            input_float8_dtype=input_float8_dtype,
        )
        out.quantize_weight()
        return out

    def extra_repr(self) -> str:
        return (f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}, "
                f"weight={self.float8_dtype}, input={self.input_float8_dtype}, frozen={self.frozen}")


@torch.inference_mode()
def recursive_swap_linears(model: nn.Module, float8_dtype=torch.float8_e4m3fn, input_float8_dtype=torch.float8_e5m2,
                           quantize_modulation: bool = True, ignore_keys: List[str] = []) -> None:
    """In-place: every plain nn.Linear below `model` becomes an F8Linear (reference :320-369).  Modulation
    sub-trees are skipped when quantize_modulation is False."""
    from .blocks import Modulation

    for name, child in list(model.named_children()):
        if name in ignore_keys or (isinstance(child, Modulation) and not quantize_modulation):
            continue
        if isinstance(child, nn.Linear) and not isinstance(child, F8Linear):
            setattr(model, name, F8Linear.from_linear(child, float8_dtype=float8_dtype,
                                                      input_float8_dtype=input_float8_dtype))
        else:
            recursive_swap_linears(child, float8_dtype=float8_dtype, input_float8_dtype=input_float8_dtype,
                                   quantize_modulation=quantize_modulation, ignore_keys=ignore_keys)


def swap_to_cublaslinear(model: nn.Module):
    """fp16-only third-party HGEMM swap of the reference (float8_quantize.py:372-392): not applicable to the
    bf16 flows this path supports; kept as a no-op so callers need no changes."""
    return None


_EXTRAS = ("vector_in", "img_in", "txt_in", "time_in", "guidance_in", "final_layer", "pe_embedder")


@torch.inference_mode()
def quantize_flow_transformer_and_dispatch_float8(
    flow_model: nn.Module,
    device=torch.device("cuda"),
    float8_dtype=torch.float8_e4m3fn,
    input_float8_dtype=torch.float8_e5m2,
    offload_flow=False,
    swap_linears_with_cublaslinear=True,
    flow_dtype=torch.float16,
    quantize_modulation: bool = True,
    quantize_flow_embedder_layers: bool = True,
) -> nn.Module:
    """Block by block: move to `device`, swap linears to F8Linear, quantise (reference :395-496).  The
    embedders are quantised only when quantize_flow_embedder_layers is set; final_layer never is."""
    kw = dict(float8_dtype=float8_dtype, input_float8_dtype=input_float8_dtype, quantize_modulation=quantize_modulation)
    for stack in (flow_model.double_blocks, flow_model.single_blocks):
        for block in stack:
            block.to(device).eval()
            recursive_swap_linears(block, **kw)
    for name in _EXTRAS:
        extra = getattr(flow_model, name, None)
        if extra is None:
            continue
        extra.to(device).eval()
        if not quantize_flow_embedder_layers or name == "final_layer":
            continue
        if isinstance(extra, nn.Linear) and not isinstance(extra, F8Linear):
            setattr(flow_model, name, F8Linear.from_linear(extra, float8_dtype=float8_dtype,
                                                           input_float8_dtype=input_float8_dtype))
        else:
            recursive_swap_linears(extra, **kw)
    if swap_linears_with_cublaslinear and flow_dtype != torch.float16:
        pass  # the reference warns and skips here as well (float8_quantize.py:491-492)
    if offload_flow:
        raise cabi.FluxB200Error("offload_flow is not supported: the B200 path keeps the 12 GB of fp8 weights resident")
    return flow_model
