"""Denoise-loop orchestration around Flux.forward (reference: flux_pipeline.py:234-344, 347-371, 627-651).

The reference's FluxPipeline.generate keeps the loop in Python and this stays so (BASELINE.json:
"flux_pipeline.denoise stays as orchestrator"); text encoders, VAE and JPEG encoding are out of scope
(SURVEY.md section 2).  What lives here:

* the pieces of `prepare` / `get_noise` / `get_schedule` that shape the transformer's inputs
* `denoise()`: the Euler loop of generate() (flux_pipeline.py:627-651)
* `GraphedStep`: one whole Flux.forward + Euler update captured in a CUDA graph (static shapes, static
  buffers), replayed once per step -- the B200 replacement for the reference's per-block torch.compile
* synthetic weights / inputs of the published Flux shapes (there is no checkpoint or network here)
"""
from __future__ import annotations

import contextlib
import math
from typing import Callable, Dict, List, Optional

import torch
from torch import Tensor, nn

from . import ops
from .f8linear import F8Linear, quantize_flow_transformer_and_dispatch_float8
from .model import Flux, FluxSpec

BF16 = torch.bfloat16


# ------------------------------------------------------------------------------------------------
# schedule / input shaping
# ------------------------------------------------------------------------------------------------
def time_shift(mu: float, sigma: float, t: Tensor) -> Tensor:
    return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)


def get_schedule(num_steps: int, image_seq_len: int, base_shift: float = 0.5, max_shift: float = 1.15,
                 shift: bool = True) -> List[float]:
    """num_steps+1 timesteps from 1 to 0, shifted towards high noise for large images (dev only)."""
    ts = torch.linspace(1, 0, num_steps + 1)
    if shift:
        slope = (max_shift - base_shift) / (4096 - 256)
        mu = slope * image_seq_len + (base_shift - slope * 256)
        ts = time_shift(mu, 1.0, ts)
    return ts.tolist()


def patchify(latent: Tensor) -> Tensor:
    """[B,16,h,w] latent -> [B,(h/2)(w/2),64] tokens of 2x2 patches (flux_pipeline.py:270-271)."""
    x = latent.unfold(2, 2, 2).unfold(3, 2, 2).permute(0, 2, 3, 1, 4, 5)
    return x.reshape(x.shape[0], -1, x.shape[3] * x.shape[4] * x.shape[5])


def unpack(x: Tensor, height: int, width: int) -> Tensor:
    """[B, (h w), (c 2 2)] denoised tokens -> [B, c, 2h, 2w] latent (flux_pipeline.py:440-448: the inverse of patchify;
    h = ceil(height / 16), w = ceil(width / 16))."""
    h, w = math.ceil(height / 16), math.ceil(width / 16)
    B, L, D = x.shape
    if L != h * w or D % 4:
        raise ValueError(f"unpack: {tuple(x.shape)} is not a {h} x {w} token grid of 2x2 patches")
    c = D // 4
    return x.reshape(B, h, w, c, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(B, c, 2 * h, 2 * w)


def vae_decode(ae, x: Tensor, height: int, width: int) -> Tensor:
    """FluxPipeline.vae_decode (flux_pipeline.py:422-438): tokens -> fp32 latent -> `ae.decode` (this package's AutoEncoder: the
    autocast region of the reference is what its kernels implement) -> bf16 image [B, 3, height, width] in [-1, 1]."""
    with torch.inference_mode():
        return ae.decode(unpack(x.float(), height, width).contiguous())


def init_synthetic_vae_weights(ae: nn.Module, seed: int = 77, dtype=BF16, prefixes=("decoder.",)) -> None:
    """Seeded parameters for an AutoEncoder (there are no weights to download): N(0, 1/fan_in) convolutions, GroupNorm affine
    1 + 0.1 N / 0.1 N, biases 0.05 N, the attention's q / k projections x3 so its softmax is not uniform.  Keys are visited in
    sorted order, so the same seed gives the same tensors to any module with the reference's decoder keys."""
    g = torch.Generator().manual_seed(seed)
    full = ae.state_dict()
    new = {}
    for k in sorted(full):
        if not k.startswith(tuple(prefixes)):
            continue
        shape = full[k].shape
        if ".norm" in k and k.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif ".norm" in k and k.endswith(".bias"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif k.endswith(".weight"):
            t = torch.randn(shape, generator=g) * (1.0 / math.sqrt(full[k][0].numel()))
            if ".attn_1.q." in k or ".attn_1.k." in k:
                t = t * 3.0
        else:
            t = 0.05 * torch.randn(shape, generator=g)
        new[k] = t.to(dtype)
    ae.load_state_dict(new, strict=False)


def make_img_ids(batch: int, h2: int, w2: int, device, dtype=BF16) -> Tensor:
    """Position ids (0, row, col) of the token grid (flux_pipeline.py:280-292)."""
    ids = torch.zeros(h2, w2, 3, device=device, dtype=dtype)
    ids[..., 1] = ids[..., 1] + torch.arange(h2, device=device, dtype=dtype)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(w2, device=device, dtype=dtype)[None, :]
    return ids[None].repeat(batch, 1, 1, 1).flatten(1, 2)


def synthetic_request(params, height: int, width: int, batch: int, text_len: int, device, seed: int = 0,
                      guidance: float = 3.5, sample_offset: int = 0) -> Dict[str, Tensor]:
    """Latents and text embeddings of the shapes generate() feeds the transformer, from per-sample seeds
    (sample i uses seed + sample_offset + i, so a batch shard reproduces the same samples)."""
    h8, w8 = 2 * math.ceil(height / 16), 2 * math.ceil(width / 16)
    lat, txt, y = [], [], []
    for i in range(batch):
        g = torch.Generator(device=device).manual_seed(seed + sample_offset + i)
        lat.append(torch.randn(1, 16, h8, w8, device=device, dtype=BF16, generator=g))
        txt.append(0.15 * torch.randn(1, text_len, params.context_in_dim, device=device, dtype=torch.float32, generator=g))
        y.append(torch.randn(1, params.vec_in_dim, device=device, dtype=torch.float32, generator=g))
    img = patchify(torch.cat(lat))
    return dict(
        img=img.contiguous(),
        img_ids=make_img_ids(batch, h8 // 2, w8 // 2, device),
        txt=torch.cat(txt).to(BF16),
        txt_ids=torch.zeros(batch, text_len, 3, device=device, dtype=BF16),
        y=torch.cat(y).to(BF16),
        guidance=torch.full((batch,), guidance, device=device, dtype=BF16),
    )


# ------------------------------------------------------------------------------------------------
# synthetic model
# ------------------------------------------------------------------------------------------------
@torch.inference_mode()
def init_synthetic_weights(model: nn.Module, seed: int = 1234) -> None:
    """Seeded stand-in for the BFL checkpoint: Linear W ~ N(0, 0.02^2), bias
    ~ N(0, 0.02^2), modulation W ~ N(0, 0.01^2) / bias 0, QK-norm scale = 1 + N(0, 0.05^2).
    Generated layer by layer on the model's device in fp32, stored bf16."""
    from .blocks import Modulation, RMSNorm

    mod_lins = {id(m.lin) for m in model.modules() if isinstance(m, Modulation)}
    dev = next(model.parameters()).device
    g = torch.Generator(device=dev).manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.Linear):
            is_mod = id(m) in mod_lins
            m.weight.copy_(torch.randn(m.weight.shape, device=dev, generator=g) * (0.01 if is_mod else 0.02))
            if m.bias is not None:
                if is_mod:
                    m.bias.zero_()
                else:
                    m.bias.copy_(torch.randn(m.bias.shape, device=dev, generator=g) * 0.02)
        elif isinstance(m, RMSNorm):
            m.scale.copy_(1 + 0.05 * torch.randn(m.scale.shape, device=dev, generator=g))


@torch.inference_mode()
def build_synthetic_flux(spec: FluxSpec, device, seed: int = 1234, quantize: bool = True) -> Flux:
    """bf16 Flux with synthetic weights on `device`, then the reference flow: swap linears to F8Linear and
    quantise (quantize_flow_transformer_and_dispatch_float8).  Input scales still need calibrate()."""
    spec_build = FluxSpec(params=spec.params, prequantized_flow=False, quantize_modulation=spec.quantize_modulation,
                          quantize_flow_embedder_layers=spec.quantize_flow_embedder_layers)
    with torch.device(device):
        model = Flux(spec_build, dtype=BF16).to(BF16)
    init_synthetic_weights(model, seed)
    model.eval()
    if quantize:
        quantize_flow_transformer_and_dispatch_float8(
            model, torch.device(device), offload_flow=False, swap_linears_with_cublaslinear=False, flow_dtype=BF16,
            quantize_modulation=spec.quantize_modulation,
            quantize_flow_embedder_layers=spec.quantize_flow_embedder_layers)
    return model


# ------------------------------------------------------------------------------------------------
# prequantised checkpoints.  The reference can LOAD a prequantised flow
# (`prequantized_flow: true` + F8Linear._load_from_state_dict, float8_quantize.py:91-193) but has no
# writer; this is the writer, and the artefact `parallel.broadcast_state` ships to the other ranks.
# ------------------------------------------------------------------------------------------------
PREQUANTIZED_FORMAT = "flux-fp8-b200/prequantized"
PREQUANTIZED_VERSION = 2  # 1 = round-1 torch.save pickle (still readable); 2 = safetensors


def _spec_to_dict(spec: FluxSpec) -> dict:
    import dataclasses

    d = dataclasses.asdict(spec)
    d["params"] = {k: (list(v) if isinstance(v, tuple) else v) for k, v in d["params"].items()}
    return d


def _spec_from_dict(sd: dict) -> FluxSpec:
    from .model import FluxParams

    pd = dict(sd["params"])
    pd["axes_dim"] = list(pd["axes_dim"])
    return FluxSpec(**{**sd, "params": FluxParams(**pd)})


def save_prequantized(model: Flux, path: str, spec: Optional[FluxSpec] = None) -> dict:
    """Write the quantised state (e4m3 `float8_data`, the four 0-dim fp32 scales, bf16 biases / norms / embedders,
    zeros[1] `weight` placeholders) exactly as `state_dict()` names it, as ONE .safetensors file -- the container and
    the key layout the reference loads (`util.load_flow_model` -> `load_sft(ckpt_path)` -> `load_state_dict(...,
    assign=True)`, util.py:249-255, with `prequantized_flow: true`), so the file is a valid `ckpt_path` for the
    reference as well.  Format / version / model spec travel in the safetensors metadata.  Returns the header."""
    import json

    from safetensors.torch import save_file

    if not all_frozen(model):
        raise RuntimeError("save_prequantized: input scales are not frozen yet (run calibrate() first); a checkpoint "
                           "without them would silently re-calibrate on its first 12 steps")
    state = {k: v.detach().to("cpu").contiguous() for k, v in model.state_dict().items() if v is not None}
    n_f8 = sum(1 for k in state if k.endswith(".float8_data"))
    header = {"format": PREQUANTIZED_FORMAT, "version": PREQUANTIZED_VERSION, "f8_layers": n_f8,
              "bytes": int(sum(v.numel() * v.element_size() for v in state.values())),
              "spec": _spec_to_dict(spec) if spec is not None else None}
    save_file(state, path, metadata={k: json.dumps(v) for k, v in header.items()})
    return header


def read_prequantized(path: str):
    """(header, state dict on the CPU) of a save_prequantized file (safetensors; the round-1 pickle is still read)."""
    import json

    try:
        from safetensors import safe_open

        with safe_open(path, framework="pt", device="cpu") as f:
            meta = f.metadata() or {}
            state = {k: f.get_tensor(k) for k in f.keys()}
        header = {k: json.loads(v) for k, v in meta.items()}
    except Exception as ex:  # noqa: BLE001  (not a safetensors container: the version-1 pickle)
        try:
            blob = torch.load(path, map_location="cpu", weights_only=False)
        except Exception:
            raise RuntimeError(f"{path}: neither a safetensors file nor a version-1 checkpoint ({ex})") from ex
        header, state = blob.get("header", {}), blob.get("state", {})
    if header.get("format") != PREQUANTIZED_FORMAT:
        raise RuntimeError(f"{path}: not a {PREQUANTIZED_FORMAT} file")
    if header.get("version", 0) > PREQUANTIZED_VERSION:
        raise RuntimeError(f"{path}: format version {header.get('version')} is newer than this build ({PREQUANTIZED_VERSION})")
    return header, state


def load_prequantized(path: str, device, spec: Optional[FluxSpec] = None) -> Flux:
    """Build a Flux with F8Linear layers in place (prequantized_flow) and load a `save_prequantized` file: no master
    weights, no quantisation pass, no calibration -- the model is frozen and graph-capturable straight away."""
    header, state = read_prequantized(path)
    if spec is None:
        sd = header.get("spec")
        if sd is None:
            raise ValueError("load_prequantized: the file carries no model spec; pass spec=")
        spec = _spec_from_dict(sd)
    spec = FluxSpec(**{**spec.__dict__, "prequantized_flow": True})
    with torch.device(device):
        model = Flux(spec, dtype=BF16)
    state = {k: v.to(device) for k, v in state.items()}
    model.load_state_dict(state, strict=True, assign=True)
    model.eval()
    if not all_frozen(model):
        raise RuntimeError(f"{path}: some F8Linear layers came back without frozen scales")
    return model


def set_input_float8_dtype(model: nn.Module, dtype: torch.dtype) -> None:
    """Activation format of every F8Linear of a model built with prequantized_flow=True.  The prequantised state dict
    does not record it (the reference constructs those layers with the e5m2 default, modules/flux_model.py:279-345,
    whatever `input_float8_dtype` the checkpoint was quantised and calibrated with, float8_quantize.py:298-304), so a
    checkpoint calibrated for e4m3 activations needs this after loading."""
    from .blocks import invalidate_derived

    for m in model.modules():
        if isinstance(m, F8Linear):
            m.input_float8_dtype = dtype
            m.input_max_value = torch.finfo(dtype).max
    invalidate_derived(model)


def all_frozen(model: nn.Module) -> bool:
    return all(m.frozen for m in model.modules() if isinstance(m, F8Linear))


@torch.inference_mode()
def calibrate(model: Flux, request: Dict[str, Tensor], num_steps: int = 13, shift: bool = True) -> None:
    """Freeze every F8Linear input scale the way the reference's warm-up generate does
    (flux_pipeline.py:197-212): run `num_steps` >= 13 denoise steps in eager mode."""
    denoise(model, dict(request), get_schedule(num_steps, request["img"].shape[1], shift=shift))
    if not all_frozen(model):
        raise RuntimeError("calibration did not freeze every F8Linear input scale (need > num_scale_trials calls)")


# ------------------------------------------------------------------------------------------------
# denoise loop
# ------------------------------------------------------------------------------------------------
@torch.inference_mode()
def denoise(model: Callable, request: Dict[str, Tensor], timesteps: List[float],
            step_fn: Optional[Callable] = None) -> Tensor:
    """Euler integration of the flow (flux_pipeline.py:627-651): per step t_vec.fill_(t_curr);
    pred = model(...); img = img + (t_prev - t_curr) * pred."""
    img = request["img"]
    t_vec = None
    for t_curr, t_prev in zip(timesteps[:-1], timesteps[1:]):
        if t_vec is None:
            t_vec = torch.full((img.shape[0],), t_curr, dtype=img.dtype, device=img.device)
        else:
            t_vec = t_vec.reshape((img.shape[0],)).fill_(t_curr)
        if step_fn is not None:
            img = step_fn(img, t_vec, t_prev - t_curr)
            continue
        pred = model(img=img, img_ids=request["img_ids"], txt=request["txt"], txt_ids=request["txt_ids"],
                     y=request["y"], timesteps=t_vec, guidance=request.get("guidance"))
        img = img + (t_prev - t_curr) * pred
    return img


class GraphedStep:
    """One denoise step (Flux.forward + Euler update) as a CUDA graph over static buffers.

    Every kernel of the step -- ours through the C ABI and the few torch ops of the embedders -- is captured once for
    a fixed (batch, L, T) and replayed per step; TMA descriptors are encoded at capture time and baked into the kernel
    parameters.  Per step the host passes the two scalars (t, dt) as kernel arguments of two fill kernels; the latent stays in the
    graph's static buffer between steps (the graph's last node copies the updated latent back over its input).

    Ownership: the graph's kernels read per-request tensors that Flux computes once (txt_in(txt), the vector and
    guidance embeddings, pe, cos/sin).  They live in a cache PRIVATE to this object (Flux.use_request_cache), so no
    other request, session or LoRA swap on the same model can free them under a later replay.  A LoRA that changes the
    weights below those embeddings bumps Flux._invariant_epoch; the next call re-captures."""

    def __init__(self, model: Flux, request: Dict[str, Tensor], warmup: int = 2):
        if not all_frozen(model):
            raise RuntimeError("GraphedStep needs frozen input scales: run calibrate() first")
        self.model = model
        self.req = request
        self.warmup = warmup
        dev = request["img"].device
        self.img = request["img"].clone()
        self.t_vec = torch.zeros((self.img.shape[0],), dtype=self.img.dtype, device=dev)
        #: [t as the bf16 value the reference's t_vec.fill_ would store, dt = t_prev - t_curr]
        self.scal = torch.zeros((2,), dtype=torch.float32, device=dev)
        self.out = torch.empty_like(self.img)
        self.captures = 0
        self._capture()

    @property
    def dt(self) -> Tensor:
        return self.scal[1]

    def _capture(self):
        from .model import _StepInvariantCache

        dev = self.img.device
        self.cache = _StepInvariantCache()  # strong references to everything step-invariant the kernels read
        self.epoch = getattr(self.model, "_invariant_epoch", 0)
        keep = self.img.clone()
        stream = torch.cuda.Stream(device=dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        private = self.model.use_request_cache(self.cache) if hasattr(self.model, "use_request_cache") else contextlib.nullcontext()
        with torch.cuda.stream(stream), torch.inference_mode(), private:
            for _ in range(self.warmup):
                self._step()
                self.img.copy_(keep)
        torch.cuda.current_stream(dev).wait_stream(stream)
        self.graph = torch.cuda.CUDAGraph()
        private = self.model.use_request_cache(self.cache) if hasattr(self.model, "use_request_cache") else contextlib.nullcontext()
        with torch.inference_mode(), private, torch.cuda.graph(self.graph, stream=stream):
            self._step()
        self.img.copy_(keep)
        self.captures += 1

    def _step(self):
        r = self.req
        self.t_vec.copy_(self.scal[0].expand_as(self.t_vec))  # exact: scal[0] already holds a bf16 value
        # img + (t_prev - t_curr) * pred: eager torch multiplies in fp32 by the python scalar, rounds the product to
        # bf16, then adds in bf16 -- done inside the final projection's launch with the fp32 0-dim `dt`
        if hasattr(self.model, "denoise_step"):
            self.model.denoise_step(img=self.img, img_ids=r["img_ids"], txt=r["txt"], txt_ids=r["txt_ids"], y=r["y"],
                                    timesteps=self.t_vec, guidance=r.get("guidance"), dt=self.scal[1], out=self.out)
        else:  # another container over the B200 blocks (e.g. the reference's own Flux, reference_binding replace_container=False)
            pred = self.model(img=self.img, img_ids=r["img_ids"], txt=r["txt"], txt_ids=r["txt_ids"], y=r["y"],
                              timesteps=self.t_vec, guidance=r.get("guidance"))
            if pred.dtype == BF16:
                ops.euler_update(self.img, pred, self.scal[1], out=self.out)
            else:
                self.out.copy_(self.img + (self.scal[1] * pred.float()).to(pred.dtype))
        self.img.copy_(self.out)  # the next step's input, unless the caller supplies another latent

    @staticmethod
    def _bf16_value(t: float) -> float:
        return float(torch.tensor(t, dtype=torch.bfloat16))

    def advance(self, t_curr: float, dt: float, latent: Optional[Tensor] = None, clone: bool = True) -> Tensor:
        """One step from the latent held in the static buffer (or `latent`, copied in first).  With clone=False the
        returned tensor IS the static output buffer: valid until the next call."""
        if self.epoch != getattr(self.model, "_invariant_epoch", 0):
            self._capture()
        if latent is not None and latent is not self.out:
            self.img.copy_(latent)
        # the two scalars travel as kernel arguments of two fill kernels (a pinned staging buffer rewritten by the host
        # every step would race with the previous step's still-queued copy)
        self.scal[0:1].fill_(self._bf16_value(t_curr))
        self.scal[1:2].fill_(dt)
        self.graph.replay()
        return self.out.clone() if clone else self.out

    def __call__(self, img: Tensor, t_vec: Tensor, dt: float, clone: bool = True) -> Tensor:
        """denoise(step_fn=...) signature: t_vec is the reference's filled bf16 timestep vector."""
        if self.epoch != getattr(self.model, "_invariant_epoch", 0):
            self._capture()
        if img is not self.out:
            self.img.copy_(img)
        self.scal[0:1].copy_(t_vec[0:1])
        self.scal[1:2].fill_(dt)
        self.graph.replay()
        return self.out.clone() if clone else self.out


class DenoiseSession:
    """Public per-request API: owns the (graph-captured) step for one request shape and moves latents
    between pinned host memory and the device.

        sess = DenoiseSession(model, request)              # captures the CUDA graph
        latent = sess.run(timesteps)                        # whole loop, latents stay in HBM
        host_out = sess.step_host(host_in, t_curr, t_prev)  # one step, host -> device -> host
    """

    def __init__(self, model: Flux, request: Dict[str, Tensor], use_graph: bool = True):
        self.model, self.request = model, request
        self.step = GraphedStep(model, request) if use_graph else None
        img = request["img"]
        self._host_in = torch.empty(img.shape, dtype=img.dtype, pin_memory=True)
        self._host_out = torch.empty(img.shape, dtype=img.dtype, pin_memory=True)
        self._dev_in = torch.empty_like(img)
        self._t_vec = torch.zeros((img.shape[0],), dtype=img.dtype, device=img.device)
        self._t_host = torch.zeros((img.shape[0],), dtype=img.dtype, pin_memory=True)
        self._dt_host = torch.zeros((), dtype=torch.float32, pin_memory=True)

    @property
    def h2d_bytes_per_step(self) -> int:
        latent = self._host_in.numel() * self._host_in.element_size()
        if self.step is not None:
            return latent  # (t, dt) travel as kernel arguments, not as a copy
        return latent + self._t_host.numel() * self._t_host.element_size()

    @property
    def d2h_bytes_per_step(self) -> int:
        return self._host_out.numel() * self._host_out.element_size()

    @torch.inference_mode()
    def run(self, timesteps: List[float]) -> Tensor:
        """The whole Euler loop (flux_pipeline.py:627-651) with the latent resident in HBM."""
        if self.step is None:
            return denoise(self.model, self.request, timesteps)
        g, img = self.step, self.request["img"]
        for t_curr, t_prev in zip(timesteps[:-1], timesteps[1:]):
            img = g.advance(t_curr, t_prev - t_curr, latent=img, clone=False)
        return img.clone()

    @torch.inference_mode()
    def step_device(self, img: Tensor, t_curr: float, t_prev: float) -> Tensor:
        if self.step is not None:
            return self.step.advance(t_curr, t_prev - t_curr, latent=img)
        self._t_vec.fill_(t_curr)
        r = self.request
        pred = self.model(img=img, img_ids=r["img_ids"], txt=r["txt"], txt_ids=r["txt_ids"], y=r["y"],
                          timesteps=self._t_vec, guidance=r.get("guidance"))
        return img + (t_prev - t_curr) * pred

    @torch.inference_mode()
    def step_host(self, img_host: Tensor, t_curr: float, t_prev: float) -> Tensor:
        """One denoise step with HOST buffers: pinned H2D copy of the latent and timestep, the step on the
        device, D2H copy of the updated latent (synchronises before returning)."""
        self._host_in.copy_(img_host)
        if self.step is not None:
            # straight into / out of the graph's static buffers: 2 H2D copies, one graph launch, one D2H copy
            g = self.step
            g.img.copy_(self._host_in, non_blocking=True)
            out = g.advance(t_curr, t_prev - t_curr, clone=False)
        else:
            self._t_host.fill_(t_curr)
            self._dev_in.copy_(self._host_in, non_blocking=True)
            self._t_vec.copy_(self._t_host, non_blocking=True)
            r = self.request
            pred = self.model(img=self._dev_in, img_ids=r["img_ids"], txt=r["txt"], txt_ids=r["txt_ids"], y=r["y"],
                              timesteps=self._t_vec, guidance=r.get("guidance"))
            out = self._dev_in + (t_prev - t_curr) * pred
        self._host_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._host_out
