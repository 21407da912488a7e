"""CPU restatement of the reference's VAE decode (modules/autoencoder.py) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this; the product path
(flux-fp8-api_b200/autoencoder.py) never does and has no CPU fallback.

A functional restatement over a plain state dict (reference keys `decoder.*`), each function citing the reference
lines it follows.  Two arithmetic policies:

  policy="fp32"      every op in fp32: must reproduce the UNMODIFIED reference module run in fp32 (pinned by
                     oracle/make_golden.py against /root/reference, and by tests/test_oracle_golden.py against the
                     committed reference output).
  policy="autocast"  what the reference computes on the GPU inside `torch.autocast("cuda", torch.bfloat16)`
                     (flux_pipeline.py:431-434): conv2d and scaled_dot_product_attention take bf16 inputs and return bf16
                     (bias added to the rounded convolution output, as at::_convolution does around cuDNN); group_norm is
                     on autocast's fp32 list, so GroupNorm and the swish after it run in fp32 and are rounded by the next
                     convolution's input cast; residual sums are bf16.  CPU autocast has a different op list, so this
                     policy cannot be produced by running the reference on the CPU: it is pinned on the B200 instead
                     (tests/test_gpu_vae.py runs the staged reference under CUDA autocast next to it).
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F
from torch import Tensor

BF16 = torch.bfloat16


def _conv(x: Tensor, sd: Dict[str, Tensor], name: str, padding: int, policy: str) -> Tensor:
    w, b = sd[name + ".weight"], sd.get(name + ".bias")
    if policy == "fp32":
        return F.conv2d(x.float(), w.float(), None if b is None else b.float(), padding=padding)
    y = F.conv2d(x.to(BF16).float(), w.to(BF16).float(), None, padding=padding).to(BF16)  # fp32 accumulate, bf16 result
    if b is not None:
        y = (y.float() + b.to(BF16).float().view(1, -1, 1, 1)).to(BF16)
    return y


def _gn(x: Tensor, sd: Dict[str, Tensor], name: str) -> Tensor:
    """nn.GroupNorm(32, C, eps=1e-6, affine=True) (:27-29 / :62-70 / :245-247) in fp32."""
    return F.group_norm(x.float(), 32, sd[name + ".weight"].float(), sd[name + ".bias"].float(), eps=1e-6)


def swish(x: Tensor) -> Tensor:
    """:18-19"""
    return x * torch.sigmoid(x)


def resnet_block(x: Tensor, sd: Dict[str, Tensor], p: str, policy: str) -> Tensor:
    """ResnetBlock.forward :81-94"""
    h = _conv(swish(_gn(x, sd, p + ".norm1")), sd, p + ".conv1", 1, policy)
    h = _conv(swish(_gn(h, sd, p + ".norm2")), sd, p + ".conv2", 1, policy)
    if p + ".nin_shortcut.weight" in sd:
        x = _conv(x, sd, p + ".nin_shortcut", 0, policy)
    if policy == "fp32":
        return x + h
    return (x.to(BF16).float() + h.float()).to(BF16)


def attn_block(x: Tensor, sd: Dict[str, Tensor], p: str, policy: str) -> Tensor:
    """AttnBlock.attention / forward :38-53: single head, head dim = channels, softmax scale 1/sqrt(C)."""
    h = _gn(x, sd, p + ".norm")
    q, k, v = (_conv(h, sd, p + "." + n, 0, policy) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q, k, v = (t.reshape(b, c, hh * ww).transpose(1, 2).float() for t in (q, k, v))
    s = q @ k.transpose(1, 2) / math.sqrt(c)
    pr = torch.softmax(s, dim=-1)
    if policy != "fp32":
        pr = pr.to(BF16).float()
    o = pr @ v
    if policy != "fp32":
        o = o.to(BF16)
    o = o.transpose(1, 2).reshape(b, c, hh, ww)
    y = _conv(o, sd, p + ".proj_out", 0, policy)
    if policy == "fp32":
        return x + y
    return (x.to(BF16).float() + y.float()).to(BF16)


def decoder(z: Tensor, sd: Dict[str, Tensor], ch_mult: List[int], num_res_blocks: int, policy: str = "autocast",
            prefix: str = "decoder") -> Tensor:
    """Decoder.forward :256-283."""
    h = _conv(z, sd, prefix + ".conv_in", 1, policy)
    h = resnet_block(h, sd, prefix + ".mid.block_1", policy)
    h = attn_block(h, sd, prefix + ".mid.attn_1", policy)
    h = resnet_block(h, sd, prefix + ".mid.block_2", policy)
    for i_level in reversed(range(len(ch_mult))):
        for i_block in range(num_res_blocks + 1):
            h = resnet_block(h, sd, f"{prefix}.up.{i_level}.block.{i_block}", policy)
        if i_level != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")  # Upsample.forward :120-123
            h = _conv(h, sd, f"{prefix}.up.{i_level}.upsample.conv", 1, policy)
    h = swish(_gn(h, sd, prefix + ".norm_out"))
    return _conv(h, sd, prefix + ".conv_out", 1, policy)


def encoder(x: Tensor, sd: Dict[str, Tensor], ch_mult: List[int], num_res_blocks: int, policy: str = "autocast",
            prefix: str = "encoder") -> Tensor:
    """Encoder.forward :177-200 -> the Gaussian's moments [B, 2 z_channels, H/8, W/8]; Downsample.forward :106-110 pads
    (0, 1, 0, 1) and convolves 3x3 with stride 2."""
    h = _conv(x, sd, prefix + ".conv_in", 1, policy)
    for i_level in range(len(ch_mult)):
        for i_block in range(num_res_blocks):
            h = resnet_block(h, sd, f"{prefix}.down.{i_level}.block.{i_block}", policy)
        if i_level != len(ch_mult) - 1:
            name = f"{prefix}.down.{i_level}.downsample.conv"
            w, b = sd[name + ".weight"], sd[name + ".bias"]
            hp = F.pad(h.float() if policy == "fp32" else h.to(BF16).float(), (0, 1, 0, 1))
            if policy == "fp32":
                h = F.conv2d(hp, w.float(), b.float(), stride=2)
            else:
                h = (F.conv2d(hp, w.to(BF16).float(), None, stride=2).to(BF16).float() + b.to(BF16).float().view(1, -1, 1, 1)).to(BF16)
    h = resnet_block(h, sd, prefix + ".mid.block_1", policy)
    h = attn_block(h, sd, prefix + ".mid.attn_1", policy)
    h = resnet_block(h, sd, prefix + ".mid.block_2", policy)
    h = swish(_gn(h, sd, prefix + ".norm_out"))
    return _conv(h, sd, prefix + ".conv_out", 1, policy)


def decode(z: Tensor, sd: Dict[str, Tensor], ch_mult: List[int], num_res_blocks: int, scale_factor: float, shift_factor: float,
           policy: str = "autocast") -> Tensor:
    """AutoEncoder.decode :330-333 on an fp32 latent (flux_pipeline.py:430 hands over `x.float()`)."""
    return decoder(z.float() / scale_factor + shift_factor, sd, ch_mult, num_res_blocks, policy)


def synthetic_state(ref_ae_module, seed: int, dtype=torch.bfloat16, prefixes=("decoder.",)) -> Dict[str, Tensor]:
    """Seeded, better-conditioned-than-default parameters for a reference AutoEncoder (there are no weights to download):
    default Conv2d init, GroupNorm affine 1 + 0.1 N / 0.1 N, q / k projections x3 so the attention is not uniform.  Rounded
    to `dtype` (the reference keeps the VAE in bf16, util.py:287)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    full = ref_ae_module.state_dict()
    for k in sorted(full):  # sorted: the stream of random numbers does not depend on module construction order
        v = full[k]
        if not k.startswith(tuple(prefixes)):  # ("decoder.", "encoder."): the decoder's tensors come first and are unchanged
            continue
        t = v.detach().float().clone()
        if ".norm" in k and k.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
        elif ".norm" in k and k.endswith(".bias"):
            t = 0.1 * torch.randn(t.shape, generator=g)
        elif k.endswith(".weight"):
            fan_in = t[0].numel()
            t = torch.randn(t.shape, generator=g) * (1.0 / math.sqrt(fan_in))
            if ".attn_1.q." in k or ".attn_1.k." in k:
                t = t * 3.0
        else:
            t = 0.05 * torch.randn(t.shape, generator=g)
        sd[k] = t.to(dtype)
    return sd


def state_checksum(sd: Dict[str, Tensor]) -> float:
    """Order-independent fingerprint of a synthetic state (fixtures store it instead of 25 MB of parameters)."""
    return float(sum(v.double().abs().sum().item() * (1 + (i % 7)) for i, (k, v) in enumerate(sorted(sd.items()))))
