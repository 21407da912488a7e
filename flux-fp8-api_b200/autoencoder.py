"""The VAE on the B200 kernels: the reference's `modules/autoencoder.py`.

Same class names, constructor signatures, sub-module names and state-dict keys as the reference (`AttnBlock` :22-50,
`ResnetBlock` :53-94, `Downsample` :97-110, `Upsample` :112-123, `Encoder` :126-200, `Decoder` :203-283, `AutoEncoder`
:300-337), so a reference `ae.safetensors` loads exactly as `util.py:283-286` does it.  The `nn.Conv2d` / `nn.GroupNorm`
children are parameter containers only: every forward runs through `libflux_b200.so` (`fluxb200_conv2d_nhwc`,
`fluxb200_group_norm_nhwc`, `fluxb200_upsample2x_nhwc`, `fluxb200_softmax_rows`, `fluxb200_vae_latent_prep`) on
channels-last bf16 activations; there is no torch fallback.

Numerics follow what the reference computes under `torch.autocast("cuda", torch.bfloat16)` (flux_pipeline.py:431-434):
convolutions and attention take bf16 inputs, accumulate in fp32 and round to bf16 (bias added after the rounding, as
`at::_convolution` does around cuDNN); GroupNorm and swish are evaluated in fp32 and rounded once, by the consuming
convolution's input cast; residual sums are bf16 + bf16.

The encoder half (`Encoder` :126-200, `Downsample` :97-110, `DiagonalGaussian` :286-298; image -> latent, used by img2img,
flux_pipeline.py:489-500) runs on the same kernels: the stride-2 convolution of `Downsample` is the implicit GEMM with a
TMA box that is traversed with element stride 2; the Gaussian sample and `scale_factor * (z - shift_factor)` stay eager
torch (a [B, 16, H/8, W/8] tensor, and `torch.randn_like` has to be torch's generator to match the reference's draw).
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
from pydantic import BaseModel
from torch import Tensor, nn

from . import ops
from .blocks import tensor_version

BF16 = torch.bfloat16


class AutoEncoderParams(BaseModel):
    """modules/autoencoder.py:7-16"""

    resolution: int
    in_channels: int
    ch: int
    out_ch: int
    ch_mult: List[int]
    num_res_blocks: int
    z_channels: int
    scale_factor: float
    shift_factor: float


def _to_nhwc(x: Tensor) -> Tensor:
    return x.to(BF16).permute(0, 2, 3, 1).contiguous()


def _to_nchw(x: Tensor) -> Tensor:
    return x.permute(0, 3, 1, 2).contiguous()


class _Packed:
    """Kernel-layout copy of a Conv2d's parameters, rebuilt when the parameter storage or version changes
    (load_state_dict, .to(), in-place edits)."""

    def __init__(self) -> None:
        self.key = None
        self.w: Optional[Tensor] = None
        self.b: Optional[Tensor] = None

    def get(self, conv: nn.Conv2d):
        w, b = conv.weight, conv.bias
        # (inference tensors carry no version counter: a state loaded under inference_mode is keyed on storage alone and has
        # to be re-assigned, not edited in place, to be picked up -- `invalidate()` forces a re-pack)
        key = (w.data_ptr(), tensor_version(w), w.device, w.dtype, None if b is None else (b.data_ptr(), tensor_version(b)))
        if key != self.key:
            self.w = ops.pack_conv_weight(w)
            self.b = None if b is None else b.detach().to(BF16).contiguous()
            self.key = key
        return self.w, self.b

    def invalidate(self) -> None:
        self.key = None


def invalidate_packed(module: nn.Module) -> None:
    """Drop every cached kernel-layout weight below `module` (after editing Conv2d parameters in place under inference_mode)."""
    for m in module.modules():
        for v in vars(m).values():
            if isinstance(v, _Packed):
                v.invalidate()


def _conv(conv: nn.Conv2d, cache: _Packed, x: Tensor, residual: Optional[Tensor] = None, out_mode: int = 0,
          want_stats: bool = False):
    """The convolution through the C ABI.  want_stats: also return the fp64 GroupNorm sums of the output, accumulated by the
    epilogue (None when the shape is outside what the fused path takes; the GroupNorm then reduces by itself)."""
    w, b = cache.get(conv)
    taps = conv.kernel_size[0] * conv.kernel_size[1]
    stats = None
    if want_stats and out_mode == 0 and ops.conv_can_fuse_gn_stats(x.shape[0], w.shape[0]):
        stats = torch.empty((x.shape[0], 32, 2), dtype=torch.float64, device=x.device)
    y = ops.conv2d_nhwc(x, w, b, taps, residual=residual, out_mode=out_mode, gn_stats=stats, stride=conv.stride[0])
    return (y, stats) if want_stats else y


def _norm(gn: nn.GroupNorm, x: Tensor, swish: bool, stats: Optional[Tensor] = None) -> Tensor:
    if gn.num_groups != 32:
        raise ValueError("the GroupNorm kernel is built for 32 groups (modules/autoencoder.py:27, :62, :68, :245)")
    return ops.group_norm_nhwc(x, gn.weight.detach().to(BF16), gn.bias.detach().to(BF16), gn.eps, swish, stats=stats)


class AttnBlock(nn.Module):
    """modules/autoencoder.py:22-50: GroupNorm -> q, k, v (1x1) -> single-head attention over the H*W positions
    (head dim = channels) -> proj_out (1x1) -> + x."""

    #: upper bound of the fp32 score block held at a time (elements): 64 Mi = 256 MB
    SCORE_ELEMS = 1 << 26

    def __init__(self, in_channels: int):
        super().__init__()
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)
        self.q = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.k = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.v = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self._pq, self._pk, self._pv, self._pp = _Packed(), _Packed(), _Packed(), _Packed()

    def attention_nhwc(self, x: Tensor, stats: Optional[Tensor] = None) -> Tensor:
        B, H, W, Cn = x.shape
        S = H * W
        if S % 4:
            raise ValueError(f"AttnBlock: H*W = {S} must be a multiple of 4")
        Sp = ((S + 63) // 64) * 64  # K extent of the P @ V product (zero padded)
        h = _norm(self.norm, x, swish=False, stats=stats)
        q = _conv(self.q, self._pq, h)
        k = _conv(self.k, self._pk, h)
        wv, bv = self._pv.get(self.v)
        vt = torch.zeros((B, Cn, Sp), dtype=BF16, device=x.device) if Sp != S else torch.empty((B, Cn, S), dtype=BF16, device=x.device)
        ops.conv2d_nhwc(h, wv, bv, 1, out=vt, out_mode=2, nchw_plane=Sp)  # v^T: [B, C, S], the B operand of P @ V
        o = torch.empty((B, H, W, Cn), dtype=BF16, device=x.device)
        rows_max = max(128, min(S, (self.SCORE_ELEMS // S) // 128 * 128))
        scale = 1.0 / math.sqrt(Cn)  # F.scaled_dot_product_attention's default, head dim = channels
        scores = torch.empty((rows_max, S), dtype=torch.float32, device=x.device)
        p = torch.zeros((rows_max, Sp), dtype=BF16, device=x.device)
        for b in range(B):
            qb, kb, ob = q[b].reshape(S, Cn), k[b].reshape(S, Cn), o[b].reshape(S, Cn)
            for r0 in range(0, S, rows_max):
                rows = min(rows_max, S - r0)
                sc = scores[:rows]
                ops.conv2d_nhwc(qb[r0:r0 + rows].view(1, 1, rows, Cn), kb, None, 1, out=sc.view(1, 1, rows, S), out_mode=1,
                                alpha=scale)
                ops.softmax_rows(sc, out=p[:rows, :S] if Sp == S else p[:rows].as_strided((rows, S), (Sp, 1)))
                ops.conv2d_nhwc(p[:rows].view(1, 1, rows, Sp), vt[b], None, 1, out=ob[r0:r0 + rows].view(1, 1, rows, Cn))
        return o

    def forward_nhwc(self, x: Tensor, stats: Optional[Tensor] = None):
        """(x, GroupNorm sums of x or None) -> (x + proj_out(attention(x)), sums of the result)."""
        return _conv(self.proj_out, self._pp, self.attention_nhwc(x, stats), residual=x, want_stats=True)

    def attention(self, h_: Tensor) -> Tensor:
        return _to_nchw(self.attention_nhwc(_to_nhwc(h_)))

    def forward(self, x: Tensor) -> Tensor:
        return _to_nchw(self.forward_nhwc(_to_nhwc(x))[0])


class ResnetBlock(nn.Module):
    """modules/autoencoder.py:53-94"""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.norm1 = nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.norm2 = nn.GroupNorm(num_groups=32, num_channels=out_channels, eps=1e-6, affine=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if self.in_channels != self.out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)
        self._p1, self._p2, self._ps = _Packed(), _Packed(), _Packed()

    def forward_nhwc(self, x: Tensor, stats: Optional[Tensor] = None):
        """(x, GroupNorm sums of x or None) -> (block output, its sums): every convolution that feeds a GroupNorm leaves
        that GroupNorm's statistics behind, so the normalisation reads its input once."""
        h = _norm(self.norm1, x, swish=True, stats=stats)
        h, hs = _conv(self.conv1, self._p1, h, want_stats=True)
        h = _norm(self.norm2, h, swish=True, stats=hs)
        if self.in_channels != self.out_channels:
            x = _conv(self.nin_shortcut, self._ps, x)
        return _conv(self.conv2, self._p2, h, residual=x, want_stats=True)

    def forward(self, x: Tensor) -> Tensor:
        return _to_nchw(self.forward_nhwc(_to_nhwc(x))[0])


class Upsample(nn.Module):
    """modules/autoencoder.py:112-123"""

    def __init__(self, in_channels: int):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)
        self._p = _Packed()

    def forward_nhwc(self, x: Tensor, stats: Optional[Tensor] = None):
        return _conv(self.conv, self._p, ops.upsample2x_nhwc(x), want_stats=True)

    def forward(self, x: Tensor) -> Tensor:
        return _to_nchw(self.forward_nhwc(_to_nhwc(x))[0])


class Downsample(nn.Module):
    """modules/autoencoder.py:97-110: F.pad(x, (0, 1, 0, 1)) then a 3x3 stride-2 convolution without padding."""

    def __init__(self, in_channels: int):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)
        self._p = _Packed()

    def forward_nhwc(self, x: Tensor, stats: Optional[Tensor] = None):
        return _conv(self.conv, self._p, x, want_stats=True)  # the kernel's stride-2 form pads right / bottom by itself

    def forward(self, x: Tensor) -> Tensor:
        return _to_nchw(self.forward_nhwc(_to_nhwc(x))[0])


class Encoder(nn.Module):
    """modules/autoencoder.py:126-200"""

    def __init__(self, resolution: int, in_channels: int, ch: int, ch_mult: List[int], num_res_blocks: int, z_channels: int):
        super().__init__()
        self.ch = ch
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.conv_in = nn.Conv2d(in_channels, self.ch, kernel_size=3, stride=1, padding=1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.in_ch_mult = in_ch_mult
        self.down = nn.ModuleList()
        block_in = self.ch
        for i_level in range(self.num_resolutions):
            block = nn.ModuleList()
            attn = nn.ModuleList()
            block_in = ch * in_ch_mult[i_level]
            block_out = ch * ch_mult[i_level]
            for _ in range(self.num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out))
                block_in = block_out
            down = nn.Module()
            down.block = block
            down.attn = attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in)
                curr_res = curr_res // 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in)
        self.norm_out = nn.GroupNorm(num_groups=32, num_channels=block_in, eps=1e-6, affine=True)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels, kernel_size=3, stride=1, padding=1)
        self._pin, self._pout = _Packed(), _Packed()

    def forward_nhwc(self, h: Tensor) -> Tensor:
        """h: bf16 [B, H, W, 64-padded image channels] -> bf16 NCHW moments [B, 2 z_channels, H/8, W/8]."""
        h, st = _conv(self.conv_in, self._pin, h, want_stats=True)
        for i_level in range(self.num_resolutions):
            for i_block in range(self.num_res_blocks):
                h, st = self.down[i_level].block[i_block].forward_nhwc(h, st)
                if len(self.down[i_level].attn) > 0:
                    h, st = self.down[i_level].attn[i_block].forward_nhwc(h, st)
            if i_level != self.num_resolutions - 1:
                h, st = self.down[i_level].downsample.forward_nhwc(h)
        h, st = self.mid.block_1.forward_nhwc(h, st)
        h, st = self.mid.attn_1.forward_nhwc(h, st)
        h, st = self.mid.block_2.forward_nhwc(h, st)
        h = _norm(self.norm_out, h, swish=True, stats=st)
        return _conv(self.conv_out, self._pout, h, out_mode=2)

    def forward(self, x: Tensor) -> Tensor:
        c = self.conv_in.in_channels
        if x.dim() != 4 or x.shape[1] != c:
            raise ValueError(f"Encoder: x must be [B, {c}, H, W], got {tuple(x.shape)}")
        return self.forward_nhwc(ops.vae_latent_prep(x.float().contiguous(), 1.0, 0.0, cpad=((c + 63) // 64) * 64))


class DiagonalGaussian(nn.Module):
    """modules/autoencoder.py:286-298 (eager torch: a tiny tensor, and the sample has to come from torch's generator)."""

    def __init__(self, sample: bool = True, chunk_dim: int = 1):
        super().__init__()
        self.sample = sample
        self.chunk_dim = chunk_dim

    def forward(self, z: Tensor) -> Tensor:
        mean, logvar = torch.chunk(z, 2, dim=self.chunk_dim)
        if self.sample:
            std = torch.exp(0.5 * logvar)
            return mean + std * torch.randn_like(mean)
        return mean


class Decoder(nn.Module):
    """modules/autoencoder.py:203-283"""

    def __init__(self, ch: int, out_ch: int, ch_mult: List[int], num_res_blocks: int, in_channels: int, resolution: int,
                 z_channels: int):
        super().__init__()
        self.ch = ch
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.ffactor = 2 ** (self.num_resolutions - 1)
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = nn.Conv2d(z_channels, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block = nn.ModuleList()
            attn = nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(self.num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out))
                block_in = block_out
            up = nn.Module()
            up.block = block
            up.attn = attn
            if i_level != 0:
                up.upsample = Upsample(block_in)
                curr_res = curr_res * 2
            self.up.insert(0, up)
        self.norm_out = nn.GroupNorm(num_groups=32, num_channels=block_in, eps=1e-6, affine=True)
        self.conv_out = nn.Conv2d(block_in, out_ch, kernel_size=3, stride=1, padding=1)
        self._pin, self._pout = _Packed(), _Packed()

    def forward_nhwc(self, h: Tensor) -> Tensor:
        """h: bf16 [B, H, W, 64-padded z channels] -> bf16 NCHW image [B, out_ch, 8H, 8W]."""
        h, st = _conv(self.conv_in, self._pin, h, want_stats=True)
        h, st = self.mid.block_1.forward_nhwc(h, st)
        h, st = self.mid.attn_1.forward_nhwc(h, st)
        h, st = self.mid.block_2.forward_nhwc(h, st)
        for i_level in reversed(range(self.num_resolutions)):
            for i_block in range(self.num_res_blocks + 1):
                h, st = self.up[i_level].block[i_block].forward_nhwc(h, st)
                if len(self.up[i_level].attn) > 0:
                    h, st = self.up[i_level].attn[i_block].forward_nhwc(h, st)
            if i_level != 0:
                h, st = self.up[i_level].upsample.forward_nhwc(h)
        h = _norm(self.norm_out, h, swish=True, stats=st)
        return _conv(self.conv_out, self._pout, h, out_mode=2)

    def forward(self, z: Tensor) -> Tensor:
        zc = self.conv_in.in_channels
        if z.dim() != 4 or z.shape[1] != zc:
            raise ValueError(f"Decoder: z must be [B, {zc}, H, W], got {tuple(z.shape)}")
        return self.forward_nhwc(ops.vae_latent_prep(z.float().contiguous(), 1.0, 0.0, cpad=((zc + 63) // 64) * 64))


class AutoEncoder(nn.Module):
    """modules/autoencoder.py:300-337"""

    def __init__(self, params: AutoEncoderParams):
        super().__init__()
        self.encoder = Encoder(resolution=params.resolution, in_channels=params.in_channels, ch=params.ch, ch_mult=params.ch_mult,
                               num_res_blocks=params.num_res_blocks, z_channels=params.z_channels)
        self.decoder = Decoder(resolution=params.resolution, in_channels=params.in_channels, ch=params.ch, out_ch=params.out_ch,
                               ch_mult=params.ch_mult, num_res_blocks=params.num_res_blocks, z_channels=params.z_channels)
        self.reg = DiagonalGaussian()
        self.scale_factor = params.scale_factor
        self.shift_factor = params.shift_factor

    def encode(self, x: Tensor) -> Tensor:
        """:325-328: `reg(encoder(x))` then `scale_factor * (z - shift_factor)`; x [B, 3, H, W] in [-1, 1]."""
        z = self.reg(self.encoder(x))
        return self.scale_factor * (z - self.shift_factor)

    def decode(self, z: Tensor) -> Tensor:
        """`z / scale_factor + shift_factor` (:331) then the decoder; z [B, z_channels, H, W] -> bf16 [B, out_ch, 8H, 8W]."""
        zc = self.decoder.conv_in.in_channels
        if z.dim() != 4 or z.shape[1] != zc:
            raise ValueError(f"AutoEncoder.decode: z must be [B, {zc}, H, W], got {tuple(z.shape)}")
        cpad = ((zc + 63) // 64) * 64
        if z.dtype != torch.float32:  # the pipeline always hands over fp32 (flux_pipeline.py:430); keep torch's roundings otherwise
            z = (z / self.scale_factor + self.shift_factor).float()
            return self.decoder.forward_nhwc(ops.vae_latent_prep(z.contiguous(), 1.0, 0.0, cpad=cpad))
        return self.decoder.forward_nhwc(ops.vae_latent_prep(z.contiguous(), self.scale_factor, self.shift_factor, cpad=cpad))

    def forward(self, x: Tensor) -> Tensor:
        return self.decode(self.encode(x))
