"""Run-to-run bit reproducibility of the VAE decode and of its kernels (diagnostics): python tools/vae_determinism.py"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flux_fp8_api_b200 import autoencoder as A, ops, pipeline as PL  # noqa: E402

BF16 = torch.bfloat16
dev = "cuda"


def rnd(shape, seed, scale=1.0):
    g = torch.Generator(device=dev).manual_seed(seed)
    return (torch.randn(shape, device=dev, generator=g) * scale).to(BF16)


def check(name, fn, n=6):
    ref = fn()
    bad = 0
    for _ in range(n):
        out = fn()
        outs = out if isinstance(out, (tuple, list)) else (out,)
        refs = ref if isinstance(ref, (tuple, list)) else (ref,)
        if not all(torch.equal(a, b) for a, b in zip(outs, refs)):
            bad += 1
            d = max((a.double() - b.double()).abs().max().item() for a, b in zip(outs, refs))
    print(f"{name:60s} {'REPRODUCIBLE' if bad == 0 else f'{bad}/{n} runs differ (max |d| {d:.3g})'}", flush=True)


for (B, H, W, Cin, N, k, res, stats) in [(2, 8, 8, 256, 256, 3, True, True), (2, 16, 16, 256, 256, 3, True, False),
                                         (2, 64, 64, 64, 128, 3, False, True), (1, 128, 128, 512, 512, 3, True, True),
                                         (1, 256, 256, 128, 128, 3, True, True), (2, 8, 8, 256, 256, 1, True, True)]:
    x = rnd((B, H, W, Cin), 1)
    w = ops.pack_conv_weight(rnd((N, Cin, k, k), 2, 1.0 / math.sqrt(Cin * k * k)))
    b = rnd((N,), 3, 0.1)
    r = rnd((B, H, W, N), 4) if res else None

    def run():
        st = torch.empty((B, 32, 2), dtype=torch.float64, device=dev) if stats else None
        y = ops.conv2d_nhwc(x, w, b, k * k, residual=r, gn_stats=st)
        return (y, st) if stats else y

    check(f"conv {B}x{H}x{W} {Cin}->{N} k{k} res={res} stats={stats}", run)

x = rnd((2, 64, 64, 128), 5, 2.0)
gm, bt = rnd((128,), 6) * 0.1 + 1, rnd((128,), 7, 0.1)
check("group_norm (own statistics)", lambda: ops.group_norm_nhwc(x, gm, bt, 1e-6, True))
sc = torch.randn(300, 1024, device=dev)
check("softmax_rows", lambda: ops.softmax_rows(sc))

for p in (dict(resolution=64, in_channels=3, ch=64, out_ch=3, ch_mult=[1, 2, 4, 4], num_res_blocks=2, z_channels=16, scale_factor=0.3611,
               shift_factor=0.1159),):
    m = A.AutoEncoder(A.AutoEncoderParams(**p))
    PL.init_synthetic_vae_weights(m, seed=31)
    m = m.to(dev, BF16).eval()
    z = torch.randn(2, 16, 8, 8, device=dev)
    with torch.inference_mode():
        check("tiny decoder, whole decode", lambda: m.decode(z))
        h = ops.vae_latent_prep(z, 0.3611, 0.1159)
        check("tiny decoder: conv_in", lambda: A._conv(m.decoder.conv_in, m.decoder._pin, h, want_stats=True))
        h1, s1 = A._conv(m.decoder.conv_in, m.decoder._pin, h, want_stats=True)
        check("tiny decoder: mid.block_1", lambda: m.decoder.mid.block_1.forward_nhwc(h1, s1))
        h2, s2 = m.decoder.mid.block_1.forward_nhwc(h1, s1)
        check("tiny decoder: mid.attn_1", lambda: m.decoder.mid.attn_1.forward_nhwc(h2, s2))
        check("tiny decoder: mid.attn_1.attention", lambda: m.decoder.mid.attn_1.attention_nhwc(h2, s2))
