"""Per-launch breakdown of one VAE decode (Flux's VAE, synthetic weights), CUDA events around every launch of ours, and the
end-to-end time next to the staged reference under CUDA autocast:  python tools/vae_profile.py [resolution=1024] [batch=1]"""
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flux_fp8_api_b200 import autoencoder as A, ops  # noqa: E402
from oracle import ref_loader as R, vae_oracle as V  # noqa: E402  (diagnostics: the reference arm)

res = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
params = dict(resolution=256, in_channels=3, ch=128, out_ch=3, ch_mult=[1, 2, 4, 4], num_res_blocks=2, z_channels=16,
              scale_factor=0.3611, shift_factor=0.1159)
m = A.AutoEncoder(A.AutoEncoderParams(**params))
sd = V.synthetic_state(m, seed=77)
m.load_state_dict(sd, strict=False)
m = m.to("cuda", torch.bfloat16).eval()
z = torch.randn(B, 16, res // 8, res // 8, device="cuda") * 1.2


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


with torch.inference_mode():
    ms = timed(lambda: m.decode(z))
    print(f"ours: {ms:.2f} ms per decode ({B} x {res}x{res})")
    ops.KERNEL_TIMELINE = []
    REPS = 3
    for _ in range(REPS):
        m.decode(z)
    torch.cuda.synchronize()
    tl, ops.KERNEL_TIMELINE = ops.KERNEL_TIMELINE, None
    agg = OrderedDict()
    for kind, work, s, e, detail in tl:
        a = agg.setdefault((kind, detail), [0.0, 0.0, 0])
        a[0] += work
        a[1] += s.elapsed_time(e)
        a[2] += 1
    tot = sum(a[1] for a in agg.values()) / REPS
    print(f"{'kernel':14s} {'shape':36s} {'n':>4s} {'us each':>9s} {'ms':>8s} {'share':>6s} {'TFLOP/s':>8s}")
    for (kind, detail), (work, t, n) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{kind:14s} {detail:36s} {n // REPS:4d} {t / n * 1e3:9.1f} {t / REPS:8.3f} {t / REPS / tot * 100:5.1f}% {work / (t * 1e-3) / 1e12:8.1f}")
    by_kind = OrderedDict()
    for (kind, _), (work, t, n) in agg.items():
        k = by_kind.setdefault(kind, [0.0, 0.0, 0])
        k[0] += work; k[1] += t; k[2] += n
    for kind, (work, t, n) in sorted(by_kind.items(), key=lambda kv: -kv[1][1]):
        print(f"  {kind:14s} {n // REPS:4d} launches {t / REPS:8.3f} ms  {work / (t * 1e-3) / 1e12 if work else 0:7.1f} TFLOP/s")
    print(f"sum of timed launches: {tot:.2f} ms")
    if R.available() and R.load().ae is not None:
        ref = R.load()
        rae = ref.ae.AutoEncoder(ref.ae.AutoEncoderParams(**params))
        rae.load_state_dict(sd, strict=False)
        rae = rae.to("cuda", torch.bfloat16).eval()

        def run_ref():
            with torch.autocast(device_type="cuda", dtype=torch.bfloat16, cache_enabled=False):
                return rae.decode(z)

        print(f"reference (cuDNN, autocast bf16): {timed(run_ref, 3):.2f} ms per decode")
