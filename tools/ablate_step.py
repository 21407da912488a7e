"""Marginal cost of each kernel family INSIDE the captured step (diagnostics; results of ablated runs are garbage).

The per-kernel CUDA-event shares of bench.py come from an eager-launch pass, where host gaps let the power-capped GPU
clock up.  This tool measures what each family really costs in the back-to-back CUDA-graph step: the step is re-captured
with one family removed (its launches skipped, outputs left uninitialised) and the difference to the full step is the
family's marginal time -- including its share of the power budget.

    python tools/ablate_step.py [config]        (default c2)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from flux_fp8_api_b200 import _cabi as cabi, model as M, ops, pipeline as PL  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = bench.CONFIGS[key]
dev = torch.device("cuda", 0)
spec = (M.flux_dev_spec if cfg["guidance"] else M.flux_schnell_spec)(quantize_modulation=cfg["qmod"])
net = PL.build_synthetic_flux(spec, dev, seed=1234)
req = PL.synthetic_request(spec.params, cfg["res"], cfg["res"], 1, cfg["text_len"], dev, seed=0)
if not cfg["guidance"]:
    req["guidance"] = None
L = req["img"].shape[1]
if cfg["guidance"]:
    PL.calibrate(net, req, num_steps=13)
else:
    for _ in range(4):
        PL.denoise(net, dict(req), PL.get_schedule(4, L, shift=False))
sched = PL.get_schedule(28, L, shift=cfg["guidance"])


def step_ms(steps=20, warm=5):
    sess = PL.DenoiseSession(net, req, use_graph=True)
    img = req["img"]
    for i in range(warm):
        img = sess.step.advance(sched[i], sched[i + 1] - sched[i], latent=img, clone=False)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        img = sess.step.advance(sched[i % 28], sched[i % 28 + 1] - sched[i % 28], latent=img, clone=False)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps


results = {}
results["full"] = step_ms()
print(f"full step: {results['full']:.3f} ms", flush=True)
ops.FUSE_LN_INTO_GEMM = True
results["ln_fused_into_gemm_launch"] = step_ms()
ops.FUSE_LN_INTO_GEMM = False
print(f"LayerNorm fused into the GEMM launches: {results['ln_fused_into_gemm_launch']:.3f} ms", flush=True)

lib = cabi.load()
real = {n: getattr(lib, n) for n in ("fluxb200_attention", "fluxb200_ln_mod_quant", "fluxb200_ln_mod_quant_grouped",
                                     "fluxb200_modulation_batched", "fluxb200_f8_gemm", "fluxb200_f8_gemm_grouped")}


class Skip:
    def __init__(self, *names):
        self.names = names

    def __enter__(self):
        for n in self.names:
            setattr(lib, n, lambda *a, **k: 0)

    def __exit__(self, *exc):
        for n in self.names:
            setattr(lib, n, real[n])


for label, names in (("attention", ("fluxb200_attention",)),
                     ("ln_mod_quant", ("fluxb200_ln_mod_quant", "fluxb200_ln_mod_quant_grouped")),
                     ("modulation", ("fluxb200_modulation_batched",)),
                     ("gemm", ("fluxb200_f8_gemm", "fluxb200_f8_gemm_grouped"))):
    with Skip(*names):
        ms = step_ms()
    results[f"without_{label}"] = ms
    print(f"without {label:14s}: {ms:8.3f} ms   -> marginal {results['full'] - ms:7.3f} ms", flush=True)

for mode, label in ((4, "gemm_no_epilogue"), (2, "gemm_no_mma")):
    cabi.check(lib.fluxb200_gemm_probe_mode(mode), "probe")
    try:
        ms = step_ms()
    finally:
        cabi.check(lib.fluxb200_gemm_probe_mode(0), "probe")
    results[label] = ms
    print(f"{label:22s}: {ms:8.3f} ms   -> saves {results['full'] - ms:7.3f} ms", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(results, open(f"gpurun_out/ablate_{key}.json", "w"), indent=1)
