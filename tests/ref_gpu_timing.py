"""Context measurement (not a bench line): the reference's EAGER fp8 path on the same B200.

/root/reference does not exist on the GPU box, so the reference's module code cannot run there; what can run
is the oracle's restatement of it with BACKEND="library", i.e. the very PyTorch entry points the reference
calls per op (torch._scaled_mm(use_fast_accum=True) = cuBLASLt fp8, F.scaled_dot_product_attention = torch's
fused SDPA, F.layer_norm, eager elementwise) in the same order with the same intermediate tensors.  That is
"the number to beat" of SURVEY.md section 8(d): same GPU, same torch, same weights, same shapes.

    python tests/ref_gpu_timing.py [depth_double depth_single]     (default: full 19 + 38)

Prints it/s of the eager library path, of this repo's path (CUDA-graph step), and the parity spread between
the two on one Flux.forward.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flux_fp8_api_b200 import model as M, pipeline as PL  # noqa: E402
from oracle import flux_oracle as O  # noqa: E402

dd, ds = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (19, 38)
dev = torch.device("cuda", 0)
params = M.FluxParams(depth=dd, depth_single_blocks=ds)
spec = M.FluxSpec(params=params)
net = PL.build_synthetic_flux(spec, dev)
req = PL.synthetic_request(params, 1024, 1024, 1, 512, dev, seed=0)
PL.calibrate(net, req, num_steps=13)
sd = {k: v for k, v in net.state_dict().items() if v is not None}
cfg = dict(num_heads=24, depth=dd, depth_single_blocks=ds, axes_dim=[16, 56, 56], theta=10_000, guidance_embed=True)
sched = PL.get_schedule(28, req["img"].shape[1])
t = torch.full((1,), sched[0], dtype=torch.bfloat16, device=dev)


def time_ms(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


with torch.inference_mode():
    O.BACKEND = "library"
    lib_fwd = lambda: O.flux_forward(sd, cfg, req["img"], req["img_ids"], req["txt"], req["txt_ids"], t, req["y"], req["guidance"])  # noqa: E731
    ref = lib_fwd()
    ms_lib = time_ms(lib_fwd)
    O.BACKEND = "restated"
    sess = PL.DenoiseSession(net, req, use_graph=True)
    ms_ours = time_ms(lambda: sess.step_device(req["img"], sched[0], sched[1]), warm=3, reps=10)
    ours = net(img=req["img"], img_ids=req["img_ids"], txt=req["txt"], txt_ids=req["txt_ids"], timesteps=t, y=req["y"],
               guidance=req["guidance"])
    e = (ours.float() - ref.float()).abs()
    # the floor at this depth: the oracle's restated arithmetic (fp32 accumulation, fp32 P) against the library
    # kernels (cuBLASLt fast-accum fp8, fused SDPA) -- two executions of the reference's own algorithm
    rest = O.flux_forward(sd, cfg, req["img"], req["img_ids"], req["txt"], req["txt_ids"], t, req["y"], req["guidance"])
    f = (rest.float() - ref.float()).abs()
    g = (ours.float() - rest.float()).abs()
    out = {"depth": [dd, ds], "library_eager_ms_per_step": ms_lib, "library_eager_it_s": 1000.0 / ms_lib,
           "ours_ms_per_step": ms_ours, "ours_it_s": 1000.0 / ms_ours, "speedup": ms_lib / ms_ours,
           "parity_vs_library": {"mean": e.mean().item(), "max": e.max().item(), "ref_amax": ref.abs().max().item(),
                                 "ref_rms": ref.float().pow(2).mean().sqrt().item()},
           "floor_restated_vs_library": {"mean": f.mean().item(), "max": f.max().item()},
           "parity_vs_restated": {"mean": g.mean().item(), "max": g.max().item()}}
    print(json.dumps(out))
