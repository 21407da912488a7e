"""Per-shape breakdown of one Flux-dev 1024x1024 step, timed IN the step (eager launches, CUDA events around every
launch of ours, GPU warm and power-capped as in the bench):  python tools/step_shapes.py"""
import os
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flux_fp8_api_b200 import model as M, ops, pipeline as PL  # noqa: E402

dev = torch.device("cuda", 0)
spec = M.flux_dev_spec()
net = PL.build_synthetic_flux(spec, dev)
req = PL.synthetic_request(spec.params, 1024, 1024, 1, 512, dev, seed=0)
PL.calibrate(net, req, num_steps=13)
sess = PL.DenoiseSession(net, req, use_graph=False)
sched = PL.get_schedule(28, req["img"].shape[1])
ops.KERNEL_TIMELINE = []
REPS = 6
with torch.inference_mode():
    for i in range(REPS + 2):
        if i == 2:
            ops.KERNEL_TIMELINE.clear()
        sess.step_device(req["img"], sched[0], sched[1])
torch.cuda.synchronize()
tl, ops.KERNEL_TIMELINE = ops.KERNEL_TIMELINE, None
agg = OrderedDict()
for kind, work, s, e, detail in tl:
    a = agg.setdefault((kind, detail), [0.0, 0.0, 0])
    a[0] += work
    a[1] += s.elapsed_time(e)
    a[2] += 1
tot = sum(a[1] for a in agg.values()) / REPS
print(f"{'kernel':14s} {'shape':60s} {'n/step':>6s} {'us each':>9s} {'ms/step':>8s} {'share':>6s} {'T(FLOP|B)/s':>11s}")
for (kind, detail), (work, ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{kind:14s} {detail:60s} {n // REPS:6d} {ms / n * 1e3:9.1f} {ms / REPS:8.3f} {ms / REPS / tot * 100:5.1f}% {work / (ms * 1e-3) / 1e12:11.1f}")
print(f"sum of timed launches: {tot:.2f} ms/step")
