"""Profiling driver for ncu: builds + calibrates the synthetic Flux-dev model, then runs ONE eager-launch
denoise step (1024x1024, batch 1) between cudaProfilerStart/Stop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flux_fp8_api_b200 import model as M, pipeline as PL  # noqa: E402

dev = torch.device("cuda", 0)
spec = M.flux_dev_spec()
net = PL.build_synthetic_flux(spec, dev)
req = PL.synthetic_request(spec.params, 1024, 1024, 1, 512, dev, seed=0)
PL.calibrate(net, req, num_steps=13)
sess = PL.DenoiseSession(net, req, use_graph=False)
sched = PL.get_schedule(28, req["img"].shape[1])
for _ in range(2):
    sess.step_device(req["img"], sched[0], sched[1])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
sess.step_device(req["img"], sched[0], sched[1])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step")
