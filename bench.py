"""Headline benchmark: Flux-dev 1024x1024 denoise it/s (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" is one denoise step of the hot path (Flux.forward over all 19+38 blocks + the Euler update)
for one 1024x1024 sample per GPU (BASELINE.json configs[1]; batch-parallel across GPUs, weak scaling).
Weights are synthetic (seeded N(0, 0.02^2), quantised with the reference flow, input scales calibrated
by 13 eager steps on rank 0 and replicated with one NCCL broadcast); latents/text embeddings are
synthetic tensors of the published shapes.

One JSON line on stdout (rank 0): value = sample-steps/s over all GPUs with latents resident in HBM
(CUDA-graph replay per step); e2e = the same through DenoiseSession.step_host (pinned host latents,
H2D + step + D2H per step); roofline = tcgen05 FP8 GEMM kernel FLOP/s measured with CUDA events around
every launch of an instrumented pass; cpu_baseline = the oracle's bf16 blocks timed on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

HEIGHT = WIDTH = 1024
TEXT_LEN = 512
L_IMG = (HEIGHT // 16) * (WIDTH // 16)
S_TOTAL = L_IMG + TEXT_LEN
# algorithmic work per sample per step
F8_FLOPS = 1.29101e10 * S_TOTAL + 6.456e9
ATTN_FLOPS = 700416.0 * S_TOTAL * S_TOTAL


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.path = gpu_index, None, f"/tmp/fluxb200_clocks_{os.getpid()}.csv"

    def __enter__(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:  # noqa: BLE001
                self.proc.kill()
            self.f.close()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "power_w_max": None, "reasons": [], "samples": 0}
        try:
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            rows = [[c.strip() for c in r] for r in rows if len(r) >= 9]
            if not rows:
                return out
            sm = [float(r[1]) for r in rows]
            busy = [s for s, r in zip(sm, rows) if float(r[3]) > 300.0] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=float(rows[0][2]),
                       power_w_max=max(float(r[3]) for r in rows), samples=len(rows))
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for i, n in enumerate(names):
                if any(r[5 + i].lower().startswith("active") for r in rows):
                    out["reasons"].append(n)
        except Exception as ex:  # noqa: BLE001
            out["error"] = str(ex)
        return out


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's bf16 path (reference flow_dtype=bfloat16, no quantisation) on host cores
# ------------------------------------------------------------------------------------------------
def gemm_traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum per f8_gemm_kernel launch, averaged over the GEMM launches of one
    step, from the committed ncu capture (profiles/r1_gemm_traffic.json, written by tools/ncu_traffic.py from
    `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` over tests/profile_step.py).  None if absent."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_gemm_traffic.json")
    try:
        with open(path) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


def cpu_block_sample(threads: int, seed: int = 0):
    """Build one double and one single bf16 block of Flux-dev and return a closure that runs both at the
    1024x1024 sequence length (L=4096, T=512) through the oracle; a step = 19 double + 38 single blocks."""
    from oracle import flux_oracle as O

    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    D, H, MLP = 3072, 24, 12288
    bf = torch.bfloat16

    def lin(prefix, n, k, std=0.02):
        return {prefix + "weight": (torch.randn(n, k, generator=g) * std).to(bf),
                prefix + "bias": (torch.randn(n, generator=g) * 0.02).to(bf)}

    p = {}
    for s in ("img", "txt"):
        p.update(lin(f"d.{s}_mod.lin.", 6 * D, D, 0.01))
        p.update(lin(f"d.{s}_attn.qkv.", 3 * D, D))
        p.update(lin(f"d.{s}_attn.proj.", D, D))
        p.update(lin(f"d.{s}_mlp.0.", MLP, D))
        p.update(lin(f"d.{s}_mlp.2.", D, MLP))
        p[f"d.{s}_attn.norm.query_norm.scale"] = torch.ones(128, dtype=bf)
        p[f"d.{s}_attn.norm.key_norm.scale"] = torch.ones(128, dtype=bf)
    p.update(lin("s.modulation.lin.", 3 * D, D, 0.01))
    p.update(lin("s.linear1.", 3 * D + MLP, D))
    p.update(lin("s.linear2.", D, D + MLP))
    p["s.norm.query_norm.scale"] = torch.ones(128, dtype=bf)
    p["s.norm.key_norm.scale"] = torch.ones(128, dtype=bf)
    img = torch.randn(1, L_IMG, D, generator=g).to(bf)
    txt = torch.randn(1, TEXT_LEN, D, generator=g).to(bf)
    vec = torch.randn(1, D, generator=g).to(bf)
    ids = torch.cat((torch.zeros(1, TEXT_LEN, 3, dtype=bf), O.make_img_ids(1, HEIGHT // 16, WIDTH // 16, bf)), 1)
    pe = O.embed_nd(ids, [16, 56, 56], 10_000, bf)

    def run():
        with torch.inference_mode():
            t0 = time.perf_counter()
            i2, t2 = O.double_block(img, txt, vec, pe, p, "d.", H)
            t1 = time.perf_counter()
            O.single_block(torch.cat((t2, i2), 1), vec, pe, p, "s.", H)
            t2_ = time.perf_counter()
        return t1 - t0, t2_ - t1

    return run


def cpu_baseline_value(run, reps: int):
    td, ts = [], []
    for _ in range(reps):
        a, b = run()
        td.append(a)
        ts.append(b)
    step_s = 19 * statistics.median(td) + 38 * statistics.median(ts)
    return 1.0 / step_s, statistics.median(td), statistics.median(ts)


def run_reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation of the path = the oracle port of its bf16
    flow, all host threads, on the same workload; each step is a bounded sample (1 double + 1 single block at
    the full 1024x1024 sequence length) extrapolated to the 19 + 38 blocks of a step."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    run = cpu_block_sample(threads)
    for _ in range(max(0, min(args.warmup, 1))):
        run()
    steps = max(1, min(args.steps, 6))
    t0 = time.perf_counter()
    value, td, ts = cpu_baseline_value(run, steps)
    wall = time.perf_counter() - t0
    sample = (f"{steps} x (1 DoubleStreamBlock + 1 SingleStreamBlock, bf16, B=1, L=4096, T=512) via oracle port; "
              f"step = 19*{td:.3f}s + 38*{ts:.3f}s")
    line = {
        "impl": "reference", "metric": "Flux-dev 1024x1024 denoise it/s", "value": value, "unit": "it/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "Flux-dev 1024x1024 denoise step, batch 1, bf16 reference flow on host CPU",
                   "resolution": [HEIGHT, WIDTH], "seq_len": S_TOTAL},
        "cpu_baseline": {"value": value, "unit": "it/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=28)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    from flux_fp8_api_b200 import _cabi, model as M, ops, parallel as PAR, pipeline as PL

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback "
                         "(use --impl reference for the CPU baseline arm)")
    rank, local_rank, world = PAR.init_distributed("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ops.device_check()
    peaks, peaks_src = measured_peaks()

    # ---- model: synthetic weights -> reference quantisation flow -> calibrate on rank 0 -> one broadcast
    spec = M.flux_dev_spec()
    net = PL.build_synthetic_flux(spec, dev, seed=1234)
    req = PL.synthetic_request(spec.params, HEIGHT, WIDTH, 1, TEXT_LEN, dev, seed=0, sample_offset=rank)
    bcast_bytes, bcast_ms = 0, 0.0
    if rank == 0:
        PL.calibrate(net, req, num_steps=13)
    if world > 1:
        if rank != 0:  # allocate scale buffers so shapes agree, then receive rank 0's state
            for m in net.modules():
                if isinstance(m, M.F8Linear):
                    z = torch.zeros((), dtype=torch.float32, device=dev)
                    m.input_scale, m.input_scale_reciprocal = z.clone(), z.clone()
        torch.cuda.synchronize()
        PAR.barrier()
        t0 = time.perf_counter()
        bcast_bytes = PAR.broadcast_state(net, src=0)
        torch.cuda.synchronize()
        bcast_ms = (time.perf_counter() - t0) * 1e3
        PAR.frozen_flags_sync(net)
    assert PL.all_frozen(net)

    sess = PL.DenoiseSession(net, req, use_graph=not args.no_graph)
    sched = PL.get_schedule(max(args.steps, 1), L_IMG, shift=True)

    # launches per step (counted while running one eager-launch step)
    before = _cabi.LAUNCHES
    with torch.inference_mode():
        tv = torch.full((1,), sched[0], dtype=torch.bfloat16, device=dev)
        net(img=req["img"], img_ids=req["img_ids"], txt=req["txt"], txt_ids=req["txt_ids"], y=req["y"], timesteps=tv,
            guidance=req["guidance"])
    launches_per_step = _cabi.LAUNCHES - before

    def timed_loop(step_fn, steps):
        torch.cuda.synchronize()
        PAR.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        step_fn(steps)
        e.record()
        torch.cuda.synchronize()
        PAR.barrier()
        return PAR.max_over_ranks(s.elapsed_time(e), dev)

    # ---- value: latents resident in HBM
    def device_steps(k):
        img = req["img"]
        for i in range(k):
            img = sess.step_device(img, sched[i % (len(sched) - 1)], sched[i % (len(sched) - 1) + 1])
        return img

    device_steps(args.warmup)
    with ClockSampler(local_rank) as clocks:
        ms_total = timed_loop(device_steps, args.steps)
    clock_info = clocks.summary()
    ms_per_step = ms_total / args.steps
    value = world * 1000.0 / ms_per_step

    # ---- e2e: host buffers in, host buffers out, every step
    host_img = torch.empty(req["img"].shape, dtype=req["img"].dtype, pin_memory=True)
    host_img.copy_(req["img"])

    def host_steps(k):
        h = host_img
        for i in range(k):
            h = sess.step_host(h, sched[i % (len(sched) - 1)], sched[i % (len(sched) - 1) + 1])
        return h

    host_steps(2)
    e2e_ms = timed_loop(host_steps, args.steps) / args.steps
    e2e_value = world * 1000.0 / e2e_ms

    # ---- roofline: CUDA events around every tcgen05 FP8 GEMM launch of an eager-launch pass
    roof = None
    if rank == 0:
        ops.KERNEL_TIMELINE = []
        eager = PL.DenoiseSession(net, req, use_graph=False)
        with torch.inference_mode():
            for i in range(3):
                if i == 1:
                    ops.KERNEL_TIMELINE.clear()
                eager.step_device(req["img"], sched[0], sched[1])
        torch.cuda.synchronize()
        tl, ops.KERNEL_TIMELINE = ops.KERNEL_TIMELINE, None
        agg = {}
        for kind, flops, s, e, *_ in tl:
            a = agg.setdefault(kind, [0.0, 0.0, 0])
            a[0] += flops
            a[1] += s.elapsed_time(e)
            a[2] += 1
        gf, gms, gn = agg["f8_gemm"]
        af, ams, an = agg["attention"]
        hbm = {}
        for kind in ("ln_mod_quant", "modulation_batched"):
            if kind in agg:
                bts, ms_, cnt = agg[kind]
                hbm[kind] = {"achieved": bts / (ms_ * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": bts / (ms_ * 1e-3) / 1e9 / peaks["hbm_gbs"], "launches_per_step": cnt // 2,
                             "ms_per_step": ms_ / 2}
        fp8_peak = 2.0 * peaks["bf16_tflops_sustained"]
        # The event-timed pass launches eagerly: the host-side gaps between its launches let the power-capped GPU
        # clock higher than it does inside the back-to-back CUDA-graph step.  `achieved` therefore charges the kernel
        # its SHARE of the event-timed launches applied to the timed graph step (the in-step rate); the raw
        # event-timed rate is reported beside it.
        timed_ms = sum(a[1] for a in agg.values()) / 2
        gemm_share = (gms / 2) / timed_ms
        attn_share = (ams / 2) / timed_ms
        gemm_ms_in_step = gemm_share * ms_per_step
        attn_ms_in_step = attn_share * ms_per_step
        achieved = (gf / 2) / (gemm_ms_in_step * 1e-3) / 1e12
        attn_achieved = (af / 2) / (attn_ms_in_step * 1e-3) / 1e12
        roof = {
            "bound": "tensor", "kernel": "f8_gemm_kernel (tcgen05 kind::f8f6f4)", "achieved": achieved,
            "peak": fp8_peak, "unit": "TFLOP/s", "frac": achieved / fp8_peak, "traffic": gemm_traffic_per_launch(),
            "peak_source": f"2 x bf16_tflops_sustained, {peaks_src}; fp8 tensor rate is twice bf16",
            "method": "algorithmic flops of all GEMM launches of a step / (share of the CUDA-event-timed launch time "
                      "x graph-replay ms per step)",
            "achieved_event_timed_eager": gf / (gms * 1e-3) / 1e12,
            "launches_per_step": gn // 2, "gemm_ms_per_step": gemm_ms_in_step, "gemm_share_of_step": gemm_share,
            "gemm_ms_per_step_event_timed_eager": gms / 2,
            "attention": {"achieved": attn_achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                          "frac": attn_achieved / peaks["bf16_tflops_sustained"],
                          "achieved_event_timed_eager": af / (ams * 1e-3) / 1e12,
                          "launches_per_step": an // 2, "ms_per_step": attn_ms_in_step, "share_of_step": attn_share},
            "hbm_kernels": hbm,
            "step_tensor_frac": (F8_FLOPS / (fp8_peak * 1e12) + ATTN_FLOPS / (peaks["bf16_tflops_sustained"] * 1e12))
                                / (ms_per_step * 1e-3),
            "fp8_pipe_util_nominal": F8_FLOPS / (ms_per_step * 1e-3) / 4.5e15,
        }

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        run = cpu_block_sample(threads)
        v, td, ts = cpu_baseline_value(run, 2)
        cpu = {"value": v, "unit": "it/s", "cores": threads, "kind": "port",
               "sample": f"2 x (1 DoubleStreamBlock + 1 SingleStreamBlock, bf16, L=4096, T=512) via oracle; "
                         f"step = 19*{td:.3f}s + 38*{ts:.3f}s"}

    if rank == 0:
        line = {
            "metric": "Flux-dev 1024x1024 denoise it/s", "value": value, "unit": "it/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp8 (e5m2 activations x e4m3 weights, fp32 accumulate; "
                                                             "bf16 attention)",
            "data": "synthetic",
            "config": {"workload": "Flux-dev 1024x1024 denoise step (19 double + 38 single blocks), batch 1 per GPU",
                       "resolution": [HEIGHT, WIDTH], "seq_len": S_TOTAL, "text_len": TEXT_LEN,
                       "global_batch": world, "parallelism": f"dp{world} (image batch, replicated fp8 weights)",
                       "l2": "inputs larger than L2: 11.8 GB of fp8 weights stream per step (L2 is 126 MB)",
                       "cuda_graph": not args.no_graph, "weights": "seeded synthetic, reference quantise+calibrate flow"},
            "e2e": {"value": e2e_value, "unit": "it/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": sess.h2d_bytes_per_step, "d2h_bytes_per_step": sess.d2h_bytes_per_step},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "clocks": clock_info,
            "roofline": roof,
            "cpu_baseline": cpu,
            "broadcast": {"bytes": bcast_bytes, "ms": bcast_ms},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
