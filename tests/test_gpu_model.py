"""GPU (-m gpu): block- and model-level parity of the fused CUDA path.

* against the committed reference outputs (tests/golden/flux_tiny.pt, minted from the unmodified
  reference on CPU): tolerance 2^-4 max-abs on O(1) activations, the spread the reference itself shows
  between its CPU and GPU matmul orders through two e5m2 bottlenecked blocks
* fused path == eager path == CUDA-graph replay
* batch sharding: a sample's result does not depend on which other samples share its batch (bit-exact),
  the property the multi-GPU partitioning rests on
* the drop-in surface: F8Linear.from_linear + 13-call calibration reproduces the reference's frozen scales."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
DEV = "cuda"


@pytest.fixture(scope="module")
def gold(golden_dir, lib):
    return torch.load(os.path.join(golden_dir, "flux_tiny.pt"))


@pytest.fixture()
def cpu_semantics():
    import flux_fp8_api_b200.f8linear as f8

    f8.SCALE_SEMANTICS = "cpu"
    yield
    f8.SCALE_SEMANTICS = "cuda"


def tiny_net(gold):
    from flux_fp8_api_b200 import model as M

    spec = M.FluxSpec(params=M.FluxParams(**gold["tiny"]), prequantized_flow=True)
    with torch.device(DEV):
        net = M.Flux(spec, dtype=BF16).to(BF16)
    net.load_state_dict(gold["state"], strict=True)
    return net.to(DEV).eval()


def maxdiff(a, b):
    return (a.float().cpu() - b.float().cpu()).abs().max().item()


def test_blocks_and_forward_match_reference_outputs(gold, cpu_semantics, monkeypatch):
    from flux_fp8_api_b200 import blocks

    net = tiny_net(gold)
    bi = {k: v.to(DEV) for k, v in gold["block_in"].items()}
    inp = {k: v.to(DEV) for k, v in gold["inputs"].items()}
    results = {}
    with torch.inference_mode():
        for mode in ("fused", "eager"):
            if mode == "eager":
                monkeypatch.setattr(blocks.DoubleStreamBlock, "_fusable", lambda self, a, b: False)
                monkeypatch.setattr(blocks.SingleStreamBlock, "_fusable", lambda self, a: False)
            d_img, d_txt = net.double_blocks[0](img=bi["img"], txt=bi["txt"], vec=bi["vec"], pe=bi["pe"])
            s_out = net.single_blocks[0](torch.cat((bi["txt"], bi["img"]), 1), vec=bi["vec"], pe=bi["pe"])
            y = net(**inp)
            assert maxdiff(d_img, gold["double_img"]) <= 2.0 ** -4
            assert maxdiff(d_txt, gold["double_txt"]) <= 2.0 ** -4
            assert maxdiff(s_out, gold["single"]) <= 2.0 ** -4
            assert maxdiff(y, gold["y_fp8"]) <= 2.0 ** -4
            assert torch.equal(net(**inp), y), "non-deterministic forward"
            results[mode] = y
    assert maxdiff(results["fused"], results["eager"]) <= 2.0 ** -5


def test_e4m3_activations_and_quantised_embedders_match_reference_outputs(golden_dir, lib, cpu_semantics, monkeypatch):
    """flux_tiny_e4m3.pt (reference quantised with input_float8_dtype=float8_e4m3fn, quantize_flow_embedder_layers=True):
    the fused path, the eager kernel path and the whole forward in the e4m3 x e4m3 arithmetic north_star names."""
    from flux_fp8_api_b200 import blocks, model as M, pipeline as PL
    from flux_fp8_api_b200.f8linear import F8Linear

    g = torch.load(os.path.join(golden_dir, "flux_tiny_e4m3.pt"))
    spec = M.FluxSpec(params=M.FluxParams(**g["tiny"]), prequantized_flow=True, quantize_flow_embedder_layers=True)
    with torch.device(DEV):
        net = M.Flux(spec, dtype=BF16).to(BF16)
    net.load_state_dict(g["state"], strict=True)
    net = net.to(DEV).eval()
    PL.set_input_float8_dtype(net, torch.float8_e4m3fn)
    assert sum(isinstance(m, F8Linear) for m in net.modules()) == g["n_f8"] == 21
    assert isinstance(net.img_in, F8Linear) and isinstance(net.vector_in.out_layer, F8Linear) and PL.all_frozen(net)
    bi = {k: v.to(DEV) for k, v in g["block_in"].items()}
    inp = {k: v.to(DEV) for k, v in g["inputs"].items()}
    results = {}
    with torch.inference_mode():
        for mode in ("fused", "eager"):
            if mode == "eager":
                monkeypatch.setattr(blocks.DoubleStreamBlock, "_fusable", lambda self, a, b: False)
                monkeypatch.setattr(blocks.SingleStreamBlock, "_fusable", lambda self, a: False)
            d_img, d_txt = net.double_blocks[0](img=bi["img"], txt=bi["txt"], vec=bi["vec"], pe=bi["pe"])
            s_out = net.single_blocks[0](torch.cat((bi["txt"], bi["img"]), 1), vec=bi["vec"], pe=bi["pe"])
            y = net(**inp)
            assert maxdiff(d_img, g["double_img"]) <= 2.0 ** -4
            assert maxdiff(d_txt, g["double_txt"]) <= 2.0 ** -4
            assert maxdiff(s_out, g["single"]) <= 2.0 ** -4
            assert maxdiff(y, g["y_fp8"]) <= 2.0 ** -4
            results[mode] = y
    assert maxdiff(results["fused"], results["eager"]) <= 2.0 ** -5
    # the e5m2 interpretation of the same state is a different function: the format switch is live
    PL.set_input_float8_dtype(net, torch.float8_e5m2)
    with torch.inference_mode():
        assert not torch.equal(net(**inp), results["eager"])


def test_four_step_trajectory_matches_reference(gold, golden_dir, cpu_semantics):
    """flux_tiny_traj.pt: 4 Euler steps through the reference (flux_pipeline.py:627-651).  Teacher-forced per step
    (each step starts from the reference's latent) to 2^-4, and free-running end to end -- eager launches, CUDA-graph
    replay and the host-buffer API -- within 4 x 2^-4 of the reference's final latent."""
    from flux_fp8_api_b200 import pipeline as PL

    traj = torch.load(os.path.join(golden_dir, "flux_tiny_traj.pt"))
    net = tiny_net(gold)
    req = {k: v.to(DEV) for k, v in traj["inputs"].items() if k != "timesteps"}
    ts = traj["timesteps"]
    eager = PL.DenoiseSession(net, req, use_graph=False)
    graph = PL.DenoiseSession(net, req, use_graph=True)
    prev = req["img"]
    for i, (t_curr, t_prev) in enumerate(zip(ts[:-1], ts[1:])):
        out_e = eager.step_device(prev, t_curr, t_prev)
        out_g = graph.step_device(prev, t_curr, t_prev)
        assert torch.equal(out_e, out_g), i
        assert maxdiff(out_e, traj["latents"][i]) <= 2.0 ** -4, i
        prev = traj["latents"][i].to(DEV)
    final = traj["latents"][-1]
    run_e, run_g = eager.run(ts), graph.run(ts)
    assert torch.equal(run_e, run_g)
    assert maxdiff(run_g, final) <= 4 * 2.0 ** -4
    host = req["img"].cpu().pin_memory()
    for t_curr, t_prev in zip(ts[:-1], ts[1:]):
        host = graph.step_host(host, t_curr, t_prev).clone()
    assert torch.equal(host.to(DEV), run_g)


def test_graph_survives_cache_resets_other_sessions_and_allocator_reuse(gold):
    """a captured graph reads per-request tensors (txt_in(txt), vector / guidance embeddings, pe,
    cos/sin).  They are owned by the GraphedStep; resetting the model's request cache, running other requests through
    the same model and churning the allocator in between must not change a replay."""
    from flux_fp8_api_b200 import pipeline as PL

    net = tiny_net(gold)
    req = {k: v.to(DEV) for k, v in gold["inputs"].items() if k != "timesteps"}
    sched = PL.get_schedule(4, req["img"].shape[1])
    sess = PL.DenoiseSession(net, req, use_graph=True)
    want = sess.step_device(req["img"], sched[0], sched[1])
    for rnd in range(3):
        net.reset_request_cache()
        other = {k: (v * 0.5 if v.is_floating_point() and k in ("txt", "y") else v).clone() for k, v in req.items()}
        PL.DenoiseSession(net, other, use_graph=(rnd % 2 == 0)).step_device(other["img"], sched[1], sched[2])
        net.reset_request_cache()
        torch.cuda.empty_cache()
        junk = [torch.full((1 << 18,), float(rnd + 1), device=DEV, dtype=BF16) for _ in range(64)]  # recycle freed blocks
        del junk
        assert torch.equal(sess.step_device(req["img"], sched[0], sched[1]), want), rnd
    assert sess.step.captures == 1
    # a LoRA under the cached embeddings bumps the epoch: the next call re-captures and sees the new weights
    lora = {"txt_in.lora_A.weight": torch.randn(4, net.txt_in.in_features).to(BF16),
            "txt_in.lora_B.weight": (torch.randn(net.txt_in.out_features, 4) * 0.5).to(BF16)}
    net.load_lora(lora, scale=1.0, name="t")
    changed = sess.step_device(req["img"], sched[0], sched[1])
    assert sess.step.captures == 2 and not torch.equal(changed, want)
    fresh = PL.DenoiseSession(net, req, use_graph=False).step_device(req["img"], sched[0], sched[1])
    assert torch.equal(changed, fresh)
    net.unload_lora("t")


def test_batch_beyond_the_batched_modulation_limit_falls_back(gold):
    """B = 17 exceeds fluxb200_modulation_batched's 16 rows; Flux.forward must fall back to the
    per-block Modulation path (M > 16 goes through the GEMM) instead of raising, with identical per-sample results."""
    net = tiny_net(gold)
    inp = {k: v.to(DEV) for k, v in gold["inputs"].items()}
    big = {k: v.repeat(9, *([1] * (v.dim() - 1)))[:17] for k, v in inp.items()}
    with torch.inference_mode():
        y2 = net(**inp)
        y17 = net(**big)
    assert y17.shape[0] == 17
    assert maxdiff(y17[:2], y2) <= 2.0 ** -5  # batched GEMV vs GEMM accumulation order
    assert torch.equal(y17[0], y17[2]) and torch.equal(y17[1], y17[3])


def test_batched_modulation_matches_per_layer(gold):
    """ModulationBank (one launch pair for every Modulation.lin) == each Modulation.forward on its own, to 1 bf16
    ulp (the two GEMV kernels accumulate in a different order), for several batch sizes."""
    from flux_fp8_api_b200 import blocks

    net = tiny_net(gold)
    mods = [net.double_blocks[0].img_mod, net.double_blocks[0].txt_mod, net.single_blocks[0].modulation]
    bank = blocks.ModulationBank(mods)
    assert not bank.stale()
    with torch.inference_mode():
        for B in (1, 2, 5):
            vec = torch.randn(B, net.hidden_size, device=DEV, generator=torch.Generator(device=DEV).manual_seed(B)).to(BF16)
            for m, (o1, o2) in zip(mods, bank(vec)):
                r1, r2 = m(vec)
                for ours, ref in zip(tuple(o1) + (tuple(o2) if o2 else ()), tuple(r1) + (tuple(r2) if r2 else ())):
                    assert ours.shape == ref.shape
                    assert maxdiff(ours, ref) <= 2.0 ** -7 * max(1.0, ref.abs().max().item())
        # and the model gives the same prediction with and without the bank
        inp = {k: v.to(DEV) for k, v in gold["inputs"].items()}
        y_bank = net(**inp)
        net.batch_modulation = False
        y_layer = net(**inp)
        assert maxdiff(y_bank, y_layer) <= 2.0 ** -5


def test_graph_replay_equals_eager_launch(gold):
    from flux_fp8_api_b200 import pipeline as PL

    net = tiny_net(gold)
    req = {k: v.to(DEV) for k, v in gold["inputs"].items() if k != "timesteps"}
    sched = PL.get_schedule(4, req["img"].shape[1])
    eager = PL.DenoiseSession(net, req, use_graph=False).run(sched)
    graph = PL.DenoiseSession(net, req, use_graph=True).run(sched)
    assert torch.isfinite(graph.float()).all()
    assert torch.equal(eager, graph)
    sess = PL.DenoiseSession(net, req, use_graph=True)
    host = req["img"].cpu().pin_memory()
    out_h = sess.step_host(host, sched[0], sched[1])
    assert torch.equal(out_h.to(DEV), sess.step_device(req["img"], sched[0], sched[1]))


def test_batch_sharding_is_bit_exact(gold):
    """Rows [lo,hi) of a batched run == the same samples run alone (what rank r computes under dp)."""
    from flux_fp8_api_b200 import parallel as PAR

    net = tiny_net(gold)
    inp = {k: v.to(DEV) for k, v in gold["inputs"].items()}
    with torch.inference_mode():
        full = net(**inp)
        for rank in range(2):
            shard = PAR.shard_request(inp, rank, 2)
            lo, hi = PAR.shard_range(2, rank, 2)
            assert torch.equal(net(**shard), full[lo:hi])


def test_from_linear_calibration_matches_reference_state(golden_dir, cpu_semantics):
    from flux_fp8_api_b200.f8linear import F8Linear

    cases = torch.load(os.path.join(golden_dir, "f8linear.pt"))
    assert {c["in_dtype"] for c in cases} == {"torch.float8_e5m2", "torch.float8_e4m3fn"}
    for c in cases:
        in_dt = torch.float8_e5m2 if "e5m2" in c["in_dtype"] else torch.float8_e4m3fn
        K, N = c["K"], c["N"]
        lin = torch.nn.Linear(K, N, bias=True).to(BF16)
        with torch.no_grad():
            lin.weight.copy_(c["weight_bf16"])
            lin.bias.copy_(c["state"]["bias"])
        f8 = F8Linear.from_linear(lin.to(DEV), input_float8_dtype=in_dt)
        assert f8.input_float8_dtype == in_dt
        sd = c["state"]
        assert torch.equal(f8.float8_data.view(torch.uint8).cpu(), sd["float8_data"].view(torch.uint8))
        assert torch.equal(f8.scale.cpu(), sd["scale"])
        # replay the 14 calls: amaxes were recorded when the golden was minted
        g = torch.Generator().manual_seed(0)
        for i, amax in enumerate(c["x_all_amax"]):
            x = torch.randn(2, c["M"], K, generator=g).to(BF16)
            x = (x * (amax / x.abs().max().item())).to(BF16)
            x.view(-1)[0] = amax  # pin the maximum exactly (amax is a bf16 value)
            f8(x.to(DEV))
        assert f8.frozen and f8.trial_index == 12
        assert torch.equal(f8.input_scale.cpu(), sd["input_scale"])
        y = f8(c["x_last"].to(DEV))
        assert maxdiff(y, c["y_last"]) <= 2.0 ** -7 * max(1.0, c["y_last"].abs().max().item())


def test_quantize_flow_swaps_the_same_layers_as_the_reference(gold):
    """quantize_flow_transformer_and_dispatch_float8 on a bf16 model yields the reference's layer split:
    13 F8Linear for the tiny model (blocks incl. modulation), embedders / final layer untouched."""
    from flux_fp8_api_b200 import model as M, pipeline as PL
    from flux_fp8_api_b200.f8linear import F8Linear

    spec = M.FluxSpec(params=M.FluxParams(**gold["tiny"]))
    net = PL.build_synthetic_flux(spec, DEV, seed=3)
    n_f8 = sum(isinstance(m, F8Linear) for m in net.modules())
    assert n_f8 == 13
    assert isinstance(net.img_in, torch.nn.Linear) and not isinstance(net.img_in, F8Linear)
    assert not PL.all_frozen(net)
    req = PL.synthetic_request(spec.params, 128, 128, 2, 32, DEV, seed=1)
    PL.calibrate(net, req, num_steps=13, shift=True)
    assert PL.all_frozen(net)
    ref_keys = set(gold["state"].keys())
    assert set(net.state_dict().keys()) == ref_keys
    spec2 = M.FluxSpec(params=spec.params, quantize_modulation=False)
    net2 = PL.build_synthetic_flux(spec2, DEV, seed=3)
    assert sum(isinstance(m, F8Linear) for m in net2.modules()) == 10


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configurations at full width (hidden 3072, 24 heads), one double + one single block deep:
# the fused CUDA path against the oracle evaluated on the same GPU tensors (torch's CUDA semantics, the
# reference's real target).  Sequence lengths: c2 1024^2 dev (S=4608), c3 1024^2 schnell (S=4352, no guidance),
# c4 768^2 (S=2816), c5 1536^2 with quantize_modulation=False (S=9728, bf16 Modulation.lin).
#
# Tolerance.  An fp8 pipeline is chaotic at the ulp level: a 1-ulp bf16 difference anywhere upstream flips
# the e5m2 rounding (25 % relative step) of a few percent of the next layer's activations, every flipped
# activation perturbs a whole output row of the next GEMM, and the perturbation compounds through the
# double block's two sequential quantise->GEMM stages and again through the single block (measured:
# profiles/r1_parity_noise_floor.md).  No two independent implementations -- including the reference run with
# a different SDPA backend -- agree to 2^-4 max-abs at this width.  The bar is therefore the REFERENCE'S OWN
# NOISE FLOOR: the oracle is evaluated a second time with the other legitimate rounding of the attention
# probabilities (bf16 P as flash/cuDNN SDPA do, against fp32 P of the math backend; oracle.SDPA_P_DTYPE), and
# our error against the oracle must not exceed that oracle-vs-oracle spread by more than 25 % (mean) /
# 50 % (p99.99, max), at every width-preserving stage the golden fixtures cover at 2^-4 absolute.
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,res,text_len,guidance,qmod,batch", [
    ("c2-dev-1024", 1024, 512, True, True, 1),
    ("c3-schnell-1024", 1024, 256, False, True, 1),
    ("c4-dev-768", 768, 512, True, True, 2),
    ("c5-dev-1536-bf16mod", 1536, 512, True, False, 1),
])
def test_baseline_configs_full_width(lib, name, res, text_len, guidance, qmod, batch):
    from flux_fp8_api_b200 import model as M, pipeline as PL
    from oracle import flux_oracle as O

    params = M.FluxParams(depth=1, depth_single_blocks=1, guidance_embed=guidance)
    spec = M.FluxSpec(params=params, quantize_modulation=qmod)
    net = PL.build_synthetic_flux(spec, DEV, seed=7)
    req = PL.synthetic_request(params, res, res, batch, text_len, DEV, seed=3)
    if not guidance:
        req["guidance"] = None
    PL.calibrate(net, req, num_steps=13, shift=guidance)
    assert PL.all_frozen(net)
    sd = {k: v for k, v in net.state_dict().items() if v is not None}
    cfg = dict(num_heads=24, depth=1, depth_single_blocks=1, axes_dim=[16, 56, 56], theta=10_000,
               guidance_embed=guidance)
    t = torch.full((batch,), 0.7, dtype=BF16, device=DEV)
    with torch.inference_mode():
        ours = net(img=req["img"], img_ids=req["img_ids"], txt=req["txt"], txt_ids=req["txt_ids"], timesteps=t,
                   y=req["y"], guidance=req["guidance"])
        ref = O.flux_forward(sd, cfg, req["img"], req["img_ids"], req["txt"], req["txt_ids"], t, req["y"], req["guidance"])
        O.SDPA_P_DTYPE = "bf16"
        try:
            ref_b = O.flux_forward(sd, cfg, req["img"], req["img_ids"], req["txt"], req["txt_ids"], t, req["y"],
                                   req["guidance"])
        finally:
            O.SDPA_P_DTYPE = "fp32"
        assert torch.isfinite(ours.float()).all()

        def spread(a, b):
            e = (a.float() - b.float()).abs().flatten()
            return e.mean().item(), torch.quantile(e[:: max(1, e.numel() // 1_000_000)], 0.9999).item(), e.max().item()

        floor = spread(ref_b, ref)
        # the closer of the two oracle roundings (our kernel rounds P to bf16 like the flash backends)
        err = min(spread(ours, ref), spread(ours, ref_b), key=lambda s: s[0])
        worst = max(spread(ours, ref), spread(ours, ref_b), key=lambda s: s[0])
        print(f"{name}: ours-oracle mean/p99.99/max = {err[0]:.3e}/{err[1]:.3e}/{err[2]:.3e}; "
              f"oracle noise floor = {floor[0]:.3e}/{floor[1]:.3e}/{floor[2]:.3e}; ref amax {ref.abs().max().item():.3f}")
        assert worst[0] <= 1.25 * floor[0] + 2.0 ** -9
        assert worst[1] <= 1.5 * floor[1] + 2.0 ** -7
        assert worst[2] <= 1.5 * floor[2] + 2.0 ** -6
        # no systematic offset: the mean signed error is far below the mean absolute error
        bias = (ours.float() - ref.float()).mean().abs().item()
        assert bias <= 0.05 * err[0] + 1e-4, bias
        # and the CUDA-graph session reproduces the eager-launch step bit for bit at this shape
        sched = PL.get_schedule(4, req["img"].shape[1], shift=guidance)
        sess_g = PL.DenoiseSession(net, req, use_graph=True)
        sess_e = PL.DenoiseSession(net, req, use_graph=False)
        assert torch.equal(sess_g.step_device(req["img"], sched[0], sched[1]), sess_e.step_device(req["img"], sched[0], sched[1]))


@pytest.mark.parametrize("res,batch", [(1024, 1), (512, 3)])
def test_layernorm_fused_into_the_gemm_launch_is_bit_identical(lib, res, batch):
    """fluxb200_f8_gemm_ln (LayerNorm-modulate-quantise as a prologue phase of the persistent GEMM grid + a grid barrier)
    against the two separate launches: same fp8 operands, same outputs, three launches fewer per double + single block;
    repeated calls and CUDA-graph replays re-use the self-re-arming barrier words."""
    from flux_fp8_api_b200 import _cabi, model as M, ops, pipeline as PL

    params = M.FluxParams(depth=1, depth_single_blocks=1)
    net = PL.build_synthetic_flux(M.FluxSpec(params=params), DEV, seed=11)
    req = PL.synthetic_request(params, res, res, batch, 512, DEV, seed=5)
    PL.calibrate(net, req, num_steps=13)
    t = torch.full((batch,), 0.6, dtype=BF16, device=DEV)
    call = dict(img=req["img"], img_ids=req["img_ids"], txt=req["txt"], txt_ids=req["txt_ids"], timesteps=t, y=req["y"],
                guidance=req["guidance"])
    prev = ops.FUSE_LN_INTO_GEMM
    try:
        with torch.inference_mode():
            ops.FUSE_LN_INTO_GEMM = False
            n0 = _cabi.LAUNCHES
            y_sep = net(**call)
            n_sep = _cabi.LAUNCHES - n0
            ops.FUSE_LN_INTO_GEMM = True
            n0 = _cabi.LAUNCHES
            y_fused = net(**call)
            n_fused = _cabi.LAUNCHES - n0
            assert n_fused == n_sep - 3, (n_sep, n_fused)
            assert torch.equal(y_fused, y_sep)
            for _ in range(5):
                assert torch.equal(net(**call), y_sep)
        sched = PL.get_schedule(4, req["img"].shape[1])
        g = PL.DenoiseSession(net, req, use_graph=True)    # captured with the fused launches
        ops.FUSE_LN_INTO_GEMM = False
        e = PL.DenoiseSession(net, req, use_graph=False)   # eager, separate launches
        for i in range(3):
            assert torch.equal(g.step_device(req["img"], sched[i], sched[i + 1]), e.step_device(req["img"], sched[i], sched[i + 1]))
    finally:
        ops.FUSE_LN_INTO_GEMM = prev


def test_prequantised_safetensors_roundtrip_and_the_reference_loads_it(gold, golden_dir, tmp_path, cpu_semantics):
    """pipeline.save_prequantized writes ONE safetensors file in the reference's key layout: load_prequantized brings
    back a frozen, graph-capturable model with bit-identical outputs, and -- when the staged reference is present -- the
    reference's own Flux(prequantized_flow=True) loads the same file the way util.load_flow_model does
    (load_sft + load_state_dict(assign=True), util.py:245-256) and runs it on its library kernels."""
    from safetensors.torch import load_file

    from flux_fp8_api_b200 import model as M, pipeline as PL
    from oracle import ref_loader as R

    net = tiny_net(gold)
    inp = {k: v.to(DEV) for k, v in gold["inputs"].items()}
    spec = M.FluxSpec(params=M.FluxParams(**gold["tiny"]), prequantized_flow=True)
    path = str(tmp_path / "tiny.f8.safetensors")
    header = PL.save_prequantized(net, path, spec)
    assert header["f8_layers"] == 13 and header["version"] == PL.PREQUANTIZED_VERSION
    back = PL.load_prequantized(path, DEV)
    assert PL.all_frozen(back)
    with torch.inference_mode():
        y = net(**inp)
        assert torch.equal(back(**inp), y)
        sched = PL.get_schedule(2, inp["img"].shape[1])
        req = {k: v for k, v in inp.items() if k != "timesteps"}
        assert torch.equal(PL.DenoiseSession(back, req).step_device(req["img"], sched[0], sched[1]),
                           PL.DenoiseSession(net, req).step_device(req["img"], sched[0], sched[1]))
    if not R.available():
        pytest.skip("oracle/_ref not staged: the reference-side load is checked where it is")
    ref = R.load()
    with torch.device("meta"):
        theirs = ref.fm.Flux(R.model_spec(ref, gold["tiny"], prequantized_flow=True), dtype=BF16)
    sd = load_file(path, device="cpu")
    missing, unexpected = theirs.load_state_dict(sd, strict=False, assign=True)
    assert not missing and not unexpected
    theirs = theirs.to(DEV).eval()
    assert R.all_frozen(ref, theirs)
    with torch.inference_mode():
        y_ref = theirs(**inp)
    assert maxdiff(y_ref, y) <= 2.0 ** -4
    assert maxdiff(y_ref, gold["y_fp8"]) <= 2.0 ** -4
