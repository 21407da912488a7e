"""Mint the golden fixtures under tests/golden/ from the UNMODIFIED reference and pin the oracle.

Run in the authoring container only (needs /root/reference, CPU torch):

    python oracle/make_golden.py

The reference (aredden/flux-fp8-api) ships no tests or golden vectors for the hot path
(SURVEY.md section 4 / 8c), so the vectors are produced here by importing its modules
(float8_quantize.py, modules/flux_model.py) and running them on seeded synthetic inputs.  For every
case the oracle (oracle/flux_oracle.py) is run on the same inputs and must agree with the reference
before the fixture is written; tests/test_oracle_golden.py re-checks the oracle against the
committed reference outputs without needing /root/reference.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FLUX_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

# the reference logs through loguru; keep the import working if it is absent
try:
    import loguru  # noqa: F401
except ImportError:  # pragma: no cover
    stub = types.ModuleType("loguru")
    stub.logger = types.SimpleNamespace(info=print, warning=print, error=print, debug=print)
    sys.modules["loguru"] = stub

import float8_quantize as ref_f8  # noqa: E402  (reference)
from modules import flux_model as ref_fm  # noqa: E402  (reference)

from oracle import flux_oracle as O  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
BF16 = torch.bfloat16

TINY = dict(in_channels=64, vec_in_dim=64, context_in_dim=96, hidden_size=256, mlp_ratio=2.0, num_heads=2,
            depth=1, depth_single_blocks=1, axes_dim=[16, 56, 56], theta=10_000, qkv_bias=True,
            guidance_embed=True)


def maxdiff(a, b):
    return (a.float() - b.float()).abs().max().item()


def report(name, ref, ora, tol, exact=False):
    if exact:
        same = torch.equal(ref.view(torch.uint8) if ref.dtype.itemsize == 1 else ref,
                           ora.view(torch.uint8) if ora.dtype.itemsize == 1 else ora)
        print(f"  {name:34s} exact={same}")
        assert same, name
        return
    d = maxdiff(ref, ora)
    frac = ((ref.float() - ora.float()).abs() > 0).float().mean().item()
    print(f"  {name:34s} max|diff|={d:.3e}  mismatching={frac*100:.3f}%  (tol {tol:g})")
    assert d <= tol, f"oracle disagrees with the reference on {name}: {d} > {tol}"


def special_values():
    g = torch.Generator().manual_seed(7)
    base = torch.randn(4096, generator=g) * 3
    mids = []
    # values straddling e5m2 / e4m3 rounding mid-points after scaling, +-0, subnormals, +-max, huge
    for e in range(-20, 17):
        for m in (1.0, 1.0625, 1.125, 1.1875, 1.25, 1.375, 1.5, 1.625, 1.75, 1.875, 1.9375):
            mids += [m * 2.0 ** e, -m * 2.0 ** e]
    extra = torch.tensor(mids + [0.0, -0.0, 448.0, -448.0, 57344.0, -57344.0, 1e30, -1e30, 3.4e38, 1e-30, 6e-8])
    return torch.cat([base, extra, torch.randn(1000, generator=g) * 1e-3]).to(BF16)


def golden_quantize():
    print("[quantize]")
    x = special_values()
    lin = ref_f8.F8Linear(16, 16, bias=False, dtype=BF16)
    out = {"x": x, "cases": []}
    for dt in (torch.float8_e5m2, torch.float8_e4m3fn):
        mx = torch.finfo(dt).max
        for amax in (0.003, 0.7, 1.0, 5.0, 300.0, 1e-14):
            scale = lin.amax_to_scale(torch.tensor(amax, dtype=torch.float32), mx)
            yref = lin.to_fp8_saturated(x, scale, mx).to(dt)
            yora = O.quantize(x, scale, dt)
            report(f"{dt} amax={amax}", yref, yora, 0, exact=True)
            assert torch.equal(scale, O.amax_to_scale(torch.tensor(amax, dtype=torch.float32), mx))
            out["cases"].append({"dtype": str(dt), "amax": amax, "scale": scale, "y": yref.view(torch.uint8)})
    torch.save(out, os.path.join(OUT, "quantize.pt"))


def golden_f8linear():
    print("[F8Linear]")
    g = torch.Generator().manual_seed(11)
    cases = []
    for (M, N, K, in_dt) in [(128, 256, 128, torch.float8_e5m2), (200, 384, 320, torch.float8_e5m2),
                             (96, 128, 256, torch.float8_e4m3fn), (3, 512, 256, torch.float8_e5m2)]:
        lin = torch.nn.Linear(K, N, bias=True)
        with torch.no_grad():
            lin.weight.normal_(0, 0.02, generator=g)
            lin.bias.normal_(0, 0.02, generator=g)
        lin = lin.to(BF16)
        f8 = ref_f8.F8Linear.from_linear(lin, input_float8_dtype=in_dt)
        xs = [(torch.randn(2, M, K, generator=g) * (1.0 + 0.3 * i)).to(BF16) for i in range(14)]
        with torch.inference_mode():
            outs = [f8(x) for x in xs]  # 12 calibration calls, 1 freezing call, 1 frozen call
        sd = {k: v.clone() for k, v in f8.state_dict().items()}
        assert f8.input_scale_initialized
        # oracle: calibration trace
        cal = O.CalibratingLinear(12, in_dt)
        p = {"float8_data": sd["float8_data"], "scale_reciprocal": sd["scale_reciprocal"], "bias": sd["bias"]}
        for i, x in enumerate(xs):
            xq = cal.quantize_input(x)
            y = O.scaled_mm(xq.reshape(-1, K), p["float8_data"], cal.input_scale.reciprocal(), p["scale_reciprocal"],
                            p["bias"]).reshape(2, M, N)
            report(f"M{M} N{N} K{K} call {i}", outs[i], y, 2.0 ** -6 * max(1.0, outs[i].abs().max().item()))
        assert torch.equal(cal.input_scale, sd["input_scale"]), "calibration trace differs"
        wq, ws, wsr = O.quantize_weight(lin.weight.data)
        assert torch.equal(wq.view(torch.uint8), sd["float8_data"].view(torch.uint8)) and torch.equal(ws, sd["scale"])
        cases.append({"M": M, "N": N, "K": K, "in_dtype": str(in_dt), "weight_bf16": lin.weight.data.clone(),
                      "state": sd, "x_last": xs[-1], "y_last": outs[-1].clone(), "x_all_amax": [x.abs().max().item() for x in xs]})
    torch.save(cases, os.path.join(OUT, "f8linear.pt"))


def golden_ops():
    print("[ops]")
    g = torch.Generator().manual_seed(23)
    B, H, S, D = 2, 2, 160, 128
    ids = torch.zeros(B, S, 3)
    ids[:, 32:, 1] = torch.arange(S - 32).float()[None] // 16
    ids[:, 32:, 2] = torch.arange(S - 32).float()[None] % 16 + 80  # positions up to 95
    ids = ids.to(BF16)
    emb = ref_fm.EmbedND(dim=128, theta=10_000, axes_dim=[16, 56, 56], dtype=BF16)
    pe = emb(ids)
    report("EmbedND", pe, O.embed_nd(ids, [16, 56, 56], 10_000, BF16), 0, exact=True)
    q = torch.randn(B, H, S, D, generator=g).to(BF16)
    k = torch.randn(B, H, S, D, generator=g).to(BF16)
    v = torch.randn(B, H, S, D, generator=g).to(BF16)
    qr, kr = ref_fm.apply_rope(q, k, pe)
    qo, ko = O.apply_rope(q, k, pe)
    report("apply_rope q", qr, qo, 0, exact=True)
    report("apply_rope k", kr, ko, 0, exact=True)
    norm = ref_fm.QKNorm(D)
    with torch.no_grad():
        norm.query_norm.scale.copy_(1 + 0.05 * torch.randn(D, generator=g))
        norm.key_norm.scale.copy_(1 + 0.05 * torch.randn(D, generator=g))
    norm = norm.to(BF16)
    qn, kn = norm(q, k, v)
    report("QKNorm q", qn, O.rms_norm(q, norm.query_norm.scale.data), 2.0 ** -7)
    report("QKNorm k", kn, O.rms_norm(k, norm.key_norm.scale.data), 2.0 ** -7)
    att = ref_fm.attention(q, k, v, pe)
    report("attention", att, O.attention(q, k, v, pe), 2.0 ** -6)
    x = (torch.randn(B, S, 256, generator=g) * 2 + 0.3).to(BF16)
    shift = (torch.randn(B, 1, 256, generator=g) * 0.2).to(BF16)
    scale = (torch.randn(B, 1, 256, generator=g) * 0.2).to(BF16)
    ln = torch.nn.LayerNorm(256, elementwise_affine=False, eps=1e-6)
    lnm = (1 + scale) * ln(x) + shift
    report("LN-modulate", lnm, O.layernorm_modulate(x, shift, scale), 2.0 ** -5)
    t = torch.tensor([1.0, 0.73, 0.002], dtype=BF16)
    te = ref_fm.timestep_embedding(t, 256)
    report("timestep_embedding", te, O.timestep_embedding(t, 256), 0, exact=True)
    gel = torch.nn.GELU(approximate="tanh")(x)
    sil = torch.nn.SiLU()(x)
    torch.save({"ids": ids, "pe": pe, "q": q, "k": k, "v": v, "q_rope": qr, "k_rope": kr,
                "qnorm_w": norm.query_norm.scale.data.clone(), "knorm_w": norm.key_norm.scale.data.clone(),
                "q_norm": qn, "k_norm": kn, "attention": att, "x": x, "shift": shift, "scale": scale, "ln_mod": lnm,
                "t": t, "t_emb": te, "gelu": gel, "silu": sil}, os.path.join(OUT, "ops.pt"))


def init_reference_flux(seed: int):
    cfg = types.SimpleNamespace(params=ref_fm.FluxParams(**TINY), prequantized_flow=False, quantize_modulation=True,
                                quantize_flow_embedder_layers=False)
    g = torch.Generator().manual_seed(seed)
    model = ref_fm.Flux(cfg, dtype=BF16)
    with torch.no_grad():
        for name, mod in model.named_modules():
            if isinstance(mod, torch.nn.Linear):
                std = 0.01 if name.endswith("mod.lin") or name.endswith("modulation.lin") else 0.04
                mod.weight.normal_(0, std, generator=g)
                if mod.bias is not None:
                    mod.bias.normal_(0, 0.02, generator=g)
            if isinstance(mod, ref_fm.RMSNorm):
                mod.scale.copy_(1 + 0.05 * torch.randn(mod.scale.shape, generator=g))
    return model.type(BF16).eval()


def flux_inputs(seed: int, B: int = 2, hw=(8, 8), T: int = 32, t: float = 1.0):
    g = torch.Generator().manual_seed(seed)
    L = hw[0] * hw[1]
    return dict(
        img=torch.randn(B, L, TINY["in_channels"], generator=g).to(BF16),
        img_ids=O.make_img_ids(B, hw[0], hw[1], BF16),
        txt=(0.5 * torch.randn(B, T, TINY["context_in_dim"], generator=g)).to(BF16),
        txt_ids=torch.zeros(B, T, 3, dtype=BF16),
        timesteps=torch.full((B,), t, dtype=BF16),
        y=torch.randn(B, TINY["vec_in_dim"], generator=g).to(BF16),
        guidance=torch.full((B,), 3.5, dtype=BF16),
    )


def golden_flux():
    print("[Flux tiny: bf16 path]")
    model = init_reference_flux(1234)
    cfg = {k: TINY[k] for k in ("num_heads", "depth", "depth_single_blocks", "axes_dim", "theta", "guidance_embed")}
    inp = flux_inputs(99)
    with torch.inference_mode():
        y_bf16 = model(**inp)
    sd_bf16 = {k: v.clone() for k, v in model.state_dict().items()}
    report("Flux.forward bf16", y_bf16, O.flux_forward(sd_bf16, cfg, **inp), 2.0 ** -4)

    print("[Flux tiny: fp8 path, reference quantise + 13-call calibration]")
    ref_f8.quantize_flow_transformer_and_dispatch_float8(
        model, torch.device("cpu"), offload_flow=False, swap_linears_with_cublaslinear=False, flow_dtype=BF16,
        quantize_modulation=True, quantize_flow_embedder_layers=False)
    sched = O.get_schedule(13, 64)
    with torch.inference_mode():
        for i in range(13):
            model(**flux_inputs(100 + i, t=sched[i]))
    n_f8 = sum(isinstance(m, ref_f8.F8Linear) for m in model.modules())
    assert all(m.input_scale_initialized for m in model.modules() if isinstance(m, ref_f8.F8Linear))
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    print(f"  {n_f8} F8Linear layers, {sum(v.numel() * v.element_size() for v in sd.values()) / 1e6:.2f} MB state")
    with torch.inference_mode():
        y_f8 = model(**inp)
        # per-block references on fresh activations
        gg = torch.Generator().manual_seed(5)
        B, L, T, Dm = 2, 64, 32, TINY["hidden_size"]
        img = torch.randn(B, L, Dm, generator=gg).to(BF16)
        txt = torch.randn(B, T, Dm, generator=gg).to(BF16)
        vec = torch.randn(B, Dm, generator=gg).to(BF16)
        pe = model.pe_embedder(torch.cat((inp["txt_ids"], inp["img_ids"]), dim=1))
        d_img, d_txt = model.double_blocks[0](img=img, txt=txt, vec=vec, pe=pe)
        xs = torch.cat((txt, img), 1)
        s_out = model.single_blocks[0](xs, vec=vec, pe=pe)
        mod1, mod2 = model.double_blocks[0].img_mod(vec)
    o_img, o_txt = O.double_block(img, txt, vec, pe, sd, "double_blocks.0.", TINY["num_heads"])
    report("DoubleStreamBlock img", d_img, o_img, 2.0 ** -4)
    report("DoubleStreamBlock txt", d_txt, o_txt, 2.0 ** -4)
    report("SingleStreamBlock", s_out, O.single_block(xs, vec, pe, sd, "single_blocks.0.", TINY["num_heads"]), 2.0 ** -4)
    (osh, osc, oga), _ = O.modulation(vec, sd, "double_blocks.0.img_mod.", True)
    report("Modulation shift", mod1.shift, osh, 2.0 ** -8)
    report("Flux.forward fp8", y_f8, O.flux_forward(sd, cfg, **inp), 2.0 ** -4)
    print(f"  fp8-vs-bf16 reference max|diff| = {maxdiff(y_f8, y_bf16):.4f}  (output amax {y_bf16.abs().max():.3f})")
    torch.save({"tiny": TINY, "cfg": cfg, "state": sd, "inputs": inp, "y_fp8": y_f8, "y_bf16": y_bf16,
                "block_in": {"img": img, "txt": txt, "vec": vec, "pe": pe},
                "double_img": d_img, "double_txt": d_txt, "single": s_out,
                "mod1": [mod1.shift, mod1.scale, mod1.gate]}, os.path.join(OUT, "flux_tiny.pt"))


def golden_flux_variants():
    """Two more reference-minted fixtures (e4m3 activations untested beyond one GEMM; no multi-step golden):

    flux_tiny_e4m3.pt  the tiny Flux quantised with input_float8_dtype=float8_e4m3fn AND
                       quantize_flow_embedder_layers=True (img_in / txt_in / time_in / vector_in / guidance_in become
                       F8Linear too): state, per-block outputs, full forward.
    flux_tiny_traj.pt  a 4-step Euler trajectory of the e5m2 model of flux_tiny.pt through the reference Flux.forward,
                       with the loop of flux_pipeline.py:627-651 (t_vec in bf16, img + (t_prev - t_curr) * pred)."""
    E4M3 = torch.float8_e4m3fn
    cfg = {k: TINY[k] for k in ("num_heads", "depth", "depth_single_blocks", "axes_dim", "theta", "guidance_embed")}
    print("[Flux tiny: e4m3 activations + quantised embedders]")
    model = init_reference_flux(4321)
    ref_f8.quantize_flow_transformer_and_dispatch_float8(
        model, torch.device("cpu"), input_float8_dtype=E4M3, offload_flow=False, swap_linears_with_cublaslinear=False,
        flow_dtype=BF16, quantize_modulation=True, quantize_flow_embedder_layers=True)
    sched = O.get_schedule(13, 64)
    with torch.inference_mode():
        for i in range(13):
            model(**flux_inputs(300 + i, t=sched[i]))
    f8s = [m for m in model.modules() if isinstance(m, ref_f8.F8Linear)]
    assert all(m.input_scale_initialized and m.input_float8_dtype == E4M3 for m in f8s)
    assert isinstance(model.img_in, ref_f8.F8Linear) and isinstance(model.time_in.in_layer, ref_f8.F8Linear)
    assert not isinstance(model.final_layer.linear, ref_f8.F8Linear)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    inp = flux_inputs(199)
    with torch.inference_mode():
        y = model(**inp)
        gg = torch.Generator().manual_seed(6)
        B, L, T, Dm = 2, 64, 32, TINY["hidden_size"]
        img = torch.randn(B, L, Dm, generator=gg).to(BF16)
        txt = torch.randn(B, T, Dm, generator=gg).to(BF16)
        vec = torch.randn(B, Dm, generator=gg).to(BF16)
        pe = model.pe_embedder(torch.cat((inp["txt_ids"], inp["img_ids"]), dim=1))
        d_img, d_txt = model.double_blocks[0](img=img, txt=txt, vec=vec, pe=pe)
        xs = torch.cat((txt, img), 1)
        s_out = model.single_blocks[0](xs, vec=vec, pe=pe)
    o_img, o_txt = O.double_block(img, txt, vec, pe, sd, "double_blocks.0.", TINY["num_heads"], E4M3)
    report("DoubleStreamBlock img (e4m3)", d_img, o_img, 2.0 ** -4)
    report("DoubleStreamBlock txt (e4m3)", d_txt, o_txt, 2.0 ** -4)
    report("SingleStreamBlock (e4m3)", s_out, O.single_block(xs, vec, pe, sd, "single_blocks.0.", TINY["num_heads"], E4M3),
           2.0 ** -4)
    report("Flux.forward fp8 (e4m3, f8 embedders)", y, O.flux_forward(sd, cfg, **inp, in_dtype=E4M3), 2.0 ** -4)
    print(f"  {len(f8s)} F8Linear layers")
    torch.save({"tiny": TINY, "cfg": cfg, "state": sd, "inputs": inp, "y_fp8": y,
                "block_in": {"img": img, "txt": txt, "vec": vec, "pe": pe},
                "double_img": d_img, "double_txt": d_txt, "single": s_out, "n_f8": len(f8s),
                "input_float8_dtype": "float8_e4m3fn", "quantize_flow_embedder_layers": True},
               os.path.join(OUT, "flux_tiny_e4m3.pt"))

    print("[Flux tiny: 4-step Euler trajectory, e5m2 model of flux_tiny.pt]")
    gold = torch.load(os.path.join(OUT, "flux_tiny.pt"))
    spec = types.SimpleNamespace(params=ref_fm.FluxParams(**TINY), prequantized_flow=True, quantize_modulation=True,
                                 quantize_flow_embedder_layers=False)
    model = ref_fm.Flux(spec, dtype=BF16)
    model.load_state_dict(gold["state"], strict=True, assign=True)
    model.eval()
    inp = {k: v.clone() for k, v in gold["inputs"].items()}
    timesteps = O.get_schedule(4, inp["img"].shape[1])
    img = inp["img"]
    preds, latents = [], []
    o_img = inp["img"]
    with torch.inference_mode():
        t_vec = None
        for t_curr, t_prev in zip(timesteps[:-1], timesteps[1:]):   # flux_pipeline.py:634-651
            if t_vec is None:
                t_vec = torch.full((img.shape[0],), t_curr, dtype=img.dtype)
            else:
                t_vec = t_vec.reshape((img.shape[0],)).fill_(t_curr)
            pred = model(img=img, img_ids=inp["img_ids"], txt=inp["txt"], txt_ids=inp["txt_ids"], y=inp["y"],
                         timesteps=t_vec, guidance=inp["guidance"])
            img = img + (t_prev - t_curr) * pred
            preds.append(pred.clone())
            latents.append(img.clone())
            o_pred = O.flux_forward(gold["state"], cfg, o_img, inp["img_ids"], inp["txt"], inp["txt_ids"],
                                    t_vec.clone(), inp["y"], inp["guidance"])
            o_img = O.euler_step(o_img, o_pred, t_curr, t_prev)
    report("4-step trajectory, final latent", img, o_img, 2.0 ** -3)
    torch.save({"timesteps": timesteps, "preds": preds, "latents": latents, "inputs": inp},
               os.path.join(OUT, "flux_tiny_traj.pt"))


def golden_lora():
    """LoRA fuse / unfuse through the reference's lora_loading functions on reference F8Linear layers."""
    import lora_loading as ref_lora  # reference

    print("lora fuse (reference lora_loading.py)")
    g = torch.Generator().manual_seed(77)
    cases = []
    # (N, K, rank of lora_B, rows of lora_A, alpha, lora_scale)
    for name, N, K, r, ra, alpha, ls in [("even", 384, 256, 16, 16, None, 1.0),
                                          ("alpha", 256, 512, 8, 8, 4.0, 0.75),
                                          ("uneven", 768, 256, 4, 12, None, 1.25),
                                          ("big-delta", 256, 256, 32, 32, 16.0, 3.0)]:
        lin = torch.nn.Linear(K, N, bias=True).to(BF16)
        lin.weight.data = (torch.randn(N, K, generator=g) * 0.02).to(BF16)
        mod = ref_f8.F8Linear.from_linear(lin)
        mod.quantize_weight()
        before = dict(float8_data=mod.float8_data.clone(), scale=mod.scale.clone(),
                      scale_reciprocal=mod.scale_reciprocal.clone())
        lora_A = (torch.randn(ra, K, generator=g) * 0.05).to(BF16)
        lora_B = (torch.randn(N, r, generator=g) * 0.05).to(BF16)
        lora_sd = (lora_A, lora_B, alpha)
        # exactly the loop body of apply_lora_to_model (lora_loading.py:679-687)
        weight, is_f8, dtype = ref_lora.extract_weight_from_linear(mod)
        assert is_f8 and dtype == BF16
        weight = ref_lora.apply_lora_weight_to_module(weight, lora_sd, lora_scale=ls)
        mod.set_weight_tensor(weight.type(dtype))
        fused = dict(weight=weight.type(dtype).clone(), float8_data=mod.float8_data.clone(), scale=mod.scale.clone(),
                     scale_reciprocal=mod.scale_reciprocal.clone())
        # ... and of remove_lora_from_module (:742-749) on the fused layer
        weight, _, _ = ref_lora.extract_weight_from_linear(mod)
        weight = ref_lora.unfuse_lora_weight_from_module(weight, lora_sd, lora_scale=ls)
        mod.set_weight_tensor(weight.type(dtype))
        unfused = dict(weight=weight.type(dtype).clone(), float8_data=mod.float8_data.clone(), scale=mod.scale.clone(),
                       scale_reciprocal=mod.scale_reciprocal.clone())
        ow, oq, os_, osr = O.lora_fuse_f8(before["float8_data"], before["scale_reciprocal"], lora_A, lora_B, alpha, ls)
        report(f"{name}: fused weight", fused["weight"], ow, 0, exact=True)
        report(f"{name}: fused float8_data", fused["float8_data"], oq, 0, exact=True)
        report(f"{name}: fused scale", fused["scale"], os_, 0, exact=True)
        uw, uq, us, usr = O.lora_fuse_f8(fused["float8_data"], fused["scale_reciprocal"], lora_A, lora_B, alpha, ls,
                                         unfuse=True)
        report(f"{name}: unfused weight", unfused["weight"], uw, 0, exact=True)
        report(f"{name}: unfused float8_data", unfused["float8_data"], uq, 0, exact=True)
        cases.append(dict(name=name, lora_A=lora_A, lora_B=lora_B, alpha=alpha, lora_scale=ls, before=before,
                          fused=fused, unfused=unfused))
    torch.save(cases, os.path.join(OUT, "lora.pt"))


VAE_TINY = dict(resolution=64, in_channels=3, ch=64, out_ch=3, ch_mult=[1, 2, 4, 4], num_res_blocks=2, z_channels=16,
                scale_factor=0.3611, shift_factor=0.1159)


def golden_vae():
    """tests/golden/vae_tiny.pt: the UNMODIFIED reference AutoEncoder.decode (fp32) on a seeded tiny configuration, the
    oracle pinned against it, and the oracle's CUDA-autocast restatement of the same decode (the kernels' CPU target)."""
    from modules import autoencoder as ref_ae  # noqa: E402  (reference)

    from oracle import vae_oracle as V

    print("vae_tiny.pt")
    ae = ref_ae.AutoEncoder(ref_ae.AutoEncoderParams(**VAE_TINY))
    sd = V.synthetic_state(ae, seed=31)                                      # decoder tensors (fingerprinted below)
    sd_all = V.synthetic_state(ae, seed=31, prefixes=("decoder.", "encoder."))  # + encoder tensors, decoder ones unchanged
    assert all(torch.equal(sd[k], sd_all[k]) for k in sd)
    missing, unexpected = ae.load_state_dict({k: v.float() for k, v in sd_all.items()}, strict=True)
    ae = ae.float().eval()
    g = torch.Generator().manual_seed(32)
    z = torch.randn(2, 16, 8, 8, generator=g) * 1.2
    with torch.inference_mode():
        y_ref = ae.decode(z)  # reference, fp32
        y_ora32 = V.decode(z, sd, VAE_TINY["ch_mult"], VAE_TINY["num_res_blocks"], VAE_TINY["scale_factor"],
                           VAE_TINY["shift_factor"], policy="fp32")
        y_auto = V.decode(z, sd, VAE_TINY["ch_mult"], VAE_TINY["num_res_blocks"], VAE_TINY["scale_factor"],
                          VAE_TINY["shift_factor"], policy="autocast")
        # intermediate taps of the reference for the kernel tests: mid-block input / output, fp32
        h0 = ae.decoder.conv_in(z / VAE_TINY["scale_factor"] + VAE_TINY["shift_factor"])
        h1 = ae.decoder.mid.block_1(h0)
        h2 = ae.decoder.mid.attn_1(h1)
        # encoder half (img2img): the Gaussian's moments of a seeded image, reference fp32 vs oracle
        img = torch.randn(2, 3, 64, 64, generator=g).clamp(-1, 1)
        m_ref = ae.encoder(img)
        m_ora32 = V.encoder(img, sd_all, VAE_TINY["ch_mult"], VAE_TINY["num_res_blocks"], policy="fp32")
        m_auto = V.encoder(img, sd_all, VAE_TINY["ch_mult"], VAE_TINY["num_res_blocks"], policy="autocast")
    dm = maxdiff(m_ref, m_ora32)
    print(f"  encoder: oracle(fp32) vs reference(fp32) max|d| {dm:.3g} on amax {m_ref.abs().max().item():.3g}; "
          f"oracle(autocast) mean|d| {(m_ref - m_auto.float()).abs().mean().item():.3g}")
    assert dm <= 2e-4 * max(1.0, m_ref.abs().max().item()), dm
    d = maxdiff(y_ref, y_ora32)
    print(f"  oracle(fp32) vs reference(fp32): max|d| {d:.3g} on amax {y_ref.abs().max().item():.3g}")
    assert d <= 2e-4 * max(1.0, y_ref.abs().max().item()), d
    da = maxdiff(y_ref, y_auto)
    print(f"  oracle(autocast) vs reference(fp32): max|d| {da:.3g}  mean|d| {(y_ref - y_auto.float()).abs().mean().item():.3g}")
    # the 12 M parameters are not stored: tests regenerate them with V.synthetic_state(<any module with the reference's
    # decoder keys>, seed=31) and check the fingerprint
    torch.save({"params": VAE_TINY, "state_seed": 31, "state_checksum": V.state_checksum(sd), "z": z, "y_ref_fp32": y_ref, "y_oracle_autocast": y_auto,
                "h_conv_in": h0, "h_mid_block_1": h1, "h_mid_attn_1": h2,
                "img": img, "moments_ref_fp32": m_ref, "moments_oracle_autocast": m_auto}, os.path.join(OUT, "vae_tiny.pt"))


def golden_text():
    """tests/golden/text_tiny.pt: outputs of the installed Hugging Face T5EncoderModel / CLIPTextModel (the third-party code
    the reference's HFEmbedder calls, conditioner.py:80-114) on tiny configurations with seeded weights, and the oracle
    pinned against them."""
    import transformers
    from transformers import CLIPTextConfig, CLIPTextModel, T5Config, T5EncoderModel

    from oracle import text_oracle as T

    print("text_tiny.pt   (transformers", transformers.__version__ + ")")
    t5 = T5EncoderModel(T5Config(**T.T5_TINY)).eval()
    sd_t5 = T.seeded_state(t5, 3)
    t5.load_state_dict(sd_t5, strict=False)
    g = torch.Generator().manual_seed(41)
    ids_t5 = torch.randint(0, T.T5_TINY["vocab_size"], (2, 40), generator=g)
    clip = CLIPTextModel(CLIPTextConfig(**T.CLIP_TINY)).eval()
    sd_clip = T.seeded_state(clip, 4)
    clip.load_state_dict(sd_clip, strict=False)
    ids_clip = torch.randint(3, T.CLIP_TINY["vocab_size"] - 1, (2, 77), generator=g)
    ids_clip[:, 0] = 0
    ids_clip[0, 19], ids_clip[0, 20:] = 511, 1   # end-of-text = the largest id, then padding
    ids_clip[1, 12], ids_clip[1, 13:] = 511, 1
    with torch.inference_mode():
        y_t5 = t5(input_ids=ids_t5, attention_mask=None, output_hidden_states=False)["last_hidden_state"]
        r = clip(input_ids=ids_clip, attention_mask=None, output_hidden_states=False)
        o_t5 = T.t5_encoder(sd_t5, T.T5_TINY, ids_t5)
        o_h, o_p = T.clip_text(sd_clip, T.CLIP_TINY, ids_clip)
    report("t5 last_hidden_state", y_t5, o_t5, 1e-5)
    report("clip last_hidden_state", r["last_hidden_state"], o_h, 1e-5)
    report("clip pooler_output", r["pooler_output"], o_p, 1e-5)
    torch.save({"transformers": transformers.__version__, "t5_seed": 3, "clip_seed": 4, "ids_t5": ids_t5, "ids_clip": ids_clip,
                "y_t5": y_t5, "y_clip_hidden": r["last_hidden_state"], "y_clip_pooled": r["pooler_output"],
                "t5_shapes": {k: (tuple(v.shape), str(v.dtype)) for k, v in sd_t5.items()},
                "clip_shapes": {k: (tuple(v.shape), str(v.dtype)) for k, v in sd_clip.items()}},
               os.path.join(OUT, "text_tiny.pt"))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    golden_quantize()
    golden_f8linear()
    golden_ops()
    golden_flux()
    golden_flux_variants()
    golden_lora()
    golden_vae()
    golden_text()
    print("golden fixtures written to", OUT)
