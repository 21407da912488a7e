"""CPU: the oracle (oracle/flux_oracle.py) against the committed reference outputs in tests/golden/
(minted from the unmodified reference by oracle/make_golden.py).  No /root/reference needed."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import flux_oracle as O

BF16 = torch.bfloat16


def load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name))


def dt_of(s):
    return torch.float8_e5m2 if "e5m2" in s else torch.float8_e4m3fn


def maxdiff(a, b):
    return (a.float() - b.float()).abs().max().item()


def test_quantize_is_bit_exact(golden_dir):
    g = load(golden_dir, "quantize.pt")
    assert len(g["cases"]) == 12
    for c in g["cases"]:
        dt = dt_of(c["dtype"])
        scale = O.amax_to_scale(torch.tensor(c["amax"], dtype=torch.float32), torch.finfo(dt).max)
        assert torch.equal(scale, c["scale"]), c
        assert torch.equal(O.quantize(g["x"], c["scale"], dt).view(torch.uint8), c["y"]), c


def test_scale_is_clamped_to_fp8_max():
    # tensors with amax < 1 get the constant scale 448 / 57344, not "full range"
    assert O.amax_to_scale(torch.tensor(0.05), 448.0).item() == 448.0
    assert O.amax_to_scale(torch.tensor(0.0), 57344.0).item() == 57344.0
    assert O.amax_to_scale(torch.tensor(4.0), 448.0).item() == 112.0


def test_quantize_double_rounding_differs_from_single():
    # fp8(bf16(x*s)) != fp8(x*s) for some inputs; the oracle must do the former
    x = torch.linspace(-3, 3, 20001).to(BF16)
    s = torch.tensor(57344.0 / 5.0)
    double = O.quantize(x, s, torch.float8_e5m2).float()
    single = (x.float() * s).clamp(-57344, 57344).to(torch.float8_e5m2).float()
    assert (double != single).any()


def test_f8linear_and_calibration(golden_dir):
    for c in load(golden_dir, "f8linear.pt"):
        in_dt = dt_of(c["in_dtype"])
        sd = c["state"]
        wq, ws, wsr = O.quantize_weight(c["weight_bf16"])
        assert torch.equal(wq.view(torch.uint8), sd["float8_data"].view(torch.uint8))
        assert torch.equal(ws, sd["scale"]) and torch.equal(wsr, sd["scale_reciprocal"])
        # calibration: frozen scale = amax_to_scale(max of the first 12 call amaxes)
        amax12 = torch.tensor(max(c["x_all_amax"][:12]), dtype=torch.float32)
        assert torch.equal(O.amax_to_scale(amax12, torch.finfo(in_dt).max), sd["input_scale"])
        y = O.f8linear(c["x_last"], sd, "", in_dt)
        tol = 2.0 ** -7 * max(1.0, c["y_last"].abs().max().item())
        assert maxdiff(y, c["y_last"]) <= tol
        assert ((y.float() - c["y_last"].float()) != 0).float().mean().item() < 0.005


def test_calibrating_linear_trace():
    cal = O.CalibratingLinear(12)
    g = torch.Generator().manual_seed(0)
    amaxes = []
    for i in range(14):
        x = (torch.randn(64, 32, generator=g) * (1 + i % 5)).to(BF16)
        amaxes.append(x.abs().max().float())
        cal.quantize_input(x)
        if i < 12:
            assert not cal.initialized
            assert torch.equal(cal.input_scale, O.amax_to_scale(torch.stack(amaxes).max(), 57344.0))
    assert cal.initialized and cal.index == 12
    assert torch.equal(cal.input_scale, O.amax_to_scale(torch.stack(amaxes[:12]).max(), 57344.0))


def test_ops_against_reference(golden_dir):
    g = load(golden_dir, "ops.pt")
    pe = O.embed_nd(g["ids"], [16, 56, 56], 10_000, BF16)
    assert torch.equal(pe, g["pe"])
    q, k = O.apply_rope(g["q"], g["k"], g["pe"])
    assert torch.equal(q, g["q_rope"]) and torch.equal(k, g["k_rope"])
    assert maxdiff(O.rms_norm(g["q"], g["qnorm_w"]), g["q_norm"]) <= 2.0 ** -7
    assert maxdiff(O.rms_norm(g["k"], g["knorm_w"]), g["k_norm"]) <= 2.0 ** -7
    assert maxdiff(O.attention(g["q"], g["k"], g["v"], g["pe"]), g["attention"]) <= 2.0 ** -6
    assert maxdiff(O.layernorm_modulate(g["x"], g["shift"], g["scale"]), g["ln_mod"]) <= 2.0 ** -5
    assert torch.equal(O.timestep_embedding(g["t"], 256), g["t_emb"])
    assert torch.equal(F.gelu(g["x"], approximate="tanh"), g["gelu"]) and torch.equal(F.silu(g["x"]), g["silu"])


def test_rope_is_identity_for_text_positions(golden_dir):
    g = load(golden_dir, "ops.pt")
    q, _ = O.apply_rope(g["q"], g["k"], g["pe"])
    assert torch.equal(q[:, :, :32], g["q"][:, :, :32])  # ids == 0 -> cos 1, sin 0


def test_blocks_and_forward_against_reference(golden_dir):
    g = load(golden_dir, "flux_tiny.pt")
    sd, cfg, bi = g["state"], g["cfg"], g["block_in"]
    heads = g["tiny"]["num_heads"]
    img, txt = O.double_block(bi["img"], bi["txt"], bi["vec"], bi["pe"], sd, "double_blocks.0.", heads)
    assert maxdiff(img, g["double_img"]) <= 2.0 ** -4 and maxdiff(txt, g["double_txt"]) <= 2.0 ** -4
    xs = torch.cat((bi["txt"], bi["img"]), 1)
    assert maxdiff(O.single_block(xs, bi["vec"], bi["pe"], sd, "single_blocks.0.", heads), g["single"]) <= 2.0 ** -4
    (sh, sc, ga), _ = O.modulation(bi["vec"], sd, "double_blocks.0.img_mod.", True)
    for ours, ref in zip((sh, sc, ga), g["mod1"]):
        assert maxdiff(ours, ref) <= 2.0 ** -8
    y = O.flux_forward(sd, cfg, **g["inputs"])
    assert maxdiff(y, g["y_fp8"]) <= 2.0 ** -4
    assert maxdiff(g["y_fp8"], g["y_bf16"]) < 0.1  # the fp8 path tracks the bf16 path


def test_e4m3_activations_and_quantised_embedders_against_reference(golden_dir):
    """flux_tiny_e4m3.pt: reference quantised with input_float8_dtype=float8_e4m3fn and
    quantize_flow_embedder_layers=True (float8_quantize.py:298-304, 447-484)."""
    g = load(golden_dir, "flux_tiny_e4m3.pt")
    E4M3 = torch.float8_e4m3fn
    sd, cfg, bi = g["state"], g["cfg"], g["block_in"]
    heads = g["tiny"]["num_heads"]
    assert g["n_f8"] == 21 and "img_in.float8_data" in sd and "time_in.in_layer.float8_data" in sd
    assert "final_layer.linear.float8_data" not in sd
    img, txt = O.double_block(bi["img"], bi["txt"], bi["vec"], bi["pe"], sd, "double_blocks.0.", heads, E4M3)
    assert maxdiff(img, g["double_img"]) <= 2.0 ** -4 and maxdiff(txt, g["double_txt"]) <= 2.0 ** -4
    xs = torch.cat((bi["txt"], bi["img"]), 1)
    assert maxdiff(O.single_block(xs, bi["vec"], bi["pe"], sd, "single_blocks.0.", heads, E4M3), g["single"]) <= 2.0 ** -4
    assert maxdiff(O.flux_forward(sd, cfg, **g["inputs"], in_dtype=E4M3), g["y_fp8"]) <= 2.0 ** -4
    # the e5m2 evaluation of the same state is a DIFFERENT function: the fixture really pins the activation format
    assert maxdiff(O.flux_forward(sd, cfg, **g["inputs"]), g["y_fp8"]) > 0


def test_four_step_trajectory_against_reference(golden_dir):
    """flux_tiny_traj.pt: the Euler loop of flux_pipeline.py:627-651 over the reference Flux.forward, 4 steps."""
    g = load(golden_dir, "flux_tiny_traj.pt")
    model = load(golden_dir, "flux_tiny.pt")
    inp, ts = g["inputs"], g["timesteps"]
    assert len(ts) == 5 and len(g["latents"]) == 4
    img = inp["img"]
    for i, (t_curr, t_prev) in enumerate(zip(ts[:-1], ts[1:])):
        t_vec = torch.full((img.shape[0],), t_curr, dtype=BF16)
        pred = O.flux_forward(model["state"], model["cfg"], img, inp["img_ids"], inp["txt"], inp["txt_ids"], t_vec,
                              inp["y"], inp["guidance"])
        # teacher-forced per step (the fixture's latent is the next input), so the bound does not compound
        assert maxdiff(pred, g["preds"][i]) <= 2.0 ** -4, i
        assert maxdiff(O.euler_step(g["latents"][i - 1] if i else inp["img"], g["preds"][i], t_curr, t_prev),
                       g["latents"][i]) == 0.0, i
        img = g["latents"][i]


def test_schedule_matches_reference_formula():
    ts = O.get_schedule(28, 4096)
    assert len(ts) == 29 and ts[0] == 1.0 and ts[-1] == 0.0
    assert all(a > b for a, b in zip(ts[:-1], ts[1:]))
    lin = O.get_schedule(4, 4096, shift=False)
    assert lin == pytest.approx([1.0, 0.75, 0.5, 0.25, 0.0])


def test_lora_fuse_and_unfuse_are_bit_exact(golden_dir):
    """oracle.lora_fuse_f8 against the reference's extract_weight_from_linear -> apply / unfuse ->
    set_weight_tensor chain (lora_loading.py), incl. alpha != rank and the uneven-rank chunked fuse."""
    cases = load(golden_dir, "lora.pt")
    assert [c["name"] for c in cases] == ["even", "alpha", "uneven", "big-delta"]
    for c in cases:
        w, q, s, sr = O.lora_fuse_f8(c["before"]["float8_data"], c["before"]["scale_reciprocal"], c["lora_A"],
                                     c["lora_B"], c["alpha"], c["lora_scale"])
        assert torch.equal(w, c["fused"]["weight"]), c["name"]
        assert torch.equal(q.view(torch.uint8), c["fused"]["float8_data"].view(torch.uint8)), c["name"]
        assert torch.equal(s, c["fused"]["scale"]) and torch.equal(sr, c["fused"]["scale_reciprocal"])
        w, q, s, sr = O.lora_fuse_f8(c["fused"]["float8_data"], c["fused"]["scale_reciprocal"], c["lora_A"],
                                     c["lora_B"], c["alpha"], c["lora_scale"], unfuse=True)
        assert torch.equal(w, c["unfused"]["weight"]), c["name"]
        assert torch.equal(q.view(torch.uint8), c["unfused"]["float8_data"].view(torch.uint8)), c["name"]
        # the LoRA really changed the layer, and un-fusing brings it back to within the e4m3 requantisation step
        assert not torch.equal(c["fused"]["float8_data"].view(torch.uint8), c["before"]["float8_data"].view(torch.uint8))
        w0 = c["before"]["float8_data"].float() * c["before"]["scale_reciprocal"]
        w2 = c["unfused"]["float8_data"].float() * c["unfused"]["scale_reciprocal"]
        assert (w0 - w2).abs().max() <= 0.13 * w0.abs().max()


def test_vae_oracle_reproduces_the_reference_decode(golden_dir):
    """oracle/vae_oracle.py (fp32 policy) against the UNMODIFIED reference AutoEncoder.decode output committed in
    tests/golden/vae_tiny.pt; the CUDA-autocast policy stays within bf16 distance of it."""
    from flux_fp8_api_b200 import autoencoder as A
    from oracle import vae_oracle as V

    g = torch.load(os.path.join(golden_dir, "vae_tiny.pt"))
    p = g["params"]
    shapes = A.AutoEncoder(A.AutoEncoderParams(**p))  # only its state-dict keys / shapes are used
    sd = V.synthetic_state(shapes, g["state_seed"])
    assert V.state_checksum(sd) == g["state_checksum"]
    y32 = V.decode(g["z"], sd, p["ch_mult"], p["num_res_blocks"], p["scale_factor"], p["shift_factor"], policy="fp32")
    assert (y32 - g["y_ref_fp32"]).abs().max().item() <= 2e-4 * g["y_ref_fp32"].abs().max().item()
    ya = V.decode(g["z"], sd, p["ch_mult"], p["num_res_blocks"], p["scale_factor"], p["shift_factor"], policy="autocast")
    assert torch.equal(ya, g["y_oracle_autocast"])
    assert (ya.float() - g["y_ref_fp32"]).abs().mean().item() <= 0.01 * g["y_ref_fp32"].abs().max().item()
    # encoder half: the Gaussian's moments of the reference Encoder (fp32)
    sd_all = V.synthetic_state(shapes, g["state_seed"], prefixes=("decoder.", "encoder."))
    m32 = V.encoder(g["img"], sd_all, p["ch_mult"], p["num_res_blocks"], policy="fp32")
    assert (m32 - g["moments_ref_fp32"]).abs().max().item() <= 2e-4 * g["moments_ref_fp32"].abs().max().item()
    assert torch.equal(V.encoder(g["img"], sd_all, p["ch_mult"], p["num_res_blocks"], policy="autocast"), g["moments_oracle_autocast"])


def test_text_encoder_oracle_reproduces_the_hugging_face_modules(golden_dir):
    """oracle/text_oracle.py against the outputs of the installed transformers T5EncoderModel / CLIPTextModel committed in
    tests/golden/text_tiny.pt (the third-party arithmetic behind the reference's HFEmbedder, conditioner.py:80-114)."""
    from oracle import text_oracle as T

    g = torch.load(os.path.join(golden_dir, "text_tiny.pt"))
    sd = T.seeded_state_from_shapes(g["t5_shapes"], g["t5_seed"])
    y = T.t5_encoder(sd, T.T5_TINY, g["ids_t5"])
    assert (y - g["y_t5"]).abs().max().item() <= 1e-5 * g["y_t5"].abs().max().item()
    sd = T.seeded_state_from_shapes(g["clip_shapes"], g["clip_seed"])
    h, p = T.clip_text(sd, T.CLIP_TINY, g["ids_clip"])
    assert (h - g["y_clip_hidden"]).abs().max().item() <= 1e-5 * g["y_clip_hidden"].abs().max().item()
    assert (p - g["y_clip_pooled"]).abs().max().item() <= 1e-5 * g["y_clip_pooled"].abs().max().item()
    # the relative-position buckets of T5 (32 buckets, distances up to 128) at the boundaries
    rel = torch.tensor([[-200, -128, -17, -16, -8, -7, -1, 0, 1, 7, 8, 16, 17, 127, 128, 200]])
    assert T.t5_relative_position_bucket(rel).tolist() == [[15, 15, 10, 10, 8, 7, 1, 0, 17, 23, 24, 26, 26, 31, 31, 31]]
