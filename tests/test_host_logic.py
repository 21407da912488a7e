"""CPU: host-side mirror of the reference interface -- state-dict format, error conventions, schedule and
input shaping, request cache, batch sharding arithmetic (no kernels are launched)."""
import os

import pytest
import torch
from torch import nn

from flux_fp8_api_b200 import model as M
from flux_fp8_api_b200 import parallel as PAR
from flux_fp8_api_b200 import pipeline as PL
from flux_fp8_api_b200.f8linear import F8Linear, mul_scale
from oracle import flux_oracle as O

BF16 = torch.bfloat16


@pytest.fixture(scope="module")
def tiny(golden_dir):
    return torch.load(os.path.join(golden_dir, "flux_tiny.pt"))


def tiny_spec(tiny, **kw):
    return M.FluxSpec(params=M.FluxParams(**tiny["tiny"]), **kw)


def test_state_dict_keys_match_the_reference(tiny):
    net = M.Flux(tiny_spec(tiny, prequantized_flow=True), dtype=BF16)
    missing, unexpected = net.load_state_dict(tiny["state"], strict=True)
    assert not missing and not unexpected
    assert set(net.state_dict().keys()) == set(tiny["state"].keys())
    assert PL.all_frozen(net)
    lin = net.double_blocks[0].img_attn.qkv
    assert isinstance(lin, F8Linear) and lin.float8_data.dtype == torch.float8_e4m3fn
    assert lin.weight.shape == (1,) and lin.weight.dtype == BF16  # placeholder defines the output dtype
    assert lin.input_scale.dtype == torch.float32 and lin.input_scale.dim() == 0
    assert lin.trial_index == lin.num_scale_trials


def test_unquantised_layout_has_plain_linears(tiny):
    net = M.Flux(tiny_spec(tiny, prequantized_flow=False), dtype=BF16)
    assert not any(isinstance(m, F8Linear) for m in net.modules())
    assert isinstance(net.double_blocks[0].img_mod.lin, nn.Linear)
    net2 = M.Flux(tiny_spec(tiny, prequantized_flow=True, quantize_modulation=False), dtype=BF16)
    assert isinstance(net2.double_blocks[0].img_mod.lin, nn.Linear)
    assert isinstance(net2.double_blocks[0].img_attn.qkv, F8Linear)
    assert isinstance(net2.img_in, nn.Linear)
    net3 = M.Flux(tiny_spec(tiny, prequantized_flow=True, quantize_flow_embedder_layers=True), dtype=BF16)
    assert isinstance(net3.img_in, F8Linear) and isinstance(net3.time_in.in_layer, F8Linear)
    assert isinstance(net3.final_layer.linear, nn.Linear)


def test_f8linear_load_cases(tiny):
    sd = {k[len("double_blocks.0.img_attn.proj."):]: v for k, v in tiny["state"].items()
          if k.startswith("double_blocks.0.img_attn.proj.")}
    full = F8Linear(256, 256, dtype=BF16)
    full.load_state_dict(sd)
    assert full.frozen
    # weight scales only -> input scale must be re-calibrated (reference float8_quantize.py:154-178)
    partial = {k: v for k, v in sd.items() if not k.startswith("input_scale")}
    lin = F8Linear(256, 256, dtype=BF16)
    lin.load_state_dict(partial, strict=False)
    assert lin.weight_initialized and not lin.input_scale_initialized and lin.trial_index == 0
    # malformed
    with pytest.raises(RuntimeError):
        F8Linear(256, 256, dtype=BF16).load_state_dict({"bias": sd["bias"]}, strict=False)
    bad = dict(sd)
    bad["float8_data"] = sd["float8_data"][:128]
    with pytest.raises(RuntimeError):
        F8Linear(256, 256, dtype=BF16).load_state_dict(bad)


def test_prequantized_checkpoint_round_trip(tiny, tmp_path):
    """save_prequantized -> load_prequantized reproduces every tensor bit for bit (fp8 bytes, scales, biases),
    carries the model spec in its header, and refuses un-calibrated models / foreign files."""
    spec = tiny_spec(tiny, prequantized_flow=True)
    net = M.Flux(spec, dtype=BF16)
    net.load_state_dict(tiny["state"], strict=True)
    path = str(tmp_path / "tiny.f8.pt")
    header = PL.save_prequantized(net, path, spec)
    assert header["format"] == PL.PREQUANTIZED_FORMAT and header["f8_layers"] == 13
    back = PL.load_prequantized(path, "cpu")
    assert PL.all_frozen(back)
    a, b = net.state_dict(), back.state_dict()
    assert set(a) == set(b)
    for k in a:
        if a[k] is None:
            assert b[k] is None
            continue
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
        assert torch.equal(a[k].view(torch.uint8) if a[k].dtype == torch.float8_e4m3fn else a[k],
                           b[k].view(torch.uint8) if b[k].dtype == torch.float8_e4m3fn else b[k]), k
    assert back.params.depth == net.params.depth and isinstance(back.double_blocks[0].img_mod.lin, F8Linear)
    # an un-calibrated model must not be written
    fresh = M.Flux(spec, dtype=BF16)
    with pytest.raises(RuntimeError):
        PL.save_prequantized(fresh, str(tmp_path / "x.pt"), spec)
    torch.save({"header": {"format": "something-else"}, "state": {}}, str(tmp_path / "y.pt"))
    with pytest.raises(RuntimeError):
        PL.load_prequantized(str(tmp_path / "y.pt"), "cpu")


def test_flux_constructor_errors():
    with pytest.raises(ValueError):
        M.Flux(M.FluxSpec(params=M.FluxParams(hidden_size=250, num_heads=4)))
    with pytest.raises(ValueError):
        M.Flux(M.FluxSpec(params=M.FluxParams(hidden_size=256, num_heads=2, axes_dim=[16, 56, 40], depth=0,
                                              depth_single_blocks=0)))


def test_scale_semantics_switch():
    import flux_fp8_api_b200.f8linear as f8

    s = torch.tensor(57344.0 / 5.0)
    assert f8.SCALE_SEMANTICS == "cuda"
    assert mul_scale(s).item() == s.to(BF16).float().item() != s.item()
    f8.SCALE_SEMANTICS = "cpu"
    try:
        assert mul_scale(s) is s
    finally:
        f8.SCALE_SEMANTICS = "cuda"
    assert mul_scale(torch.tensor(448.0)).item() == 448.0  # weight scale of real checkpoints is bf16-exact


def test_schedule_and_input_shaping_match_the_oracle():
    for steps, L, shift in [(28, 4096, True), (4, 4096, False), (50, 9216, True), (13, 64, True)]:
        assert PL.get_schedule(steps, L, shift=shift) == O.get_schedule(steps, L, shift=shift)
    ids = PL.make_img_ids(2, 6, 10, "cpu")
    assert torch.equal(ids, O.make_img_ids(2, 6, 10, BF16))
    lat = torch.arange(2 * 16 * 4 * 6, dtype=torch.float32).reshape(2, 16, 4, 6)
    tok = PL.patchify(lat)
    assert tok.shape == (2, 6, 64)
    assert torch.equal(tok[0, 0].reshape(16, 2, 2), lat[0, :, 0:2, 0:2])
    params = M.FluxParams()
    req = PL.synthetic_request(params, 64, 48, 3, 32, "cpu", seed=5)
    assert req["img"].shape == (3, 12, 64) and req["txt"].shape == (3, 32, 4096) and req["y"].shape == (3, 768)
    shard = PL.synthetic_request(params, 64, 48, 1, 32, "cpu", seed=5, sample_offset=2)
    assert torch.equal(shard["img"][0], req["img"][2]) and torch.equal(shard["txt"][0], req["txt"][2])


def test_step_invariant_cache_never_aliases():
    c = M._StepInvariantCache()
    calls = []
    a = torch.zeros(3)
    f = lambda: calls.append(1) or len(calls)
    assert c.get("k", (a,), f) == 1 and c.get("k", (a,), f) == 1
    a.add_(1)  # in-place change -> recompute
    assert c.get("k", (a,), f) == 2
    b = torch.zeros(3)  # a different tensor object -> recompute even if it had the same address/shape
    assert c.get("k", (b,), f) == 3


def test_shard_ranges_partition_the_batch():
    for total in range(0, 20):
        for world in (1, 2, 3, 4, 8):
            spans = [PAR.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    req = {"img": torch.arange(8)[:, None], "txt": torch.arange(8)[:, None] * 10, "note": 3}
    got = torch.cat([PAR.shard_request(req, r, 4)["img"] for r in range(4)])
    assert torch.equal(got, req["img"]) and PAR.shard_request(req, 1, 4)["note"] == 3


def test_lora_operands_reproduce_the_reference_delta(golden_dir):
    """lora.lora_operands (the factors handed to fluxb200_lora_fuse) multiplied out on the CPU equal the oracle's
    calculate_lora_weight restatement for every golden case (alpha scaling, uneven-rank chunking), and the host
    mirror keeps the reference's key helpers' behaviour."""
    from flux_fp8_api_b200 import lora

    cases = torch.load(os.path.join(golden_dir, "lora.pt"))
    for c in cases:
        down, up, chunks = lora.lora_operands((c["lora_A"], c["lora_B"], c["alpha"]), None, "cpu")
        assert down.dtype == torch.float32 and up.dtype == torch.float32
        delta = torch.zeros(down.shape[0], up.shape[1])
        for part in up.chunk(chunks, dim=0):
            delta = delta + (c["lora_scale"] * torch.mm(down, part))
        assert torch.equal(delta, O.lora_delta(c["lora_A"], c["lora_B"], c["alpha"], None, c["lora_scale"])), c["name"]
    w = {"a.b.lora_A.weight": torch.ones(2, 4), "a.b.lora_B.weight": torch.ones(3, 2), "a.b.alpha": 2.0,
         "c.lora_A.weight": torch.ones(2, 4)}
    assert lora.get_lora_for_key("a.b", w)[2] == 2.0
    assert lora.get_lora_for_key("c", w) is None          # lora_B missing -> skipped, as in the reference
    assert sorted(lora._keys_without_ab(w)) == ["a.b", "c"]
    with pytest.raises(ValueError):
        lora.lora_operands((torch.ones(5, 4), torch.ones(3, 2), None), None, "cpu")   # 5 rows do not chain with rank 2
    with pytest.raises(FileNotFoundError):   # a path is loaded as a BFL-layout .safetensors file
        lora.apply_lora_to_model(torch.nn.Linear(2, 2), "some/file.safetensors")


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the arm the driver runs beside ours): one JSON line with the contract's keys, the
    reference's own modules when oracle/_ref is staged, --steps / --warmup honoured, config identical to our arm's."""
    import json
    import subprocess
    import sys

    import bench

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FLUXB200_BENCH_THREADS=str(min(8, os.cpu_count() or 1)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "c4",
                          "--steps", "2", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 0
    assert line["metric"] == bench.metric_name(bench.CONFIGS["c4"]) and line["unit"] == "it/s" and line["higher_is_better"]
    assert line["config"] == bench.config_block(bench.CONFIGS["c4"], "c4", 1)
    cb = line["cpu_baseline"]
    assert cb["value"] == line["value"] and cb["cores"] >= 1 and cb["kind"] in ("reference", "port")
    from oracle import ref_loader as R

    assert cb["kind"] == ("reference" if R.available() else "port")
    assert line["e2e"] == {"value": line["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(line["value"] - 1000.0 / (19.0 * line["ms_per_step"])) < 1e-9 * line["value"] + 1e-12
    # work accounting used by our arm
    f8, attn, mod = bench.algorithmic_flops(bench.CONFIGS["c2"])
    assert abs(f8 - (1.29101e10 * 4608 + 6.456e9)) < 1 and abs(attn - 700416.0 * 4608 ** 2) < 1 and mod == 0.0
    assert bench.algorithmic_flops(bench.CONFIGS["c5"])[2] == 6.456e9


def test_flux_from_pretrained_and_lora_bookkeeping(tmp_path, golden_dir):
    """Flux.from_pretrained (reference modules/flux_model.py:718-734): a JSON spec whose ckpt_path names a safetensors
    state dict -- bf16 master weights and a prequantised flow -- comes back on the CPU with the same keys and values
    (meta-device construction + load_state_dict(assign=True)); get_lora / has_lora follow the reference's bookkeeping."""
    import json

    from safetensors.torch import save_file

    from flux_fp8_api_b200 import lora as L, model as M

    gold = torch.load(os.path.join(golden_dir, "flux_tiny.pt"))
    # (1) prequantised flow: the reference-minted golden state
    ckpt = str(tmp_path / "tiny.f8.safetensors")
    save_file({k: v.contiguous() for k, v in gold["state"].items()}, ckpt)
    cfg = tmp_path / "tiny.json"
    cfg.write_text(json.dumps({"params": gold["tiny"], "prequantized_flow": True, "ckpt_path": ckpt,
                               "version": "flux-dev", "flow_dtype": "bfloat16"}))   # extra reference fields are ignored
    net = M.Flux.from_pretrained(str(cfg), dtype=torch.bfloat16)
    sd = net.state_dict()
    assert set(sd) == set(gold["state"])
    for k, v in gold["state"].items():
        assert sd[k].dtype == v.dtype and sd[k].device.type == "cpu"
        a, b = (sd[k].view(torch.uint8), v.view(torch.uint8)) if v.dtype.itemsize == 1 else (sd[k], v)
        assert torch.equal(a, b), k
    lin = net.double_blocks[0].img_attn.qkv
    assert lin.frozen and lin.float8_data.dtype == torch.float8_e4m3fn
    # (2) un-quantised bf16 flow
    with torch.device("cpu"):
        ref = M.Flux(M.FluxSpec(params=M.FluxParams(**gold["tiny"])), dtype=torch.bfloat16).to(torch.bfloat16)
    ckpt2 = str(tmp_path / "tiny.bf16.safetensors")
    save_file({k: v.contiguous() for k, v in ref.state_dict().items()}, ckpt2)
    cfg2 = tmp_path / "tiny_bf16.json"
    cfg2.write_text(json.dumps({"params": gold["tiny"], "prequantized_flow": False, "ckpt_path": ckpt2}))
    net2 = M.Flux.from_pretrained(str(cfg2), dtype=torch.bfloat16)
    assert all(torch.equal(v, ref.state_dict()[k]) for k, v in net2.state_dict().items())
    with pytest.raises(ValueError):
        M.Flux.from_pretrained(str(tmp_path / "missing.json"))
    # LoRA bookkeeping (no device work): identifiers resolve by path or name, as in the reference
    net.loras.append(L.LoraWeights({}, "/x/style.safetensors", None, 0.7))
    assert net.has_lora("style.safetensors") and net.get_lora("/x/style.safetensors").scale == 0.7
    assert not net.has_lora("other") and net.get_lora("other") is None
    assert net.unload_lora("other") is False


def test_dispatch_predicates_refuse_what_the_kernels_cannot_take():
    """The own-kernel paths of the embedders / final layer are taken only for plain bf16 CUDA linears of supported
    shapes; everything else (CPU tensors, fp32 layers, F8Linear embedders, K % 32 != 0) stays on the generic path."""
    from flux_fp8_api_b200 import model as M

    lin = torch.nn.Linear(64, 32).to(torch.bfloat16)
    x = torch.zeros(2, 64, dtype=torch.bfloat16)
    assert not M._skinny_bf16(x, lin) and not M._small_bf16_linear(x, lin)          # CPU tensors
    assert not M._skinny_bf16(torch.zeros(2, 64), torch.nn.Linear(64, 32))             # fp32
    assert not M._small_bf16_linear(torch.zeros(2, 48, dtype=torch.bfloat16), torch.nn.Linear(48, 32).to(torch.bfloat16))


def test_vae_host_side_keys_packing_and_loud_failure():
    """AutoEncoder mirrors the reference's decoder state dict (138 tensors for Flux's VAE), packs Conv2d weights into the
    kernel layout, and refuses to compute without the CUDA library path (no CPU fallback)."""
    from flux_fp8_api_b200 import autoencoder as A
    from flux_fp8_api_b200 import ops

    m = A.AutoEncoder(A.AutoEncoderParams(resolution=256, in_channels=3, ch=128, out_ch=3, ch_mult=[1, 2, 4, 4], num_res_blocks=2,
                                          z_channels=16, scale_factor=0.3611, shift_factor=0.1159))
    keys = set(m.state_dict())
    assert len(keys) == 244 and sum(k.startswith("decoder.") for k in keys) == 138 and sum(k.startswith("encoder.") for k in keys) == 106
    for k in ("decoder.conv_in.weight", "decoder.mid.attn_1.q.weight", "decoder.mid.block_1.norm1.weight",
              "decoder.up.0.block.0.nin_shortcut.weight", "decoder.up.3.upsample.conv.bias", "decoder.up.1.block.2.conv2.weight",
              "decoder.norm_out.bias", "decoder.conv_out.weight"):
        assert k in keys, k
    assert "decoder.up.0.upsample.conv.weight" not in keys  # the highest resolution level has no upsampler
    assert m.state_dict()["decoder.conv_in.weight"].shape == (512, 16, 3, 3)
    assert m.state_dict()["decoder.up.1.block.0.nin_shortcut.weight"].shape == (256, 512, 1, 1)

    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    pk = ops.pack_conv_weight(w)
    assert pk.shape == (2, 9 * 64) and pk.dtype == torch.bfloat16
    for n in range(2):
        for ky in range(3):
            for kx in range(3):
                row = pk[n, (ky * 3 + kx) * 64:(ky * 3 + kx + 1) * 64].float()
                assert torch.equal(row[:3], w[n, :, ky, kx]) and not row[3:].any()
    with pytest.raises(ValueError):
        ops.pack_conv_weight(torch.zeros(4, 4, 5, 5))
    if not torch.cuda.is_available():
        with pytest.raises(Exception) as ei:
            m.decode(torch.zeros(1, 16, 8, 8))
        assert "CUDA" in str(ei.value) or "cuda" in str(ei.value)
    assert m.state_dict()["encoder.down.1.downsample.conv.weight"].shape == (256, 256, 3, 3) and m.encoder.down[1].downsample.conv.stride == (2, 2)
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            m.encode(torch.zeros(1, 3, 64, 64))


def test_unpack_inverts_patchify_like_the_reference_rearrange():
    """pipeline.unpack == einops `b (h w) (c ph pw) -> b c (h ph) (w pw)` (flux_pipeline.py:440-448)."""
    from einops import rearrange

    from flux_fp8_api_b200 import pipeline as PL

    lat = torch.randn(2, 16, 12, 20)
    tok = PL.patchify(lat)
    assert torch.equal(PL.unpack(tok, 96, 160), lat)
    assert torch.equal(PL.unpack(tok, 96, 160), rearrange(tok, "b (h w) (c ph pw) -> b c (h ph) (w pw)", h=6, w=10, ph=2, pw=2))
    with pytest.raises(ValueError):
        PL.unpack(tok, 96, 176)


def test_synthetic_vae_weights_match_the_oracle_generator():
    """The package's seeded VAE parameters (bench.py's own arm) are the oracle's (the reference arm's): same tensors."""
    from flux_fp8_api_b200 import autoencoder as A, pipeline as PL
    from oracle import vae_oracle as V

    p = dict(resolution=64, in_channels=3, ch=64, out_ch=3, ch_mult=[1, 2], num_res_blocks=1, z_channels=16, scale_factor=1.0,
             shift_factor=0.0)
    ae = A.AutoEncoder(A.AutoEncoderParams(**p))
    want = V.synthetic_state(ae, seed=5)
    PL.init_synthetic_vae_weights(ae, seed=5)
    got = ae.state_dict()
    assert set(want) == {k for k in got if k.startswith("decoder.")} and all(torch.equal(got[k].to(torch.bfloat16), want[k]) for k in want)
    want_all = V.synthetic_state(ae, seed=5, prefixes=("decoder.", "encoder."))
    PL.init_synthetic_vae_weights(ae, seed=5, prefixes=("decoder.", "encoder."))
    got = ae.state_dict()
    assert set(want_all) == set(got) and all(torch.equal(got[k].to(torch.bfloat16), want_all[k]) for k in want_all)
