"""GPU (-m gpu): the text-encoder path (second half) against torch references of its kernels and
against the Hugging Face modules the reference's HFEmbedder wraps (modules/conditioner.py:80-114), run on the same B200."""
import json
import math
import os

import pytest
import torch
import torch.nn.functional as F

from flux_fp8_api_b200 import conditioner as CD
from flux_fp8_api_b200 import ops
from oracle import text_oracle as T

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ULP = 2.0 ** -7


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(shape, device=DEV, generator=g) * scale).to(BF16)


@pytest.mark.parametrize("rows,D", [(77, 768), (512, 4096), (5, 256)])
def test_rows_norm_t5_and_layernorm(rows, D):
    x, w, b = _rand((rows, D), 1, 2.0) + 0.3, (1 + 0.1 * _rand((D,), 2).float()).to(BF16), _rand((D,), 3, 0.1)
    y = ops.rows_norm(x, w, None, 1e-6)
    xf = x.float()
    ref = w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).to(BF16)  # T5LayerNorm.forward, op by op
    d = (y.float() - ref.float()).abs()
    assert d.max().item() <= ULP * ref.float().abs().max().item() and (d > 0).float().mean().item() < 0.01
    y2 = ops.rows_norm(x, w, b, 1e-5)
    ref2 = F.layer_norm(xf, (D,), w.float(), b.float(), 1e-5).to(BF16)
    d2 = (y2.float() - ref2.float()).abs()
    assert d2.max().item() <= ULP * ref2.float().abs().max().item() and (d2 > 0).float().mean().item() < 0.02


def test_gated_act_modes():
    u = _rand((100, 2 * 512), 4, 2.0)
    a, b = u[:, :512].float(), u[:, 512:].float()
    g = ops.gated_act(u, 512, 0)
    ref = (T.gelu_new(a).to(BF16).float() * b).to(BF16)
    d = (g.float() - ref.float()).abs()
    assert d.max().item() <= 2 * ULP * ref.float().abs().max().item() and (d > 0).float().mean().item() < 0.02
    qg = ops.gated_act(u, 512, 1)
    refq = (a * torch.sigmoid(1.702 * a)).to(BF16)
    dq = (qg.float() - refq.float()).abs()
    assert dq.max().item() <= ULP * refq.float().abs().max().item() and (dq > 0).float().mean().item() < 0.02


@pytest.mark.parametrize("B,S,H,bias,scale,causal", [(1, 512, 64, True, 1.0, False), (2, 77, 12, False, 0.125, True),
                                                     (3, 100, 4, True, 1.0, False), (1, 64, 2, False, 0.125, False)])
def test_attention_d64_against_the_eager_formula(B, S, H, bias, scale, causal):
    qkv = _rand((B * S, 3 * H * 64), 5)
    bt = _rand((H, S, S), 6) if bias else None
    out = ops.attention_d64(qkv, B, S, H, bt, scale, causal)
    q, k, v = (qkv[:, i * H * 64:(i + 1) * H * 64].reshape(B, S, H, 64).transpose(1, 2) for i in range(3))
    s = torch.matmul(q, k.transpose(-1, -2))  # bf16 result, as the eager modules compute it
    if scale != 1.0:
        s = s * scale
    if bias:
        s = s + bt[None]
    if causal:
        s = s + torch.full((S, S), float("-inf"), device=DEV).triu(1).to(BF16)
    p = torch.softmax(s.float(), -1)
    ref = (p @ v.float()).transpose(1, 2).reshape(B * S, H * 64)
    d = (out.float() - ref).abs()
    assert d.max().item() <= 3 * ULP * ref.abs().max().item(), d.max().item()
    assert d.mean().item() <= 0.35 * ULP * ref.abs().max().item()


def _hf(kind, cfg, seed, dtype=BF16):
    from transformers import CLIPTextConfig, CLIPTextModel, T5Config, T5EncoderModel

    m = (T5EncoderModel(T5Config(**cfg)) if kind == "t5" else CLIPTextModel(CLIPTextConfig(**cfg))).eval()
    m.load_state_dict(T.seeded_state(m, seed), strict=False)
    return m.to(DEV, dtype)


def test_tiny_encoders_against_hugging_face_and_the_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "text_tiny.pt"))
    t5 = _hf("t5", T.T5_TINY, g["t5_seed"])
    ours = CD.accelerate(t5)
    ids = g["ids_t5"].to(DEV)
    with torch.inference_mode():
        y = ours(input_ids=ids, attention_mask=None, output_hidden_states=False)["last_hidden_state"].float().cpu()
        y_hf = t5(input_ids=ids, attention_mask=None, output_hidden_states=False)["last_hidden_state"].float().cpu()
    floor = (y_hf - g["y_t5"]).abs().mean().item()
    mine = (y - g["y_t5"]).abs().mean().item()
    print(f"tiny T5: ours-vs-fp32 {mine:.4g}, HF-bf16-vs-fp32 {floor:.4g}, ours-vs-HF-bf16 {(y - y_hf).abs().mean().item():.4g}")
    assert y.shape == g["y_t5"].shape and mine <= 1.25 * floor

    clip = _hf("clip", T.CLIP_TINY, g["clip_seed"])
    oursc = CD.accelerate(clip)
    idc = g["ids_clip"].to(DEV)
    with torch.inference_mode():
        r = oursc(input_ids=idc, attention_mask=None, output_hidden_states=False)
        rh = clip(input_ids=idc, attention_mask=None, output_hidden_states=False)
    for key, gold in (("last_hidden_state", g["y_clip_hidden"]), ("pooler_output", g["y_clip_pooled"])):
        floor = (rh[key].float().cpu() - gold).abs().mean().item()
        mine = (r[key].float().cpu() - gold).abs().mean().item()
        print(f"tiny CLIP {key}: ours-vs-fp32 {mine:.4g}, HF-bf16-vs-fp32 {floor:.4g}")
        assert mine <= 1.25 * floor
    assert r.pooler_output.shape == (2, T.CLIP_TINY["hidden_size"])


def _device_init(m, seed, gain=1.0):
    """On-device seeded parameters for the full-size encoders (4.7 B values: too slow through the CPU generator).  gain < 1
    keeps a randomly initialised 24-block stack contractive, so that bf16 rounding noise is not amplified to O(1) by the
    depth (a trained network is stable; N(0, 1/fan_in) blocks are not) and the comparison stays discriminating."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    with torch.no_grad():
        for k, p in sorted(m.state_dict().items()):
            if not p.dtype.is_floating_point:
                continue
            if "norm" in k and k.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, device=DEV, generator=g))
            elif k.endswith("bias") and p.dim() == 1:
                p.copy_(0.05 * torch.randn(p.shape, device=DEV, generator=g))
            elif "embed" in k or "shared" in k or "relative_attention_bias" in k:
                p.copy_(0.5 * torch.randn(p.shape, device=DEV, generator=g))
            else:
                p.copy_(torch.randn(p.shape, device=DEV, generator=g) * (gain / math.sqrt(p.shape[-1])))


def _timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        out = fn()
    e.record()
    torch.cuda.synchronize()
    return out, s.elapsed_time(e) / n


@pytest.mark.parametrize("kind", ["clip-l", "t5-xxl"])
def test_full_size_encoders_against_hugging_face_on_this_gpu(kind):
    """clip-vit-large-patch14's text tower (12 x 768, 77 tokens) and the t5-v1_1-xxl encoder (24 x 4096, 512 tokens) with
    seeded weights: ours vs the Hugging Face module in bf16, against that module's own bf16-vs-fp32 distance; both timed."""
    from transformers import CLIPTextConfig, CLIPTextModel, T5Config, T5EncoderModel

    with torch.device(DEV):
        if kind == "clip-l":
            cfg = CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                                 max_position_embeddings=77, hidden_act="quick_gelu", eos_token_id=2, bos_token_id=0, pad_token_id=1)
            m, S, key = CLIPTextModel(cfg).to(BF16).eval(), 77, "pooler_output"
        else:
            cfg = T5Config(vocab_size=32128, d_model=4096, d_kv=64, d_ff=10240, num_layers=24, num_heads=64,
                           relative_attention_num_buckets=32, relative_attention_max_distance=128, dropout_rate=0.0,
                           layer_norm_epsilon=1e-6, feed_forward_proj="gated-gelu", is_encoder_decoder=False, use_cache=False,
                           tie_word_embeddings=False)
            m, S, key = T5EncoderModel(cfg).to(BF16).eval(), 512, "last_hidden_state"
    _device_init(m, 17, gain=1.0 if kind == "clip-l" else 0.5)
    ids = torch.randint(3, 30000, (1, S), device=DEV, generator=torch.Generator(device=DEV).manual_seed(18))
    if kind == "clip-l":
        ids[0, 0], ids[0, 30], ids[0, 31:] = 0, 49407, 1
    ours = CD.accelerate(m)
    with torch.inference_mode():
        y, ms_ours = _timed(lambda: ours(input_ids=ids, attention_mask=None, output_hidden_states=False)[key])
        y_hf, ms_hf = _timed(lambda: m(input_ids=ids, attention_mask=None, output_hidden_states=False)[key])
        y, y_hf = y.float(), y_hf.float()
        y32 = m.float()(input_ids=ids, attention_mask=None, output_hidden_states=False)[key]
    floor, mine = (y_hf - y32).abs(), (y - y32).abs()
    rep = {"encoder": kind, "tokens": S, "amax": y32.abs().max().item(), "hf_bf16_vs_fp32": {"mean": floor.mean().item(), "max": floor.max().item()},
           "ours_vs_fp32": {"mean": mine.mean().item(), "max": mine.max().item()},
           "ours_vs_hf_bf16": {"mean": (y - y_hf).abs().mean().item(), "max": (y - y_hf).abs().max().item()},
           "ms_hugging_face": ms_hf, "ms_ours": ms_ours}
    print(json.dumps(rep))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"text_parity_{kind}.json"), "w") as f:
        json.dump(rep, f, indent=1)
    assert torch.isfinite(y).all()
    assert mine.mean().item() <= 1.25 * floor.mean().item()
    assert mine.max().item() <= 2.0 * floor.max().item()


def test_graph_replay_equals_eager_launches_and_follows_weight_updates():
    """The CUDA-graph replay of an encoder returns the bits of its launch-by-launch run, for new token ids too, and is
    re-captured when a parameter of the wrapped module changes."""
    t5 = _hf("t5", T.T5_TINY, 3)
    ours = CD.accelerate(t5)
    g = torch.Generator(device=DEV).manual_seed(9)
    ids1 = torch.randint(0, 512, (2, 40), device=DEV, generator=g)
    ids2 = torch.randint(0, 512, (2, 40), device=DEV, generator=g)
    with torch.inference_mode():
        ours.use_graph = False
        e1, e2 = ours(input_ids=ids1).last_hidden_state.clone(), ours(input_ids=ids2).last_hidden_state.clone()
        ours.use_graph = True
        g1, g2, g1b = ours(input_ids=ids1).last_hidden_state, ours(input_ids=ids2).last_hidden_state, ours(input_ids=ids1).last_hidden_state
        assert torch.equal(g1, e1) and torch.equal(g2, e2) and torch.equal(g1b, e1)
        with torch.no_grad():
            t5.encoder.block[0].layer[1].DenseReluDense.wo.weight.mul_(0.5)
        changed = ours(input_ids=ids1).last_hidden_state
        assert not torch.equal(changed, e1)
        ours.use_graph = False
        assert torch.equal(ours(input_ids=ids1).last_hidden_state, changed)


def test_loud_failures_of_the_wrapper():
    from transformers import T5Config, T5EncoderModel

    m = T5EncoderModel(T5Config(**T.T5_TINY)).eval()  # fp32, CPU
    with pytest.raises(ValueError):
        CD.accelerate(m)
    with pytest.raises(ValueError):
        CD.accelerate(torch.nn.Linear(4, 4))
    t5 = _hf("t5", T.T5_TINY, 3)
    with pytest.raises(NotImplementedError):
        CD.accelerate(t5)(input_ids=torch.zeros(1, 8, dtype=torch.long, device=DEV), attention_mask=torch.ones(1, 8, device=DEV))
