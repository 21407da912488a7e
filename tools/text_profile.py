"""Per-launch breakdown of the t5-v1_1-xxl encoder (512 tokens) and the CLIP-L text tower (77 tokens) through our kernels,
next to the Hugging Face modules on the same GPU (seeded weights):  python tools/text_profile.py"""
import math
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transformers import CLIPTextConfig, CLIPTextModel, T5Config, T5EncoderModel  # noqa: E402

from flux_fp8_api_b200 import conditioner as CD, ops  # noqa: E402

DEV, BF16 = "cuda", torch.bfloat16


def init(m, gain):
    g = torch.Generator(device=DEV).manual_seed(17)
    with torch.no_grad():
        for k, p in sorted(m.state_dict().items()):
            if p.dtype.is_floating_point:
                if "norm" in k and k.endswith("weight"):
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, device=DEV, generator=g))
                else:
                    p.copy_(torch.randn(p.shape, device=DEV, generator=g) * (gain / math.sqrt(p.shape[-1]) if p.dim() > 1 else 0.05))


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


with torch.device(DEV):
    t5 = T5EncoderModel(T5Config(vocab_size=32128, d_model=4096, d_kv=64, d_ff=10240, num_layers=24, num_heads=64,
                                 relative_attention_num_buckets=32, relative_attention_max_distance=128, dropout_rate=0.0,
                                 layer_norm_epsilon=1e-6, feed_forward_proj="gated-gelu", is_encoder_decoder=False, use_cache=False,
                                 tie_word_embeddings=False)).to(BF16).eval()
    clip = CLIPTextModel(CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                                        num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu", eos_token_id=2,
                                        bos_token_id=0, pad_token_id=1)).to(BF16).eval()
init(t5, 0.5), init(clip, 1.0)
for name, m, S in (("t5-v1_1-xxl encoder", t5, 512), ("clip-vit-large-patch14 text", clip, 77)):
    ids = torch.randint(3, 30000, (1, S), device=DEV)
    ours = CD.accelerate(m)
    with torch.inference_mode():
        ms = timed(lambda: ours(input_ids=ids, attention_mask=None))
        ms_hf = timed(lambda: m(input_ids=ids, attention_mask=None))
        ours.use_graph = False
        ms_eager = timed(lambda: ours(input_ids=ids, attention_mask=None))
        ours.use_graph = True
        print(f"{name}: ours {ms:.2f} ms (CUDA-graph replay; {ms_eager:.2f} ms launch by launch), Hugging Face (bf16) {ms_hf:.2f} ms")
        ops.KERNEL_TIMELINE = []
        ours.use_graph = False  # launch by launch, so that every launch can be timed
        ours(input_ids=ids, attention_mask=None)
        torch.cuda.synchronize()
        tl, ops.KERNEL_TIMELINE = ops.KERNEL_TIMELINE, None
    agg = OrderedDict()
    for kind, work, s, e, detail in tl:
        a = agg.setdefault((kind, detail), [0.0, 0.0, 0])
        a[0] += work; a[1] += s.elapsed_time(e); a[2] += 1
    tot = sum(a[1] for a in agg.values())
    for (kind, detail), (work, t, n) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"   {kind:14s} {detail:34s} {n:4d} x {t / n * 1e3:8.1f} us = {t:7.3f} ms {t / tot * 100:5.1f}%  {work / (t * 1e-3) / 1e12:7.1f} TFLOP/s")
    print(f"   sum of timed launches {tot:.2f} ms")
