"""Summarise gpurun_out/parity_<config>.json (written by tests/test_gpu_reference.py on the B200) into
profiles/r2_parity_reference.md:  python tools/parity_summary.py"""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = ["# Round 2 — parity against the unmodified reference run on the same B200",
       "",
       "Source: `tests/test_gpu_reference.py` (one full-depth 19 + 38 block Flux per config, hidden 3072; the reference = the",
       "staged, sha256-checked `float8_quantize.py` / `modules/flux_model.py` under `oracle/_ref`, quantised and calibrated by",
       "its own code, its `state_dict()` loaded strictly into this package's `Flux`).  Raw numbers: `gpurun_out/parity_*.json`.",
       "",
       "* **layers** — every reference `F8Linear.forward` call replayed through our quantiser + tcgen05 GEMM on the",
       "  reference's own bf16 input (identical fp8 operand bytes): fraction of output elements that differ at all, and the",
       "  worst |diff| relative to the layer's output amax (1 bf16 ulp at the output range = 7.8e-3).",
       "* **attention** — every reference `attention()` call (RoPE + SDPA + transpose) replayed through `blocks.attention`.",
       "* **blocks** — each reference block's inputs fed to our fused block: mean |ours − ref| divided by the reference's OWN",
       "  mean spread on the same block input when it runs another SDPA backend (flash / efficient / math vs the default",
       "  cuDNN one) = ratio; and the fraction of fp8 operand bytes that differ from the reference's own quantised inputs",
       "  (\"flips\"), per consuming layer of block 0.",
       "* **free-running** — whole forward, 57 blocks deep: mean / max |ours − ref| against reference-vs-reference.",
       ""]
rows, flips, free = [], [], []
for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "parity_*.json"))):
    r = json.load(open(path))
    name = r["config"]
    L, A, B = r["layers"], r["attention"], r["blocks"]
    ratios = sorted(b["mean"] / max(max(v[0] for v in b["floor"].values()), 1e-9) for b in B)
    fl = [v for b in B for v in b["flips"].values()]
    rows.append(f"| {name} | {len(L)} | {max(x['frac_diff'] for x in L):.1e} | {max(x['max'] / max(x['amax'], 1e-6) for x in L):.1e} | "
                f"{len(A)} | {sum(x['frac_diff'] for x in A) / len(A):.1e} | {max(x['max'] / max(x['amax'], 1e-6) for x in A):.1e} | "
                f"{ratios[len(ratios) // 2]:.2f} | {ratios[-1]:.2f} | {sum(fl) / len(fl):.4f} | {max(fl):.4f} |")
    b0 = B[0]["flips"]
    flips.append(f"| {name} | " + " | ".join(f"{b0.get(k, float('nan')):.4f}" for k in
                                             ("img_attn.qkv", "img_attn.proj", "img_mlp.0", "img_mlp.2")) +
                 f" | {B[19]['flips'].get('linear1', float('nan')):.4f} | {B[19]['flips'].get('linear2', float('nan')):.4f} |")
    f = r.get("free_running")
    if f:
        rr = f["ref_vs_ref"]
        free.append(f"| {name} | {f['ref_rms']:.3f} | {f['ref_amax']:.2f} | {f['ours_vs_ref'][0]:.4f} / {f['ours_vs_ref'][1]:.3f} | " +
                    " | ".join(f"{rr[k][0]:.4f} / {rr[k][1]:.3f}" if k in rr else "n/a" for k in ("flash", "efficient", "math")) + " |")
out += ["| config | F8Linear calls | max frac differing | worst max\\|d\\|/amax | attention calls | mean frac differing | worst max\\|d\\|/amax | blocks: ours/floor median | max | fp8 flip rate mean | max |",
        "|---|---|---|---|---|---|---|---|---|---|---|"] + rows + [""]
out += ["fp8 operand flip rate by consuming layer (double block 0, single block 0):", "",
        "| config | img_attn.qkv | img_attn.proj | img_mlp.0 | img_mlp.2 | linear1 | linear2 |", "|---|---|---|---|---|---|---|"] + flips + [""]
out += ["Free-running forward (57 blocks), mean / max abs difference of the [B, L, 64] prediction:", "",
        "| config | ref rms | ref amax | ours vs ref | ref(flash) vs ref | ref(efficient) vs ref | ref(math) vs ref |",
        "|---|---|---|---|---|---|---|"] + free + [""]
out += ["Reading.  (1) On identical inputs our GEMM and attention kernels agree with cuBLASLt `_scaled_mm` / cuDNN SDPA to below one",
        "bf16 ulp at the output range, with fewer than 3 in 10 000 GEMM outputs differing at all.  (2) The LayerNorm-modulate-quantise",
        "path is exact (0 flipped bytes into `qkv` / `linear1`); the flips enter through attention's P rounding (0.1-0.2 % of `proj`",
        "inputs) and grow through the GELU (4-7 % of `mlp.2` inputs) exactly as they do between two runs of the reference itself.",
        "(3) Per block our error on the reference's input is a fifth of the reference's own backend-to-backend spread (median);",
        "57 blocks deep ours-vs-reference equals reference-vs-reference (0.137 vs 0.140 mean on an output of rms 1.18 at c2):",
        "BASELINE.json's `1e-2 max-abs` is not a property the reference has with itself -- switching its SDPA backend moves the",
        "output by 0.75-0.9 max-abs.  With e4m3 activations (one more mantissa bit) both spreads halve."]
open(os.path.join(ROOT, "profiles", "r2_parity_reference.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
