"""Stage the UNMODIFIED reference modules of the hot path under oracle/_ref/ so they travel to the GPU box.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): only tests/, __graft_entry__.smoke() and bench.py's
reference / cpu_baseline legs may import what this stages.  The product path never does.

The reference (aredden/flux-fp8-api) is pure Python over PyTorch: there is nothing to compile, and it has no
setup.py / pyproject, so `pip install --target baseline/_ref /root/reference` has nothing to install.  The
"reference build" for this repository is therefore a byte-for-byte staging of the files SURVEY.md
section 8(a) / 8(f) cite,

    float8_quantize.py        F8Linear, recursive_swap_linears, quantize_flow_transformer_and_dispatch_float8
    modules/flux_model.py     Flux, DoubleStreamBlock, SingleStreamBlock, Modulation, attention, rope, QKNorm ...
    lora_loading.py           apply_lora_to_model / remove_lora_from_module (config c4 merges a LoRA with it)
    modules/autoencoder.py    AutoEncoder / Decoder (the VAE decode that follows the denoise loop)

into the git-ignored directory oracle/_ref/ (listed in .gitignore, NOT in .gpurunignore: it ships with the
gpurun snapshot exactly like the built libflux_b200.so), plus MANIFEST.json with the sha256 of every file so a
test can prove the staged copy is the unmodified reference.  No reference source enters the git history.

    python oracle/fetch_ref.py            # needs /root/reference (or $FLUX_REFERENCE); run by build()

On the GPU box /root/reference does not exist; tests and bench.py import the staged copy through
oracle/ref_loader.py and skip / report "unavailable" when it is absent.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ("float8_quantize.py", "modules/flux_model.py", "lora_loading.py", "modules/autoencoder.py")


def sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(reference_root: str = None) -> dict:
    root = reference_root or os.environ.get("FLUX_REFERENCE", "/root/reference")
    if not os.path.isdir(root):
        raise FileNotFoundError(f"reference tree {root} not found (expected in the authoring container only)")
    manifest = {"source": root, "files": {}}
    for rel in FILES:
        src, dst = os.path.join(root, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest["files"][rel] = {"sha256": sha256(dst), "bytes": os.path.getsize(dst)}
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return manifest


if __name__ == "__main__":
    m = stage(sys.argv[1] if len(sys.argv) > 1 else None)
    for rel, info in m["files"].items():
        print(f"staged {rel}  {info['bytes']} B  sha256 {info['sha256'][:16]}")
