"""Import the staged, unmodified reference modules (oracle/_ref/, see oracle/fetch_ref.py) and drive them.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__ and bench.py's reference legs; never by the product
package (tests/test_cabi.py::test_product_does_not_import_the_oracle enforces it).

`load()` returns a namespace with the reference's own modules

    ref.f8   = float8_quantize        (F8Linear, recursive_swap_linears, quantize_flow_transformer_and_dispatch_float8)
    ref.fm   = modules.flux_model     (Flux, DoubleStreamBlock, SingleStreamBlock, Modulation, attention, ...)
    ref.lora = lora_loading           (apply_lora_to_model, remove_lora_from_module)   [None if its imports fail]

or raises ReferenceUnavailable.  Nothing here restates reference arithmetic: the helpers below only construct the
reference's classes and feed them tensors (what flux_pipeline.py / util.py do around them, which cannot be imported
here because of pydash / accelerate / quanto).
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import sys
import types
from typing import Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


class ReferenceUnavailable(RuntimeError):
    pass


_cached = None


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "MANIFEST.json"))


def verify_manifest() -> dict:
    """sha256 of every staged file against MANIFEST.json (and against /root/reference when that is present)."""
    with open(os.path.join(REF_DIR, "MANIFEST.json")) as f:
        manifest = json.load(f)
    for rel, info in manifest["files"].items():
        with open(os.path.join(REF_DIR, rel), "rb") as f:
            digest = hashlib.sha256(f.read()).hexdigest()
        if digest != info["sha256"]:
            raise ReferenceUnavailable(f"oracle/_ref/{rel} differs from its manifest: not the unmodified reference")
        src = os.path.join(manifest.get("source", ""), rel)
        if os.path.exists(src):
            with open(src, "rb") as f:
                if hashlib.sha256(f.read()).hexdigest() != digest:
                    raise ReferenceUnavailable(f"oracle/_ref/{rel} differs from {src}")
    return manifest


def load():
    """Import the staged reference (once).  The reference logs through loguru; a stub keeps it importable."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise ReferenceUnavailable("oracle/_ref is not staged: run `python oracle/fetch_ref.py` where /root/reference "
                                   "exists (build() does)")
    verify_manifest()
    try:
        import loguru  # noqa: F401
    except ImportError:  # pragma: no cover
        stub = types.ModuleType("loguru")
        stub.logger = types.SimpleNamespace(info=print, warning=print, error=print, debug=print, success=print)
        sys.modules["loguru"] = stub
    for name in ("float8_quantize", "modules", "modules.flux_model", "lora_loading"):
        mod = sys.modules.get(name)
        if mod is not None and not os.path.abspath(getattr(mod, "__file__", "") or "").startswith(REF_DIR):
            if name != "modules" or getattr(mod, "__file__", None):
                raise ReferenceUnavailable(f"a different module named {name!r} is already imported")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    ns = types.SimpleNamespace()
    ns.fm = importlib.import_module("modules.flux_model")
    ns.f8 = importlib.import_module("float8_quantize")
    try:
        ns.lora = importlib.import_module("lora_loading")
    except Exception as ex:  # noqa: BLE001  (optional third-party imports of that file)
        ns.lora, ns.lora_error = None, repr(ex)
    try:
        ns.ae = importlib.import_module("modules.autoencoder")
    except Exception as ex:  # noqa: BLE001  (staged by newer fetch_ref only)
        ns.ae, ns.ae_error = None, repr(ex)
    ns.dir = REF_DIR
    _cached = ns
    return ns


def model_spec(ref, params: dict, prequantized_flow: bool = False, quantize_modulation: bool = True,
               quantize_flow_embedder_layers: bool = False):
    """What reference Flux.__init__ reads from util.ModelSpec (modules/flux_model.py:510-520): `.params` (the
    reference's own pydantic FluxParams) and three flags."""
    return types.SimpleNamespace(params=ref.fm.FluxParams(**params), prequantized_flow=prequantized_flow,
                                 quantize_modulation=quantize_modulation,
                                 quantize_flow_embedder_layers=quantize_flow_embedder_layers)


@torch.inference_mode()
def build_reference_flux(ref, params: dict, bf16_state: dict, device, quantize: bool = True,
                         quantize_modulation: bool = True, quantize_flow_embedder_layers: bool = False,
                         input_float8_dtype=torch.float8_e5m2):
    """Reference Flux (bf16) carrying `bf16_state`, then the reference's own quantisation flow
    (float8_quantize.py:395-496), exactly as flux_pipeline.load_pipeline_from_config_path does (:706-716)."""
    spec = model_spec(ref, params, False, quantize_modulation, quantize_flow_embedder_layers)
    with torch.device(device):
        net = ref.fm.Flux(spec, dtype=torch.bfloat16).to(torch.bfloat16)
    missing, unexpected = net.load_state_dict(bf16_state, strict=True)
    assert not missing and not unexpected
    net.eval()
    if quantize:
        ref.f8.quantize_flow_transformer_and_dispatch_float8(
            net, torch.device(device), offload_flow=False, swap_linears_with_cublaslinear=False,
            flow_dtype=torch.bfloat16, input_float8_dtype=input_float8_dtype,
            quantize_modulation=quantize_modulation, quantize_flow_embedder_layers=quantize_flow_embedder_layers)
    return net


def all_frozen(ref, net) -> bool:
    return all(m.input_scale_initialized for m in net.modules() if isinstance(m, ref.f8.F8Linear))


def cpu_info() -> dict:
    name = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    name = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return {"model": name, "logical_cpus": os.cpu_count()}
