"""The reference-side binding: redirect aredden/flux-fp8-api's hot-path names to the B200 classes.

The reference has no FFI / plugin registry; its boundary for this path is the module surface of
`float8_quantize.py` and `modules/flux_model.py`.  A maintainer drops the B200 path in with

    import float8_quantize, lora_loading
    import modules.flux_model as flux_model
    from flux_fp8_api_b200 import reference_binding
    reference_binding.bind(float8_quantize, flux_model, lora_loading)      # before the model is constructed
    # ... bind(..., autoencoder=modules.autoencoder) also redirects the VAE's decode half (before util.load_autoencoder)

after which `util.load_flow_model`, `quantize_flow_transformer_and_dispatch_float8`, `FluxPipeline.generate` and
`Flux.load_lora` run unchanged on the sm_100a kernels.  tests/test_reference_binding.py executes exactly this against
the staged reference (oracle/_ref): the reference's OWN `Flux` container is constructed over the B200 blocks, loads a
reference-minted prequantised state dict, and (on the GPU) produces the same output as this package's `Flux`.

Nothing here imports the reference: the caller passes the reference's module objects in.
"""
from __future__ import annotations

from types import ModuleType
from typing import Optional

from . import autoencoder as _ae
from . import blocks as _blocks
from . import f8linear as _f8
from . import lora as _lora
from . import model as _model

#: modules/flux_model.py names replaced by blocks.py (block layer)
BLOCK_NAMES = ("attention", "rope", "apply_rope", "EmbedND", "RMSNorm", "QKNorm", "SelfAttention", "Modulation",
               "ModulationOut", "DoubleStreamBlock", "SingleStreamBlock")
#: float8_quantize.py names replaced by f8linear.py (operator layer, rows a1-a3)
F8_NAMES = ("F8Linear", "recursive_swap_linears", "quantize_flow_transformer_and_dispatch_float8",
            "swap_to_cublaslinear")
#: modules/flux_model.py container names replaced by model.py when replace_container=True (row a11)
CONTAINER_NAMES = ("Flux", "MLPEmbedder", "LastLayer", "timestep_embedding")


#: modules/autoencoder.py names replaced by autoencoder.py
AE_NAMES = ("AttnBlock", "ResnetBlock", "Upsample", "Downsample", "Encoder", "Decoder", "DiagonalGaussian", "AutoEncoder")


def bind(float8_quantize: ModuleType, flux_model: ModuleType, lora_loading: Optional[ModuleType] = None,
         replace_container: bool = True, autoencoder: Optional[ModuleType] = None) -> dict:
    """Point the reference's module attributes at the B200 implementations.  Returns {qualified name: original object}
    so `unbind` can restore them.

    replace_container=False keeps the reference's own `Flux.forward` (modules/flux_model.py:672-716) as the caller
    of the B200 blocks: everything still runs on our kernels, without the step-invariant cache, batched modulation
    and cos/sin hand-down that `model.Flux` adds."""
    saved = {}

    def put(mod: ModuleType, name: str, obj) -> None:
        saved[(mod.__name__, name)] = (mod, getattr(mod, name, None))
        setattr(mod, name, obj)

    for name in F8_NAMES:
        put(float8_quantize, name, getattr(_f8, name))
    # float8_quantize.py:10 imported Modulation by value for its isinstance() check in recursive_swap_linears
    put(float8_quantize, "Modulation", _blocks.Modulation)
    for name in BLOCK_NAMES:
        put(flux_model, name, getattr(_blocks, name))
    if replace_container:
        for name in CONTAINER_NAMES:
            put(flux_model, name, getattr(_model, name))
    if autoencoder is not None:
        # util.load_autoencoder (util.py:280-287) builds `AutoEncoder(config.ae_params)` by this name and loads the
        # checkpoint with strict=False; FluxPipeline.vae_decode (flux_pipeline.py:422-438) then calls `.decode`
        for name in AE_NAMES:
            put(autoencoder, name, getattr(_ae, name))
    if lora_loading is not None:
        # lora_loading.py:13-14 imported F8Linear / Flux by value; its fuse / unfuse entry points are replaced by the
        # on-device versions (same names, same argument order; the key-layout conversion helpers stay the reference's)
        put(lora_loading, "F8Linear", _f8.F8Linear)
        if replace_container:
            put(lora_loading, "Flux", _model.Flux)
        resolve = getattr(lora_loading, "resolve_lora_state_dict", None)
        get_weights = getattr(lora_loading, "get_lora_weights", None)

        def _bfl(model, lora_path):
            """A file path goes through the reference's own loader + key conversion (lora_loading.py:608-612, 580-605);
            dicts / LoraWeights are 'already loaded' BFL-layout state, as in the reference (:638-640)."""
            if isinstance(lora_path, str) and get_weights is not None and resolve is not None:
                weights, _ = get_weights(lora_path)
                _, weights = resolve(weights, model.params.guidance_embed)
                return weights
            return lora_path

        def apply_lora_to_model(model, lora_path, lora_scale: float = 1.0, return_lora_resolved: bool = False):
            return _lora.apply_lora_to_model(model, _bfl(model, lora_path), lora_scale, return_lora_resolved)

        def remove_lora_from_module(model, lora_path, lora_scale: float = 1.0):
            return _lora.remove_lora_from_module(model, _bfl(model, lora_path), lora_scale)

        put(lora_loading, "apply_lora_to_model", apply_lora_to_model)
        put(lora_loading, "remove_lora_from_module", remove_lora_from_module)
        put(lora_loading, "LoraWeights", _lora.LoraWeights)
    return saved


def unbind(saved: dict) -> None:
    for (_, name), (mod, obj) in saved.items():
        if obj is None:
            try:
                delattr(mod, name)
            except AttributeError:
                pass
        else:
            setattr(mod, name, obj)
