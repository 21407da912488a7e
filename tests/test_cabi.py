"""CPU: libflux_b200.so loads without a GPU, exports every symbol include/flux_b200.h declares, the ctypes
struct mirrors match the C layout, and argument validation fails loudly (no compute is launched)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "flux_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fluxb200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from flux_fp8_api_b200 import _cabi

    assert declared_symbols() == sorted(_cabi.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in flux_b200.h but not exported"
    assert lib.fluxb200_version() == 100
    assert lib.fluxb200_last_error() is not None


def test_struct_layout_matches_c(tmp_path):
    from flux_fp8_api_b200 import _cabi

    src = tmp_path / "layout.c"
    fields_g = [f[0] for f in _cabi.GemmArgs._fields_]
    fields_a = [f[0] for f in _cabi.AttentionArgs._fields_]
    fields_l = [f[0] for f in _cabi.LnArgs._fields_]
    fields_v = [f[0] for f in _cabi.GemvLayer._fields_]
    fields_c = [f[0] for f in _cabi.ConvArgs._fields_]
    body = ['#include <stdio.h>', '#include <stddef.h>', '#include "flux_b200.h"', "int main(void){"]
    body.append('printf("%zu\\n", sizeof(fluxb200_gemm_args));')
    body += [f'printf("%zu\\n", offsetof(fluxb200_gemm_args, {f}));' for f in fields_g]
    body.append('printf("%zu\\n", sizeof(fluxb200_attention_args));')
    body += [f'printf("%zu\\n", offsetof(fluxb200_attention_args, {f}));' for f in fields_a]
    body.append('printf("%zu\\n", sizeof(fluxb200_ln_args));')
    body += [f'printf("%zu\\n", offsetof(fluxb200_ln_args, {f}));' for f in fields_l]
    body.append('printf("%zu\\n", sizeof(fluxb200_gemv_layer));')
    body += [f'printf("%zu\\n", offsetof(fluxb200_gemv_layer, {f}));' for f in fields_v]
    body.append('printf("%zu\\n", sizeof(fluxb200_conv_args));')
    body += [f'printf("%zu\\n", offsetof(fluxb200_conv_args, {f}));' for f in fields_c]
    body.append("return 0;}")
    src.write_text("\n".join(body))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    nums = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert nums[0] == C.sizeof(_cabi.GemmArgs)
    assert nums[1:1 + len(fields_g)] == [getattr(_cabi.GemmArgs, f).offset for f in fields_g]
    rest = nums[1 + len(fields_g):]
    assert rest[0] == C.sizeof(_cabi.AttentionArgs)
    assert rest[1:1 + len(fields_a)] == [getattr(_cabi.AttentionArgs, f).offset for f in fields_a]
    rest = rest[1 + len(fields_a):]
    assert rest[0] == C.sizeof(_cabi.LnArgs)
    assert rest[1:1 + len(fields_l)] == [getattr(_cabi.LnArgs, f).offset for f in fields_l]
    rest = rest[1 + len(fields_l):]
    assert rest[0] == C.sizeof(_cabi.GemvLayer)
    assert rest[1:1 + len(fields_v)] == [getattr(_cabi.GemvLayer, f).offset for f in fields_v]
    rest = rest[1 + len(fields_v):]
    assert rest[0] == C.sizeof(_cabi.ConvArgs)
    assert rest[1:] == [getattr(_cabi.ConvArgs, f).offset for f in fields_c]


def test_invalid_arguments_fail_loudly(lib):
    from flux_fp8_api_b200 import _cabi

    assert lib.fluxb200_quantize(None, None, 16, None, 0, None) == _cabi.ERR_INVALID
    assert b"null" in lib.fluxb200_last_error()
    assert lib.fluxb200_f8_gemm(None, None) == _cabi.ERR_INVALID
    g = _cabi.GemmArgs()
    assert lib.fluxb200_f8_gemm(C.byref(g), None) == _cabi.ERR_INVALID
    a = _cabi.AttentionArgs()
    assert lib.fluxb200_attention(C.byref(a), None) == _cabi.ERR_INVALID
    assert lib.fluxb200_ln_mod_quant_grouped(None, 1, 1, 3072, 1e-6, None) == _cabi.ERR_INVALID
    ln = (_cabi.LnArgs * 2)()
    assert lib.fluxb200_ln_mod_quant_grouped(ln, 3, 1, 3072, 1e-6, None) == _cabi.ERR_INVALID      # at most two row sets
    assert lib.fluxb200_ln_mod_quant_grouped(ln, 2, 1, 3000, 1e-6, None) == _cabi.ERR_INVALID      # D % 256
    assert lib.fluxb200_lora_fuse(None, 0, None, None, None, 8, 8, 4, 1, 1.0, 0, None, None, None) == _cabi.ERR_INVALID
    assert b"null" in lib.fluxb200_last_error()
    assert lib.fluxb200_modulation_batched(None, None, 1, 1, None, None, 0, 1, 3072, 1, 0, None) == _cabi.ERR_INVALID
    with pytest.raises(ValueError):
        _cabi.check(_cabi.ERR_INVALID, "x")
    with pytest.raises(_cabi.FluxB200Error):
        _cabi.check(_cabi.ERR_CUDA, "x")


def test_cpu_tensors_are_rejected_not_computed():
    """There is no CPU fallback: product entry points refuse CPU tensors."""
    import torch

    from flux_fp8_api_b200 import _cabi, ops
    from flux_fp8_api_b200.f8linear import F8Linear

    x = torch.zeros(4, 16, dtype=torch.bfloat16)
    with pytest.raises(_cabi.FluxB200Error):
        ops.quantize(x, torch.tensor(1.0), torch.float8_e5m2)
    with pytest.raises(_cabi.FluxB200Error):
        F8Linear.from_linear(torch.nn.Linear(16, 16).to(torch.bfloat16))
    with pytest.raises(_cabi.FluxB200Error):
        ops.ln_mod_quant_pair([(torch.zeros(1, 8, 256, dtype=torch.bfloat16), x[:1, :0], x[:1, :0], torch.tensor(1.0))] * 2,
                              torch.float8_e5m2)
    with pytest.raises(_cabi.FluxB200Error):
        ops.lora_fuse(torch.zeros(8, 16, dtype=torch.float8_e4m3fn), torch.tensor(1.0), torch.zeros(8, 2), torch.zeros(2, 16), 1.0)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure; nothing under flux-fp8-api_b200/ may reach it."""
    pkg = os.path.join(ROOT, "flux-fp8-api_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "flux_oracle" not in text, f


def test_a_plain_c_program_links_and_calls_the_library(tmp_path, lib):
    """INTEGRATION.md option B: a C host includes include/flux_b200.h and links libflux_b200.so.  No GPU needed: the
    program checks the version, that invalid arguments come back as FLUXB200_ERR_INVALID with a message (never a
    crash), and that the measurement overrides validate their input."""
    from flux_fp8_api_b200 import _cabi

    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "flux_b200.h"
int main(void) {
  if (fluxb200_version() != FLUXB200_VERSION) return 1;
  fluxb200_gemm_args g;
  memset(&g, 0, sizeof g);
  int rc = fluxb200_f8_gemm(&g, NULL);
  if (rc != FLUXB200_ERR_INVALID || strlen(fluxb200_last_error()) == 0) return 2;
  if (fluxb200_quantize(NULL, NULL, 16, NULL, FLUXB200_E4M3, NULL) != FLUXB200_ERR_INVALID) return 3;
  if (fluxb200_gemm_force_tiling(7, 0) != FLUXB200_ERR_INVALID) return 4;
  if (fluxb200_gemm_probe_mode(99) != FLUXB200_ERR_INVALID) return 5;
  if (fluxb200_bf16_gemv(NULL, NULL, NULL, NULL, NULL, 0, NULL, 0, 1, 64, 64, 0, NULL) != FLUXB200_ERR_INVALID) return 6;
  printf("ok %d %s\n", fluxb200_version(), fluxb200_last_error());
  return 0;
}
''')
    exe = tmp_path / "host"
    libdir = os.path.dirname(_cabi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-l:libflux_b200.so", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([str(exe)]).decode()
    assert out.startswith("ok 100 ")
