#!/bin/bash
# Builds tools/ab/libflux_probe.so: the library with the attention phase timers compiled in (-DFLUXB200_ATTN_PROBE).
# Use it through FLUXB200_LIB=$PWD/tools/ab/libflux_probe.so python tools/attn_probe.py <variant ...>
set -e
cd "$(dirname "$0")/../flux-fp8-api_b200/csrc"
make -j8 > /dev/null
mkdir -p ../../tools/ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr \
  -cudart static -DFLUXB200_ATTN_PROBE -DFLUXB200_ATTN_EXPERIMENTS $EXTRA -c attention.cu -o /tmp/attention_probe.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../../tools/ab/libflux_probe.so \
  host_util.o elementwise.o f8_gemm.o /tmp/attention_probe.o lora.o vae.o text_encoder.o
