#!/bin/bash
# One GPU session.  usage: tools/gpu_round.sh [tests] [bench] [configs] [shapes] [vae] [ncu]   (default: tests bench)
# Everything is written under gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
what="${*:-tests bench}"
has() { [[ " $what " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv,noheader
if has tests; then
  python -m pytest tests/ -q -m gpu --durations=15 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
fi
if has bench; then
  python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
fi
if has configs; then
  for c in c3 c4 c5; do
    python bench.py --config $c --steps 12 --warmup 3 --gpu-reference eager --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c rc=$?"; cat gpurun_out/bench_$c.json; tail -2 gpurun_out/bench_$c.err
  done
fi
if has shapes; then
  python tools/step_shapes.py > gpurun_out/step_shapes.txt 2>&1; tail -14 gpurun_out/step_shapes.txt
  FLUXB200_GEMM_WIDE=0 python tools/step_shapes.py > gpurun_out/step_shapes_narrow.txt 2>&1; tail -14 gpurun_out/step_shapes_narrow.txt
fi
if has vae; then
  for r in 1024 768 1536; do python tools/vae_profile.py $r 1 > gpurun_out/vae_profile_$r.txt 2>&1; head -1 gpurun_out/vae_profile_$r.txt; tail -8 gpurun_out/vae_profile_$r.txt; done
fi
if has quad; then
  FLUXB200_GEMM_MC=2 python tools/step_shapes.py > gpurun_out/step_shapes_quad.txt 2>&1; tail -14 gpurun_out/step_shapes_quad.txt
fi
if has ablate; then
  python tools/ablate_step.py c2 > gpurun_out/ablate_c2.txt 2>&1; cat gpurun_out/ablate_c2.txt
fi
if has sdpa; then
  python tools/sdpa_compare.py > gpurun_out/sdpa_compare.txt 2>&1; cat gpurun_out/sdpa_compare.txt
fi
if has ncu; then
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
  ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/traffic.csv python tools/profile_step.py > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
  ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:f8_gemm_kernel -c 4 -o gpurun_out/prof_gemm_double -f python tools/profile_step.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm double rc=$?"
  ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:f8_gemm_kernel --launch-skip 76 -c 2 -o gpurun_out/prof_gemm_single -f python tools/profile_step.py >> gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm single rc=$?"
  ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:attention_kernel -c 1 -o gpurun_out/prof_attn -f python tools/profile_step.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
  ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:ln_mod_quant -c 2 -o gpurun_out/prof_ln -f python tools/profile_step.py > gpurun_out/ncu_ln.log 2>&1; echo "ncu ln rc=$?"
fi
ls -la gpurun_out | tail -30
