#!/bin/bash
# One GPU session: parity tests, smoke, bench (both arms), per-shape step table, ncu launch list + DRAM traffic +
# full captures of the top kernels.   usage: tools/gpu_round.sh [noncu]
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 28 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
python tests/step_shapes.py > gpurun_out/step_shapes.txt 2>&1; tail -12 gpurun_out/step_shapes.txt
if [ "$1" != "noncu" ]; then
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tests/profile_step.py > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/traffic.csv python tests/profile_step.py > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:f8_gemm_kernel -c 8 -o gpurun_out/prof_gemm -f python tests/profile_step.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:attention_kernel -c 2 -o gpurun_out/prof_attn -f python tests/profile_step.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
fi
ls -la gpurun_out | tail -20
