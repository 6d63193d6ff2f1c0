#!/bin/bash
# Round-2 visit A: the full GPU test suite (new full-size / audit / trajectory tests), default bench (with CPU + cuDNN legs),
# config #4 bench line, launch list of one step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt; nproc >> gpurun_out/a_gpu.txt; free -g >> gpurun_out/a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 900 -x > gpurun_out/a_pytest_gpu.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|error|config #|K1 @|backward audit|summary|Error" gpurun_out/a_pytest_gpu.log | tail -30
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?"; cut -c1-3000 gpurun_out/a_bench.json; tail -3 gpurun_out/a_bench.err
timeout 600 python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/a_bench_cfg4.json 2> gpurun_out/a_bench_cfg4.err
echo "bench4 rc=$?"; cut -c1-1500 gpurun_out/a_bench_cfg4.json; tail -3 gpurun_out/a_bench_cfg4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/a_launches_x3.csv python tools/profile_step.py bf16x3 1 1 > gpurun_out/a_prof_step.log 2>&1
tail -2 gpurun_out/a_prof_step.log
