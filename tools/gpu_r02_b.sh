#!/bin/bash
# quick visit: full GPU suite (no -x), short bench (no baseline legs), launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 900 > gpurun_out/b_pytest_gpu.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|error|config #|K1 @|backward audit|summary|Error|^FAILED|^step|^ +[0-9]+ " gpurun_out/b_pytest_gpu.log | tail -60
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-torch-gpu $BENCH_FLAGS > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
echo "bench rc=$?"; cut -c1-1800 gpurun_out/b_bench.json; tail -3 gpurun_out/b_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/b_launches_x3.csv python tools/profile_step.py bf16x3 1 1 > gpurun_out/b_prof_step.log 2>&1
tail -2 gpurun_out/b_prof_step.log
