#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 900 > gpurun_out/d_pytest_gpu.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|Error" gpurun_out/d_pytest_gpu.log | tail -12
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-torch-gpu --no-alt > gpurun_out/d_bench_streams.json 2> gpurun_out/d_bench_streams.err
echo "bench(streams) rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/d_bench_streams.json')); print('streams on ', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['ms_per_launch'], d['clocks'], d['cuda_graph'])"; tail -2 gpurun_out/d_bench_streams.err
HM_STREAMS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-torch-gpu --no-alt > gpurun_out/d_bench_nostreams.json 2> gpurun_out/d_bench_nostreams.err
echo "bench(no streams) rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/d_bench_nostreams.json')); print('streams off', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['ms_per_launch'], d['clocks'], d['cuda_graph'])"
HM_STREAMS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/d_launches_x3.csv python tools/profile_step.py bf16x3 1 1 > gpurun_out/d_prof_step.log 2>&1
tail -2 gpurun_out/d_prof_step.log
