#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/g_pytest_gpu.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|Error" gpurun_out/g_pytest_gpu.log | tail -8
for v in "HM_FUSED3=1" "HM_FUSED3=0"; do
  env $v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-torch-gpu --no-alt > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
  python -c "
import json; d=json.load(open('gpurun_out/g_bench.json')); print('$v', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'])"
done
HM_STREAMS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/g_launches_x3.csv python tools/profile_step.py bf16x3 1 1 > gpurun_out/g_prof_step.log 2>&1
tail -2 gpurun_out/g_prof_step.log
