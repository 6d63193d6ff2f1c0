#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/h_pytest_gpu.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|Error" gpurun_out/h_pytest_gpu.log | tail -8
for v in "HM_FUSED3_PAIR=1" "HM_FUSED3_PAIR=0"; do
  env $v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-torch-gpu --no-alt > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
  python -c "
import json; d=json.load(open('gpurun_out/h_bench.json')); print('$v', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['roofline']['ms_per_launch'],4), round(d['roofline']['frac'],4))"
done
HM_STREAMS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/h_launches_x3.csv python tools/profile_step.py bf16x3 1 1 > gpurun_out/h_prof_step.log 2>&1
tail -2 gpurun_out/h_prof_step.log
