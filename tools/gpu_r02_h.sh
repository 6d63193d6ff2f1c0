#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_box2mask_gpu.py tests/test_model_gpu.py -q --timeout 600 > gpurun_out/i_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|Error" gpurun_out/i_pytest.log | tail -6
timeout 600 python bench.py --config 5 --steps 20 --warmup 5 > gpurun_out/i_bench_cfg5.json 2> gpurun_out/i_bench_cfg5.err
echo "cfg5 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/i_bench_cfg5.json')); print('cfg5', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['cuda_graph'], d['gpu_launches'])"; tail -2 gpurun_out/i_bench_cfg5.err
python tools/mma_issue_bench.py > gpurun_out/i_mma_issue.log 2>&1; cat gpurun_out/i_mma_issue.log | tail -24
