#!/bin/bash
# One GPU-box visit: parity tests, bench line, launch list and ncu captures (outputs under gpurun_out/).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
tail -25 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
if [ "$1" != "noprof" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_x3.csv python tools/profile_step.py bf16x3 1 1 > gpurun_out/prof_step.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'krows|mnrows|hm_kgemm_kernel<64>|hm_kgemm_kernel<128>|mngemm' -c 14 \
    -f -o gpurun_out/prof_rows python tools/profile_step.py bf16x3 1 1 >> gpurun_out/prof_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'kgemm2' -s 3 -c 1 \
    -f -o gpurun_out/prof_k1 python tools/k1_only.py bf16x3 5 >> gpurun_out/prof_step.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'in_bwd_apply|in_bwd_reduce' -s 60 -c 8 \
    -f -o gpurun_out/prof_in python tools/profile_step.py bf16x3 1 1 >> gpurun_out/prof_step.log 2>&1
tail -5 gpurun_out/prof_step.log
fi
ls -la gpurun_out
