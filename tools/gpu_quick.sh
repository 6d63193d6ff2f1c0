#!/bin/bash
# Shorter GPU-box visit: parity tests, bench line, launch list, optional ncu capture of kernels matching $1.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_x3.csv python tools/profile_step.py bf16x3 1 1 > gpurun_out/prof_step.log 2>&1
if [ -n "$1" ]; then
ncu --set full --clock-control none --profile-from-start off -k regex:"$1" -s ${2:-0} -c ${3:-8} \
    -f -o gpurun_out/prof_sel python tools/profile_step.py bf16x3 1 1 >> gpurun_out/prof_step.log 2>&1
fi
tail -3 gpurun_out/prof_step.log
