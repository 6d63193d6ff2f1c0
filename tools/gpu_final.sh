#!/bin/bash
# Round-end evidence run (1 GPU): parity tests, smoke, bench line, launch lists (bf16x3 / bf16), ncu captures of K1 in
# both precisions and of a few other engines.  Outputs under gpurun_out/ (kept below the 64 MiB merge limit).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
for prec in bf16x3 bf16; do
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$prec.csv python tools/profile_step.py $prec 1 1 > gpurun_out/prof_step_$prec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'kgemm2' -s 3 -c 1 \
    -f -o gpurun_out/prof_k1_$prec python tools/k1_only.py $prec 5 >> gpurun_out/prof_step_$prec.log 2>&1
done
ncu --set full --clock-control none --profile-from-start off -k regex:'krows2|mngemm2|mnrows' -c 6 \
    -f -o gpurun_out/prof_engines python tools/profile_step.py bf16x3 1 1 >> gpurun_out/prof_step_bf16x3.log 2>&1
ls -la gpurun_out | head -30
