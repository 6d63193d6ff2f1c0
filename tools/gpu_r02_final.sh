#!/bin/bash
# Round-2 evidence run (1 GPU): full GPU suite, smoke, the default bench line (CPU + stock-PyTorch legs included), the
# config #4 / #5 side lines, launch lists (bf16x3 / bf16) and ncu --set full captures of K1 (both precisions) and of the
# narrow engines.  Outputs under gpurun_out/ (kept below the 64 MiB merge limit).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 900 > gpurun_out/f_pytest_gpu.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|^FAILED" gpurun_out/f_pytest_gpu.log | tail -5
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; tail -1 gpurun_out/f_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/f_bench.json; tail -2 gpurun_out/f_bench.err
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/f_bench_cfg4.json 2> gpurun_out/f_bench_cfg4.err
echo "cfg4 rc=$?"; cut -c1-300 gpurun_out/f_bench_cfg4.json
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/f_bench_cfg5.json 2> gpurun_out/f_bench_cfg5.err
echo "cfg5 rc=$?"; cut -c1-300 gpurun_out/f_bench_cfg5.json; tail -2 gpurun_out/f_bench_cfg5.err
for prec in bf16x3 bf16; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/f_launches_$prec.csv python tools/profile_step.py $prec 1 1 > gpurun_out/f_prof_step_$prec.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kgemm2' -s 3 -c 1 \
    -f -o gpurun_out/f_prof_k1_$prec python tools/k1_only.py $prec 5 >> gpurun_out/f_prof_step_$prec.log 2>&1
done
HM_STREAMS=0 timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'krows2|mnrows|hm_kgemm_kernel|hm_mngemm_kernel' -c 14 \
    -f -o gpurun_out/f_prof_engines python tools/profile_step.py bf16x3 1 1 >> gpurun_out/f_prof_step_bf16x3.log 2>&1
ls -la gpurun_out | grep " f_"
