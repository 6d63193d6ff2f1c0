#!/bin/bash
mkdir -p gpurun_out
python tools/bench_k1_variants.py > gpurun_out/c_k1_variants.log 2>&1; cat gpurun_out/c_k1_variants.log | tail -12
timeout 900 python -m pytest tests/test_streamk_gpu.py tests/test_model_gpu.py tests/test_sn_gpu.py tests/test_backward_audit_gpu.py -q -s --timeout 600 > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|stream-K vs|backward audit|^FAILED" gpurun_out/c_pytest.log | tail -20
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-torch-gpu --no-alt > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c_bench.json')); print(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['clocks'])"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'krows2|hm_kgemm_kernel<64>|hm_kgemm_kernel<32>|mnrows|hm_mngemm_kernel<1>' -c 16 \
    -f -o gpurun_out/c_prof_narrow python tools/profile_step.py bf16x3 1 1 > gpurun_out/c_prof.log 2>&1
tail -2 gpurun_out/c_prof.log; ls -la gpurun_out/c_prof_narrow.ncu-rep
