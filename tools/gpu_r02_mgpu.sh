#!/bin/bash
# N-GPU visit: product-path equivalence check (2 ranks x B == 1 rank x 2B, identical replicas) and the bench at N GPUs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tests/multigpu_check.py --out gpurun_out/mgpu_check.json > gpurun_out/mgpu_check.log 2>&1
echo "multigpu_check rc=$?"; grep -E "MULTIGPU_CHECK|Error|assert" gpurun_out/mgpu_check.log | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/mgpu_bench_n$N.json 2> gpurun_out/mgpu_bench_n$N.err
echo "bench N=$N rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/mgpu_bench_n$N.json')); print('N=$N', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d.get('replicas_identical'), d['cuda_graph'], d['clocks'])"
tail -3 gpurun_out/mgpu_bench_n$N.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-torch-gpu --no-alt > gpurun_out/mgpu_bench_n1.json 2> gpurun_out/mgpu_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/mgpu_bench_n1.json')); print('N=1', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])"
