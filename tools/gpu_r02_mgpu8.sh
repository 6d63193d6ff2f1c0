#!/bin/bash
# 8-GPU confirmation of the allreduce schedule: buckets + 8 reserved SMs (default) vs the single serial allreduce, plus N=1
N=${1:-8}
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/mgpu8_${name}_n$N.json 2> gpurun_out/mgpu8_${name}_n$N.err
  echo "$name rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/mgpu8_${name}_n$N.json')); print('$name N=$N', round(d['value'],2), round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d.get('replicas_identical'), d['cuda_graph'], d['clocks']['sm_mhz'])"
  grep -i "NVLS\|error" gpurun_out/mgpu8_${name}_n$N.err | head -3
}
run buckets8 HM_COMM_SMS=8
run single HM_BUCKETS=0 NCCL_MAX_CTAS=32
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-torch-gpu --no-alt > gpurun_out/mgpu8_n1.json 2> gpurun_out/mgpu8_n1.err
python -c "
import json; d=json.load(open('gpurun_out/mgpu8_n1.json')); print('N=1', round(d['value'],2), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
