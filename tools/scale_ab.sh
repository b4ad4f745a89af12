#!/bin/bash
# A/B of the data-parallel knobs on N GPUs of one box (dev tool): NCCL CTA cap, SM reserve of the persistent grids,
# optimizer deferred span by span.   usage: tools/scale_ab.sh N "variant args" ...
N=${1:-2}; shift
mkdir -p gpurun_out
run() {
  echo "== $*"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 --no-cpu-baseline --no-gpu-baseline "$@" 2>gpurun_out/scale_ab.err | \
    python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'clk', d['clocks']['sm_mhz'])"
}
if [ $# -eq 0 ]; then
  set -- "--nccl-ctas 0 --sm-reserve 0 --defer-optimizer 0" "--nccl-ctas 0 --sm-reserve 0 --defer-optimizer 1" \
         "--nccl-ctas 8 --defer-optimizer 1" "--nccl-ctas 4 --defer-optimizer 1" "--nccl-ctas 16 --defer-optimizer 1" \
         "--nccl-ctas 0 --sm-reserve 16 --defer-optimizer 1"
fi
for v in "$@"; do run $v; done
