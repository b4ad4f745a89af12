#!/bin/bash
# N = 1, 2, 4, 8 back to back on ONE box (what the driver's scaling run does), full JSON lines into gpurun_out/scale_N.json
mkdir -p gpurun_out
for N in ${@:-1 2 4 8}; do
  if [ "$N" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29620 + N)) \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  fi
  python -c "import json,sys; d=json.loads(open('gpurun_out/scale_$N.json').read().strip().splitlines()[-1]); print('N=$N value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'clk', d['clocks']['sm_mhz'])"
done
