#!/bin/bash
# adapter-config step time (BASELINE config 3 shape: NLVR2 pairs, Houlsby rf 16, base frozen), fused vs two-launch bottleneck
mkdir -p gpurun_out
for f in 1 0; do
  echo "== CLIMB_ADAPTER_FUSED=$f"
  CLIMB_ADAPTER_FUSED=$f MODE=adapters timeout 300 python tools/step_time.py
done
