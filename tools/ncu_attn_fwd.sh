#!/bin/bash
# ncu --set full of the attention forward variants at the bench shape (dev tool): reports land in gpurun_out/
mkdir -p gpurun_out
for v in ${VARIANTS:-3 4}; do
  CLIMB_ATTN_FWD=$v B=64 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd -s 8 -c 1 -f \
      -o gpurun_out/prof_attn_fwd$v python tools/perf_attn.py > gpurun_out/ncu_attn_fwd$v.log 2>&1
  echo "variant $v exit=$?"
done
