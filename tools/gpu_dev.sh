#!/bin/bash
# Dev loop on the GPU box: every pytest group in its own process (a trapped kernel kills the CUDA
# context) under its own timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, timeout, args...
  local name=$1; shift; local to=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to python -m pytest -q -m gpu "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? $(tail -n 3 gpurun_out/$name.log | tr '\n' ' ')" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
for g in "$@"; do
  case $g in
    gemm_k)   run gemm_k 600 tests/test_gpu_kernels.py -k "gemm_kmajor" ;;
    gemm_mn)  run gemm_mn 600 tests/test_gpu_kernels.py -k "gemm_mn_major" ;;
    gemm_x)   run gemm_x 600 tests/test_gpu_kernels.py -k "gemm_wgrad or gemm_epilogues or gemm_large or gemm_rejects" ;;
    attn)     run attn 600 tests/test_gpu_kernels.py -k "attention" ;;
    ln)       run ln 600 tests/test_gpu_kernels.py -k "layernorm" ;;
    parity)   run parity 900 tests/test_gpu_parity.py -s ;;
    cl)       run cl 900 tests/test_gpu_cl.py -s ;;
    image)    run image 600 tests/test_image_pre.py ;;
    viltbert) run viltbert 900 tests/test_gpu_viltbert.py -s ;;
    perf)     echo "=== perf" | tee -a gpurun_out/summary.txt; timeout 600 python tools/perf_kernels.py > gpurun_out/perf.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary.txt; cat gpurun_out/perf.log | tee -a gpurun_out/summary.txt ;;
    perfgen)  echo "=== perf (generic GEMM kernel only)" | tee -a gpurun_out/summary.txt; CLIMB_GEMM_GENERIC=1 timeout 600 python tools/perf_kernels.py > gpurun_out/perf_generic.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary.txt; head -n 12 gpurun_out/perf_generic.log | tee -a gpurun_out/summary.txt ;;
    gemm_fast) run gemm_fast 600 tests/test_gpu_kernels.py -k "gemm_fast" ;;
    pair)     run pair 600 tests/test_gpu_kernels.py -k "gemm_fast or wgrad" ;;      # both pair modes (pair_mode fixture)
    pair_ab)  echo "=== pair_ab (per-GEMM timings, one-CTA vs CTA-pair kernels)" | tee -a gpurun_out/summary.txt
              for p in 0 1; do echo "-- CLIMB_GEMM_PAIR=$p" | tee -a gpurun_out/summary.txt; ONLY_GEMM=1 CLIMB_GEMM_PAIR=$p timeout 120 python tools/perf_kernels.py 2>&1 | tee -a gpurun_out/summary.txt; done ;;
    ncu_pair) echo "=== ncu_pair" | tee -a gpurun_out/summary.txt; timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 2 -c 2 -f -o gpurun_out/prof_pair python tools/ncu_pair.py > gpurun_out/ncu_pair.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary.txt ;;
    trainer)  run trainer 600 tests/test_gpu_zz_trainer.py -s ;;
    smoke)    echo "=== smoke" | tee -a gpurun_out/summary.txt; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit=$? $(tail -n 2 gpurun_out/smoke.log | tr '\n' ' ')" | tee -a gpurun_out/summary.txt ;;
    bench)    echo "=== bench" | tee -a gpurun_out/summary.txt; timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit=$?" | tee -a gpurun_out/summary.txt; tail -n 5 gpurun_out/bench.err | tee -a gpurun_out/summary.txt; cat gpurun_out/bench.json | tee -a gpurun_out/summary.txt ;;
    benchref) echo "=== benchref" | tee -a gpurun_out/summary.txt; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit=$?" | tee -a gpurun_out/summary.txt; cat gpurun_out/bench_ref.json | tee -a gpurun_out/summary.txt ;;
    launches) echo "=== launches" | tee -a gpurun_out/summary.txt; timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/one_step.py > gpurun_out/launches.log 2>&1; echo "exit=$? rows=$(wc -l < gpurun_out/launches.csv)" | tee -a gpurun_out/summary.txt ;;
    launches_adapters) echo "=== launches_adapters" | tee -a gpurun_out/summary.txt; MODE=adapters timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_adapters.csv python tools/one_step.py > gpurun_out/launches_adapters.log 2>&1; echo "exit=$? rows=$(wc -l < gpurun_out/launches_adapters.csv)" | tee -a gpurun_out/summary.txt ;;
    ncu_gemm) echo "=== ncu_gemm" | tee -a gpurun_out/summary.txt; WARM=2 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16_tcgen05 -s 40 -c 4 -f -o gpurun_out/prof_gemm python tools/one_step.py > gpurun_out/ncu_gemm.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary.txt ;;
    ncu_gemm_bwd) echo "=== ncu_gemm_bwd" | tee -a gpurun_out/summary.txt; WARM=2 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16_tcgen05 -s ${NCU_SKIP:-62} -c ${NCU_COUNT:-8} -f -o gpurun_out/prof_gemm_bwd python tools/one_step.py > gpurun_out/ncu_gemm_bwd.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary.txt ;;
    ncu_fast) echo "=== ncu_fast" | tee -a gpurun_out/summary.txt; WARM=2 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_fast -s ${NCU_SKIP:-20} -c ${NCU_COUNT:-4} -f -o gpurun_out/prof_fast python tools/one_step.py > gpurun_out/ncu_fast.log 2>&1
              WARM=2 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_fast -s ${NCU_SKIP2:-62} -c ${NCU_COUNT2:-5} -f -o gpurun_out/prof_fast_bwd python tools/one_step.py >> gpurun_out/ncu_fast.log 2>&1; echo "exit=$?" | tee -a gpurun_out/summary.txt ;;
    ncu_attn) echo "=== ncu_attn" | tee -a gpurun_out/summary.txt
              WARM=2 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_tc_fwd -s 4 -c 1 -f -o gpurun_out/prof_attn_fwd python tools/one_step.py > gpurun_out/ncu_attn.log 2>&1
              WARM=2 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_tc_bwd -s 4 -c 1 -f -o gpurun_out/prof_attn_bwd python tools/one_step.py >> gpurun_out/ncu_attn.log 2>&1
              echo "exit=$?" | tee -a gpurun_out/summary.txt ;;
    *)        run "$(echo $g | tr '/:. ' '____')" 900 $g ;;
  esac
done
cat gpurun_out/summary.txt
