#!/bin/bash
# ab/run_log.sh <tag>... : log-variant parity tests + configs[1] numbers for each A/B library (tag "base" = the shipping library)
for t in "$@"; do
  if [ $t = base ]; then unset SIHMPC_B200_LIB; else export SIHMPC_B200_LIB=/root/repo/ab/lib_$t.so; fi
  python -m pytest tests/test_gpu_step.py -m gpu -q -x -k "log_variant" 2>&1 | tail -2
  echo "$t $(python tools/log_bench.py 20 5 2> gpurun_out/ab_log_${t}.err | tee gpurun_out/ab_log_${t}.json)"
done
