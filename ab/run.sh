#!/bin/bash
# ab/run.sh <tag>...  : headline window + 200-step loop for each A/B library (tag "base" = the shipping library)
for t in "$@"; do
  if [ $t = base ]; then unset SIHMPC_B200_LIB; else export SIHMPC_B200_LIB=/root/repo/ab/lib_$t.so; fi
  python bench.py --no-configs --no-cpu-baseline --e2e-steps 2 --repeats 5 > gpurun_out/ab_${t}_w.json 2> gpurun_out/ab_${t}_w.err
  python bench.py --steps 200 --no-configs --no-cpu-baseline --e2e-steps 2 --repeats 3 > gpurun_out/ab_${t}_200.json 2> gpurun_out/ab_${t}_200.err
  python - <<PY
import json
for k in ("w","200"):
    try:
        d=json.loads(open("gpurun_out/ab_${t}_%s.json"%k).read().strip().splitlines()[-1])
        print("${t}",k,"value %.4g"%d["value"],"ms %.3f"%d["ms_timed_region"]["median"],"kernel_ms",d["roofline"]["kernel_ms_per_launch"],"frac %.3f"%d["roofline"]["frac"],"handover",d.get("handed_over_to_general_kernel"),"inkernel",d.get("constrained_steps_resolved_in_fast_path"),"status_or",d.get("status_or"))
    except Exception as e: print("${t}",k,"FAILED",e)
PY
done
