#!/bin/bash
# A/B of V-cycle variants: tools/ab_vcycle.sh <workload> <steps> <warmup> "VAR=val ..." "VAR=val ..." (one run per quoted configuration)
wl=${1:-lid_driven2D_nx707}; K=${2:-10}; W=${3:-3}; shift 3
for cfg in "$@"; do
  env $cfg python bench.py --workload $wl --steps $K --warmup $W --no-cpu-baseline --no-e2e --no-extra 2>gpurun_out/ab_err.log | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$wl', '$cfg', 'ms/step %.2f' % l['ms_per_step'], 'its/step %.1f' % l['solve']['fgmres_its_per_step'], 'launches', l['gpu_launches'])
" || tail -5 gpurun_out/ab_err.log
done
