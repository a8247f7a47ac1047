#!/bin/bash
# 1 GPU: dynamic slice dealing A/B on the headline kernel
O=gpurun_out/r2r; mkdir -p $O
sw() {
  echo "== $*" >> $O/sweep.log
  env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs 2>> $O/sweep.err | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
p = d['roofline']['phases_ms']
print('GTEPS %.1f  combine %.3f ms  scatter %.3f  apply %.3f  frac %.3f  iter %.3f ms  sum %.9e' % (d['value'], p['combine'], p['scatter_gather'], p['apply'], d['roofline']['frac'], d['roofline']['iteration_ms'], d['config']['rank_sum_global']))" >> $O/sweep.log 2>&1
}
sw GT_PULL_DYNAMIC=0
sw GT_PULL_DYNAMIC=1
sw GT_PULL_DYNAMIC=4
sw GT_PULL_DYNAMIC=16
M=gpu__time_duration.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.max
GT_PULL_DYNAMIC=4 ncu --metrics $M --clock-control none -k regex:k_spmv_pull -s 8 -c 2 --csv --log-file $O/ncu_dyn4.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > $O/ncu_dyn4.log 2>&1
echo done > $O/done
