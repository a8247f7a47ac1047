#!/bin/bash
# 1 GPU: GT_PULL_ROWSORT A/B on the headline kernel
O=gpurun_out/r2q; mkdir -p $O
sw() {
  echo "== $*" >> $O/sweep.log
  env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs 2>> $O/sweep.err | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
p = d['roofline']['phases_ms']
print('GTEPS %.1f  combine %.3f ms  scatter %.3f  apply %.3f  frac %.3f  sum %.9e' % (d['value'], p['combine'], p['scatter_gather'], p['apply'], d['roofline']['frac'], d['config']['rank_sum_global']))" >> $O/sweep.log 2>&1
}
sw GT_PULL_ROWSORT=0
sw GT_PULL_ROWSORT=1
M=gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors.sum
for v in 0 1; do
GT_PULL_ROWSORT=$v ncu --metrics $M --clock-control none -k regex:k_spmv_pull -s 8 -c 2 --csv --log-file $O/ncu_rowsort$v.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > $O/ncu_rowsort$v.log 2>&1
done
echo done > $O/done
