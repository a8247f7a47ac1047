#!/bin/bash
# Round-2 GPU call A (1 GPU): tests, bench (both arms), single-GPU PageRank hot-band experiments.
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/smi.txt
nproc > $O/host.txt; free -g >> $O/host.txt; df -h /tmp >> $O/host.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
( time timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err ) 2> $O/bench_ref.time
sw() {
  echo "== $*" >> $O/sweep.log
  env GT_PULL_VERBOSE=1 "$@" python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs 2>> $O/sweep.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
p = d['roofline']['phases_ms']
print('GTEPS %.1f  combine %.3f ms  scatter %.3f  apply %.3f  frac %.3f  sum %.9e' % (d['value'], p['combine'], p['scatter_gather'], p['apply'], d['roofline']['frac'], d['config']['rank_sum_global']))" >> $O/sweep.log 2>&1
}
sw GT_PULL_BAND=0
for b in 8192 16384 24576 49152 131072; do sw GT_PULL_BAND=$b; done
for b in 8192 12288 16384 24576 28000; do sw GT_PULL_BAND=$b GT_PULL_BAND_SMEM=1; done
sw GT_PULL_BAND=12288 GT_PULL_BAND_SMEM=1 GT_PULL_VROW=128
M=gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
nc() {
  tag=$1; shift
  env "$@" ncu --metrics $M --clock-control none -k regex:k_spmv_pull -s 8 -c 4 --csv --log-file $O/ncu_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > $O/ncu_$tag.log 2>&1
}
nc band0 GT_PULL_BAND=0
nc band16k GT_PULL_BAND=16384
nc band49k GT_PULL_BAND=49152
nc smem16k GT_PULL_BAND=16384 GT_PULL_BAND_SMEM=1
nc smem28k GT_PULL_BAND=28000 GT_PULL_BAND_SMEM=1
echo done > $O/done
