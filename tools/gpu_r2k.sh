#!/bin/bash
# 1 GPU: suite, NS configs (leaner applicator), ncu launch lists + one --set full capture per dominant NS kernel, bench + its launch list
O=gpurun_out/r2k; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for cfg in "bfs --scale 22" "sssp --scale 25" "cc --scale 24"; do
  timeout 300 python tools/run_config.py $cfg --repeat 4 2>&1 | grep -v "^Execute" >> $O/configs.log
done
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct
for c in "sssp 25" "bfs 22" "cc 24"; do set -- $c
timeout 600 ncu --metrics $M --clock-control none -k regex:k_ns_ -c 300 --csv --log-file $O/ncu_$1$2.csv python tools/run_config.py $1 --scale $2 --repeat 1 > $O/ncu_$1$2.log 2>&1
done
# one full capture each: the third dense pass, the largest frontier pass and a mid-run applicator of SSSP-25
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ns_dense -s 4 -c 1 -o $O/full_ns_dense python tools/run_config.py sssp --scale 25 --repeat 1 > $O/full_ns_dense.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ns_spmspv -s 5 -c 1 -o $O/full_ns_spmspv python tools/run_config.py sssp --scale 25 --repeat 1 > $O/full_ns_spmspv.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ns_apply -s 3 -c 1 -o $O/full_ns_apply python tools/run_config.py sssp --scale 25 --repeat 1 > $O/full_ns_apply.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_spmv_pull -s 8 -c 1 -o $O/full_pull python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > $O/full_pull.log 2>&1
# launch list of the bench command itself (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > $O/launches_bench.log 2>&1
echo done > $O/done
