#!/bin/bash
# Round-2 GPU call C (1 GPU): TCSC_CF + non-stationary engine fixes — parity tests, configs, bench.
O=gpurun_out/r2c; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for cfg in "bfs --scale 22" "sssp --scale 25" "cc --scale 24"; do
  timeout 300 python tools/run_config.py $cfg --repeat 4 2>&1 | grep -v "^Execute" >> $O/configs.log
done
GT_PULL_VERBOSE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -k regex:k_ns_ -c 300 --csv --log-file $O/ncu_sssp25.csv python tools/run_config.py sssp --scale 25 --repeat 1 > $O/ncu_sssp25.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:k_ns_ -c 300 --csv --log-file $O/ncu_bfs22.csv python tools/run_config.py bfs --scale 22 --repeat 1 > $O/ncu_bfs22.log 2>&1
echo done > $O/done
