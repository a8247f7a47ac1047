#!/bin/bash
O=gpurun_out/r2h; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for cfg in "bfs --scale 22" "sssp --scale 25" "cc --scale 24"; do
  timeout 300 python tools/run_config.py $cfg --repeat 4 2>&1 | grep -v "^Execute" >> $O/configs.log
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
for c in "sssp 25" "bfs 22" "cc 24"; do set -- $c
timeout 600 ncu --metrics $M --clock-control none -k regex:k_ns_ -c 300 --csv --log-file $O/ncu_$1$2.csv python tools/run_config.py $1 --scale $2 --repeat 1 > $O/ncu_$1$2.log 2>&1
done
echo done > $O/done
