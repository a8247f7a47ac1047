#!/bin/bash
# Round-2 GPU call B (1 GPU): the rewritten non-stationary engine — parity tests, the three configs, ncu launch lists.
O=gpurun_out/r2b; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for cfg in "bfs --scale 22" "sssp --scale 25" "cc --scale 24" "bfs --scale 24" "sssp --scale 22"; do
  timeout 300 python tools/run_config.py $cfg --repeat 4 >> $O/configs.log 2>&1
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -k regex:k_ns_ -c 300 --csv --log-file $O/ncu_bfs22.csv python tools/run_config.py bfs --scale 22 --repeat 1 > $O/ncu_bfs22.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:k_ns_ -c 300 --csv --log-file $O/ncu_sssp25.csv python tools/run_config.py sssp --scale 25 --repeat 1 > $O/ncu_sssp25.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:k_ns_ -c 300 --csv --log-file $O/ncu_cc24.csv python tools/run_config.py cc --scale 24 --repeat 1 > $O/ncu_cc24.log 2>&1
GT_BFS_BU=0 timeout 300 python - >> $O/configs.log 2>&1 <<'PY'
import sys, json; sys.path.insert(0, '.')
from graphtap_b200 import engine as E
E.Env.init(); E.Env.quiet = True
G = E.Graph(); G.load_rmat(22, directed=False, transpose=False, self_loops=False, parallel_edges=False)
for ratio in (0.0, 0.01, 0.05, 0.2):
    ts = []
    for _ in range(4):
        V = E.BFS_Program(G, False, False, True, E._ROW_); V.set("bfs_bottom_up_ratio", ratio); it = V.execute(); ts.append(V.timing().execute_ms); cs = V.checksum(quiet=True); V.free()
    print(json.dumps({"bfs22_bottom_up_ratio": ratio, "ms": ts, "it": it, "cs": cs}))
PY
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
echo done > $O/done
