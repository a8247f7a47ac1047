#!/bin/bash
# Turns the artefacts of tools/gpu_r2k.sh (gpurun_out/r2k) into the tracked summaries under profiles/.
set -e
cd "$(dirname "$0")/.."
O=gpurun_out/r2k
cp $O/bench_n1.json profiles/r02_bench_n1.json
for c in sssp25 bfs22 cc24; do grep '^"' $O/ncu_$c.csv > profiles/r02_ncu_ns_$c.csv; done
{
  echo "# Round 2 — ncu launch lists of one execute() of the non-PageRank configs (1 B200; cold-cache, serialised: shares, not absolutes)"
  echo
  echo "\`ncu --metrics gpu__time_duration.sum,dram__bytes_{read,write}.sum,{l1tex,lts}__throughput…,gpu__dram_throughput…,hit rates --clock-control none -k regex:k_ns_\`"
  echo "over \`tools/run_config.py <app> --scale <s> --repeat 1\`; raw rows in \`r02_ncu_ns_<config>.csv\`, summarised by \`tools/ncu_launches.py\`."
  for c in sssp25 bfs22 cc24; do
    echo; echo "## $c"; echo; echo '```'
    python tools/ncu_launches.py profiles/r02_ncu_ns_$c.csv --per-launch --min-us 30
    echo '```'
  done
} > profiles/r02_ncu_ns_summary.md
for k in dense spmspv apply; do python tools/ncu_summary.py $O/full_ns_$k.ncu-rep > profiles/r02_ncu_ns_full_$k.txt; done
python tools/ncu_summary.py $O/full_pull.ncu-rep > profiles/r02_ncu_pull_s26_full.txt
grep '^"' $O/launches_bench.csv > profiles/r02_launches_bench_s26.csv
python tools/ncu_launches.py profiles/r02_launches_bench_s26.csv > profiles/r02_launches_bench_s26_summary.txt
