#!/bin/bash
# 1 GPU, final code: suite, smoke, NS configs, bench
O=gpurun_out/r2p; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
for cfg in "bfs --scale 22" "sssp --scale 25" "cc --scale 24"; do
  timeout 300 python tools/run_config.py $cfg --repeat 4 2>&1 | grep -v "^Execute" >> $O/configs.log
done
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
echo done > $O/done
