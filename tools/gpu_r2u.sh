#!/bin/bash
# 1 GPU: fused BFS tile pass — suite + A/B
O=gpurun_out/r2u; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for f in 1 0; do
  for cfg in "bfs --scale 22" "bfs --scale 24"; do
    echo "== GT_NS_FUSED=$f $cfg" >> $O/configs.log
    GT_NS_FUSED=$f timeout 300 python tools/run_config.py $cfg --repeat 5 2>&1 | grep -v "^Execute" >> $O/configs.log
  done
done
echo done > $O/done
