#!/bin/bash
N=${1:-4}; O=gpurun_out/r2g_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29557"
GT_TIMELINE=$O/tl_sssp timeout 300 $TR tools/run_config.py sssp --scale 25 --repeat 3 > $O/sssp.log 2>&1
GT_TIMELINE=$O/tl_cc timeout 300 $TR tools/run_config.py cc --scale 26 --repeat 3 > $O/cc.log 2>&1
echo done > $O/done
