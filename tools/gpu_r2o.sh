#!/bin/bash
# 8 GPUs, final code: bench line (with parity block and other configs), config-scale ingest over the peer window, SSSP timeline
N=8; O=gpurun_out/r2o_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29608"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
timeout 600 $TR tools/ingest_check.py cc --scale 27 > $O/ingest_cc27.json 2> $O/ingest_cc27.err; echo "rc=$?" >> $O/ingest_cc27.err
GT_TIMELINE=$O/tl_sssp timeout 300 $TR tools/run_config.py sssp --scale 25 --repeat 3 2>&1 | grep -v "^Execute" > $O/sssp.log
echo done > $O/done
