#!/bin/bash
# 2 GPUs: the ingest exchange driven by gt::make_route_plan (both transports), parity block
O=gpurun_out/r2t; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612"
timeout 600 $TR tools/multi_gpu_check.py > $O/check.log 2>&1; echo "check rc=$?" >> $O/check.log
timeout 300 $TR tools/ingest_check.py cc --scale 24 > $O/ingest_cc24.json 2> $O/ingest_cc24.err; echo "rc=$?" >> $O/ingest_cc24.err
GT_PEER=0 timeout 300 $TR tools/ingest_check.py sssp --scale 20 --file > $O/ingest_sssp20_nccl.json 2> $O/ingest_sssp20_nccl.err; echo "rc=$?" >> $O/ingest_sssp20_nccl.err
echo done > $O/done
