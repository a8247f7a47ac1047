#!/bin/bash
# N-GPU: RED + flag for follower rows, ingest exchange over a world peer window
N=${1:-4}; O=gpurun_out/r2n_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2959$N"
timeout 900 $TR tools/multi_gpu_check.py > $O/check.log 2>&1; echo "check rc=$?" >> $O/check.log
timeout 300 $TR tools/ingest_check.py cc --scale 25 > $O/ingest_cc25.json 2> $O/ingest_cc25.err; echo "rc=$?" >> $O/ingest_cc25.err
timeout 300 $TR tools/ingest_check.py pr --scale 20 --file > $O/ingest_pr20_file.json 2> $O/ingest_pr20_file.err; echo "rc=$?" >> $O/ingest_pr20_file.err
GT_PEER=0 timeout 300 $TR tools/ingest_check.py sssp --scale 20 > $O/ingest_sssp20_nccl.json 2> $O/ingest_sssp20_nccl.err; echo "rc=$?" >> $O/ingest_sssp20_nccl.err
GT_TIMELINE=$O/tl_sssp timeout 300 $TR tools/run_config.py sssp --scale 25 --repeat 4 2>&1 | grep -v "^Execute" > $O/sssp.log
timeout 300 $TR tools/run_config.py cc --scale 26 --repeat 3 2>&1 | grep -v "^Execute" > $O/cc.log
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
echo done > $O/done
