#!/bin/bash
# 8-GPU validation + A/B
N=8; O=gpurun_out/r2f_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556"
timeout 900 $TR tools/multi_gpu_check.py > $O/check.log 2>&1; echo "check rc=$?" >> $O/check.log
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
GT_TIMELINE=$O/tl timeout 300 $TR bench.py --gpus $N --steps 1 --warmup 3 --no-other-configs --no-parity > $O/bench_tl.json 2> $O/bench_tl.err
GT_PULL_SPLIT_MIN=8 timeout 300 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-other-configs --no-parity > $O/bench_split8.json 2> $O/bench_split8.err
GT_PULL_SPLIT_MIN=32 timeout 300 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-other-configs --no-parity > $O/bench_split32.json 2> $O/bench_split32.err
GT_PEER_LANES=1 timeout 300 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-other-configs --no-parity > $O/bench_lanes1.json 2> $O/bench_lanes1.err
echo done > $O/done
