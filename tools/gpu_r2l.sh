#!/bin/bash
# N-GPU validation of the final code: parity block (incl. partitioned ingest), bench line, config-scale ingest check, SSSP timeline
N=${1:-8}; O=gpurun_out/r2l_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N"
timeout 900 $TR tools/multi_gpu_check.py > $O/check.log 2>&1; echo "check rc=$?" >> $O/check.log
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
if [ $N = 8 ]; then
  timeout 600 $TR tools/ingest_check.py cc --scale 27 > $O/ingest_cc27.json 2> $O/ingest_cc27.err; echo "rc=$?" >> $O/ingest_cc27.err
  timeout 300 $TR tools/ingest_check.py pr --scale 20 --file > $O/ingest_pr20_file.json 2> $O/ingest_pr20_file.err; echo "rc=$?" >> $O/ingest_pr20_file.err
fi
GT_TIMELINE=$O/tl_sssp timeout 300 $TR tools/run_config.py sssp --scale 25 --repeat 3 2>&1 | grep -v "^Execute" > $O/sssp.log
echo done > $O/done
