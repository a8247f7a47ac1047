#!/bin/bash
# 2-GPU rehearsal of the config-scale ingest check + the final 2-GPU bench line
O=gpurun_out/r2m; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29582"
timeout 300 $TR tools/ingest_check.py cc --scale 24 > $O/ingest_cc24.json 2> $O/ingest_cc24.err; echo "rc=$?" >> $O/ingest_cc24.err
timeout 300 $TR tools/ingest_check.py pr --scale 20 --file > $O/ingest_pr20_file.json 2> $O/ingest_pr20_file.err; echo "rc=$?" >> $O/ingest_pr20_file.err
timeout 300 $TR tools/ingest_check.py sssp --scale 22 > $O/ingest_sssp22.json 2> $O/ingest_sssp22.err; echo "rc=$?" >> $O/ingest_sssp22.err
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
echo done > $O/done
