#!/bin/bash
# 2-GPU box: the whole GPU suite (includes tools/multi_gpu_check.py at p=2 with the partitioned-ingest block)
O=gpurun_out/r2i; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
timeout 600 $TR tools/multi_gpu_check.py > $O/check_p2.log 2>&1; echo "check rc=$?" >> $O/check_p2.log
echo done > $O/done
