#!/bin/bash
# 2-GPU box: full GPU suite (partitioned ingest at p=2 inside multi_gpu_check), NS configs with the filter-read variants,
# SSSP-25 / CC-25 at 2 GPUs with the leaner per-iteration path + timeline
O=gpurun_out/r2j; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for lib in default f1 f2 f3; do
  if [ $lib != default ]; then export GT_LIB=$PWD/build/variants/libgraphtap_b200.$lib.so; else unset GT_LIB; fi
  for cfg in "bfs --scale 22" "sssp --scale 25" "cc --scale 24"; do
    echo "== $lib $cfg" >> $O/configs.log
    CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/run_config.py $cfg --repeat 4 2>&1 | grep -v "^Execute" >> $O/configs.log
  done
done
unset GT_LIB
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562"
GT_TIMELINE=$O/tl_sssp timeout 300 $TR tools/run_config.py sssp --scale 25 --repeat 4 2>&1 | grep -v "^Execute" > $O/sssp_p2.log
timeout 300 $TR tools/run_config.py cc --scale 25 --repeat 4 2>&1 | grep -v "^Execute" > $O/cc_p2.log
timeout 300 $TR tools/run_config.py bfs --scale 22 --repeat 4 2>&1 | grep -v "^Execute" > $O/bfs_p2.log
echo done > $O/done
