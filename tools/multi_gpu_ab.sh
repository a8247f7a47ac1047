#!/bin/bash
# A/B runs of the multi-GPU knobs, one torchrun per configuration (run under `gpurun --gpus N -- tools/multi_gpu_ab.sh N ...`;
# an N-GPU box is charged N x its wall time, so every line here is worth its seconds).
#
#   tools/multi_gpu_ab.sh N check                     parity of all apps at N ranks (tools/multi_gpu_check.py)
#   tools/multi_gpu_ab.sh N pr   [KNOB=V ...]         PageRank RMAT-26 bench line with the given environment
#   tools/multi_gpu_ab.sh N sssp [KNOB=V ...]         SSSP RMAT-25 (config #4)
#   tools/multi_gpu_ab.sh N cc   [KNOB=V ...]         CC RMAT-27 (config #5; needs 8 GPUs' memory)
#
# Open questions these answer: GT_PULL_SPLIT_MIN=4|8|16 at N = 8 (fewer, longer part-rows);
# GT_PEER=0 vs 1 for sssp / cc at N = 4, 8 (frontier puts were only timed at N = 2).
set -u
N=${1:?number of GPUs}; what=${2:?check|pr|sssp|cc}; shift 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400))"
tag=$(echo "$*" | tr ' =' '__')
case $what in
  check)
    env "$@" timeout 200 $TR tools/multi_gpu_check.py > gpurun_out/check_p${N}_${tag}.log 2>&1
    echo "rc=$? $(grep -c ' OK' gpurun_out/check_p${N}_${tag}.log) OK lines, $(grep -o 'MULTI_GPU_CHECK PASS' gpurun_out/check_p${N}_${tag}.log | wc -l)/$N ranks PASS"
    grep -i 'fail\|error\|timed out' gpurun_out/check_p${N}_${tag}.log | head -5 ;;
  pr)
    env "$@" timeout 200 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/pr_p${N}_${tag}.json 2> gpurun_out/pr_p${N}_${tag}.err
    grep '^{' gpurun_out/pr_p${N}_${tag}.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$*', 'GTEPS %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], 'rank_sum', d['config']['rank_sum_check'], d['roofline']['phases_ms'])" ;;
  sssp|cc)
    scale=25; [ $what = cc ] && scale=27
    env "$@" timeout 300 $TR tools/run_config.py $what --scale $scale 2> gpurun_out/${what}_p${N}_${tag}.err | grep '^{' | tee gpurun_out/${what}_p${N}_${tag}.json ;;
esac
