#!/bin/bash
# Sweep of the pull-layout build knobs (GT_PULL_*); prints GTEPS and per-phase ms from bench.py.
#   tools/sweep_pull.sh <scale>                 the built-in list
#   tools/sweep_pull.sh <scale> KNOB=V ...      one configuration
scale=${1:-26}
cd "$(dirname "$0")/.."
run() {
  echo -n "$* -> "
  env "$@" python bench.py --scale $scale --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
p = d['roofline']['phases_ms']
print('GTEPS %.1f  combine %.3f ms  scatter %.3f  apply %.3f  frac %.3f' % (d['value'], p['combine'], p['scatter_gather'], p['apply'], d['roofline']['frac']))"
}
shift
if [ $# -gt 0 ]; then run "$@"; exit 0; fi
run GT_PULL_L1HOT=0
run GT_PULL_L1HOT=16000
run GT_PULL_L1HOT=24000
run GT_PULL_L1HOT=48000
run GT_PULL_L1HOT=0 GT_PULL_L2HINT=1
run GT_PULL_L1HOT=24000 GT_PULL_L2HINT=1
run GT_PULL_L1HOT=24000 GT_PULL_UNROLL=4
run GT_PULL_L1HOT=24000 GT_PULL_VROW=256
run GT_PULL_L1HOT=24000 GT_PULL_VROW=1024 GT_PULL_CTAS=1
