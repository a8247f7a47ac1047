#!/bin/bash
# Sweep of the pull-layout build knobs (GT_PULL_*); prints GTEPS and per-phase ms from bench.py.
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
run GT_PULL_VROW=2048
run GT_PULL_VROW=512
run GT_PULL_VROW=128
run GT_PULL_VROW=512 GT_PULL_HOT=12000 GT_PULL_CTAS=2
run GT_PULL_VROW=512 GT_PULL_HOT=6000 GT_PULL_CTAS=2
run GT_PULL_VROW=512 GT_PULL_HOT=6000 GT_PULL_CTAS=4 GT_PULL_THREADS=512
run GT_PULL_VROW=512 GT_PULL_HOT=0 GT_PULL_CTAS=2
