#!/bin/bash
# Rebuild the probe kernel with a few tunable combinations on the GPU box and time each.
# combo = UNROLL:MINCTAS[:CARVEOUT_PCT[:CTAS_PER_SM]]
# Usage (under gpurun): bash tools/sweep.sh "3:4 2:4:72 3:3:58:3"
for combo in ${1:-"3:4"}; do
  IFS=: read U M CV CT <<< "$combo"
  make -C strainscan_b200/csrc clean > /dev/null
  make -C strainscan_b200/csrc -j8 SS_DEFS="-DSS_PROBE_UNROLL=$U -DSS_PROBE_MIN_CTAS=$M $SS_EXTRA_DEFS" > /dev/null 2>&1 || { echo "build failed $combo"; continue; }
  echo "== UNROLL=$U MINCTAS=$M CARVEOUT=${CV:-default} CTAS=${CT:-auto}: $(grep -A3 'ILb1ELb0' strainscan_b200/csrc/ss_probe.ptxas.log | grep -o 'Used [0-9]* registers.*smem')"
  env ${CV:+SS_CARVEOUT=$CV} ${CT:+SS_PROBE_CTAS=$CT} python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
r = d['roofline']
print('%.4g kmers/s probe %.3f ms frac %.3f table_probe_rate %.4f' % (d['value'], r['launch_ms'], r['frac'], d['config'].get('table_probe_rate')))"
done
