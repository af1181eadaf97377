#!/bin/bash
# Read-count sweep of the L1 pass on one GPU (BASELINE.json configs[4] at N = 1): bench.py at 1 M .. 100 M reads,
# one JSON line each (resident throughput; end to end up to 30 M reads -- the 100 M case would pin 33 GB of host text).
# Usage (under gpurun): bash tools/bench_reads_sweep.sh gpurun_out/r01z_reads_sweep.jsonl
OUT=${1:-gpurun_out/reads_sweep.jsonl}
: > "$OUT"
for R in 1000000 3000000 10000000 30000000 100000000; do
  EXTRA=""
  [ "$R" -gt 30000000 ] && EXTRA="--no-e2e"
  python bench.py --reads $R --steps 5 --warmup 3 --no-cpu-baseline $EXTRA >> "$OUT" 2>> "${OUT%.jsonl}.err"
done
python - "$OUT" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    e = d.get("e2e") or {}
    print("%9d reads: %.4g k-mers/s resident (%.2f ms/step, probe %.2f ms), e2e %s" % (
        d["config"]["reads_total"], d["value"], d["ms_per_step"], d["kernel_ms"]["probe"],
        ("%.4g" % e["value"]) if e.get("value") else "-"))
PY
