#!/bin/bash
# A/B runs of library variants on the bench workload (resident pass only).  Usage: tools/ab.sh <tag> <variant>...
TAG=$1; shift
mkdir -p gpurun_out
for v in default "$@"; do
  if [ "$v" = default ]; then unset SS_LIB_PATH; else export SS_LIB_PATH=$PWD/build/variants/$v/libstrainscan_b200.so; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ab_$v.json 2> gpurun_out/${TAG}_ab_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_ab_$v.json"))
    print("$v", "value %.4g" % d["value"], "rescan %.4g" % d["value_rescan"], d["kernel_ms"], "probe rate %.4f" % d["config"]["table_probe_rate"])
except Exception as e:
    print("$v", "FAILED", e); print(open("gpurun_out/${TAG}_ab_$v.err").read()[-1500:])
PY
done
unset SS_LIB_PATH
