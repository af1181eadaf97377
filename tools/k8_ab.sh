#!/bin/bash
# K8 A/B on one GPU: the gzip parity tests on every decoder form, then config 3 (reduced) with the file pass timed under
# each SS_DGZ_LANES setting on the same files, then one ncu capture of the lanes kernel.  Usage: bash tools/k8_ab.sh <tag> [pairs] [variants]
TAG=$1; PAIRS=${2:-20000000}; VARIANTS=${3:-"SS_DGZ_LANES=4;SS_DGZ_LANES=8;SS_DGZ_LANES=16;SS_DGZ_LANES=2"}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "device_inflate_of_ordinary_gzip" > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
SS_DEBUG_TIMING=1 timeout 400 python bench.py --config c3 --pairs $PAIRS --steps 2 --warmup 1 --e2e-variants "$VARIANTS" > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_c3.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("default e2e ms", e["ms_per_step"], "value", e["value"])
    for v in e.get("variants", []):
        print(v["env"], "ms", v["ms_per_step"], "value", v["value"])
except Exception as ex:
    print("no bench line:", ex)
PY
grep "ss dgz" gpurun_out/${TAG}_c3.err | tail -24 | cut -c1-400
grep "rror\|Traceback\|assert" gpurun_out/${TAG}_c3.err | tail -5
if [ "$4" = "ncu" ]; then
  SS_DGZ_LANES=${5:-4} timeout 200 ncu --set full --clock-control none --import-source on -k regex:ss_dgz_decode_lanes -c 1 -f -o gpurun_out/${TAG}_dgz2 \
      python bench.py --config c3 --pairs 2000000 --steps 1 --warmup 0 > gpurun_out/${TAG}_ncu.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-300
fi
