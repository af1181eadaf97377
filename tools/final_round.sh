#!/bin/bash
# Last GPU visit of a round on one GPU: the whole GPU test suite, smoke(), config 3 (reduced) with K8 variants timed on
# the same files, one ncu capture of the gzip decode kernel that ships.  Usage: bash tools/final_round.sh <tag> [pairs]
TAG=$1; PAIRS=${2:-20000000}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/${TAG}_smoke.log 2>&1
tail -2 gpurun_out/${TAG}_smoke.log
SS_DEBUG_TIMING=1 timeout 300 python bench.py --config c3 --pairs $PAIRS --steps 2 --warmup 1 \
    --e2e-variants "SS_DGZ_RING=0;SS_DGZ_LOCKSTEP=0;SS_DGZ_LANES=1,SS_DGZ_WARPS=32;SS_DGZ_LANES=1,SS_DGZ_WARPS=24" > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
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
grep "ss dgz" gpurun_out/${TAG}_c3.err | cut -c1-330 | tail -45
grep "rror\|Traceback\|assert" gpurun_out/${TAG}_c3.err | tail -5
timeout 150 ncu --set full --clock-control none --import-source on -k regex:ss_dgz_decode_lanes -c 1 -f -o gpurun_out/${TAG}_dgz2 \
    python bench.py --config c3 --pairs 4000000 --steps 1 --warmup 0 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-200
