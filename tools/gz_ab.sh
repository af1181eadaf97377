#!/bin/bash
# A/B the device inflate kernel's launch bounds on the GPU box: bash tools/gz_ab.sh "16 8"
for M in ${1:-16}; do
  (cd strainscan_b200/csrc && rm -f ss_gunzip.o && make SS_DEFS="-DSS_GUNZIP_MINCTAS=$M $SS_EXTRA_DEFS" > /dev/null 2>&1; grep -A2 gunzip_kernel ss_gunzip.ptxas.log | tail -2 | tr '\n' ' '; echo)
  python tools/bench_ingest.py --no-reference --only-bgzf 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); [print('MINCTAS=$M', k, {a:round(b,4) for a,b in v.items() if a in ('ingest_s','stream_count_files_s','text_gb_per_s')}) for k,v in d['cases'].items()]"
done
