#!/bin/bash
# One GPU visit: parity tests, the bench line (both arms), the ncu launch list and one --set full capture of the
# probe kernel.  Usage (under gpurun): bash tools/gpu_round.sh <tag> [skip-tests] [skip-ref] [skip-ncu]
TAG=${1:-rXX}
mkdir -p gpurun_out
if [[ " $* " != *" skip-tests "* ]]; then
  python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.log
fi
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 4000 gpurun_out/${TAG}_bench.json
if [[ " $* " != *" skip-ref "* ]]; then
  ( time python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err ) 2>&1 | tail -4
  tail -c 2500 gpurun_out/${TAG}_ref.json
fi
if [[ " $* " != *" skip-ncu "* ]]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_launches_bench.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:ss_probe_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_probe \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_full_bench.log 2>&1
fi
ls -la gpurun_out/${TAG}_*
