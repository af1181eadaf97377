#!/bin/bash
# One multi-GPU visit (under gpurun --gpus N): config 3 (device and host inflate), config 5 sweep.  Usage: bash tools/n8_round.sh <tag> <N> [pairs]
TAG=$1; N=$2; PAIRS=${3:-50000000}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
nproc > gpurun_out/${TAG}_n${N}_box.txt; free -g >> gpurun_out/${TAG}_n${N}_box.txt
SS_DEBUG_TIMING=1 timeout 420 $T bench.py --gpus $N --config c3 --pairs $PAIRS --steps 3 --warmup 1 > gpurun_out/${TAG}_c3_n${N}.json 2> gpurun_out/${TAG}_c3_n${N}.err
tail -c 1200 gpurun_out/${TAG}_c3_n${N}.json; grep "rror\|Traceback" gpurun_out/${TAG}_c3_n${N}.err | tail -5
if [ "$5" != "nohost" ]; then
  SS_DGZ=0 timeout 420 $T bench.py --gpus $N --config c3 --pairs $PAIRS --steps 2 --warmup 1 > gpurun_out/${TAG}_c3_n${N}_hostinflate.json 2> gpurun_out/${TAG}_c3_n${N}_hostinflate.err
  tail -c 1200 gpurun_out/${TAG}_c3_n${N}_hostinflate.json; grep "rror\|Traceback" gpurun_out/${TAG}_c3_n${N}_hostinflate.err | tail -5
fi
if [ "$4" = "c5" ]; then
  timeout 420 $T bench.py --gpus $N --config c5 --steps 4 --warmup 2 > gpurun_out/${TAG}_c5_n${N}.json 2> gpurun_out/${TAG}_c5_n${N}.err
  tail -c 3000 gpurun_out/${TAG}_c5_n${N}.json; grep "rror\|Traceback" gpurun_out/${TAG}_c5_n${N}.err | tail -5
fi
