#!/bin/bash
# A/B builds of the library with other compile-time knobs:  tools/build_variant.sh <name> "<-D...>"
# -> build/variants/<name>/libstrainscan_b200.so ; run with SS_LIB_PATH=<that file> (strainscan_b200/_lib.py).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; DEFS=$2
D=$ROOT/build/variants/$NAME
mkdir -p $D/strainscan_b200/csrc $D/include
cp $ROOT/strainscan_b200/csrc/*.cu $ROOT/strainscan_b200/csrc/*.cuh $ROOT/strainscan_b200/csrc/*.h $ROOT/strainscan_b200/csrc/Makefile $D/strainscan_b200/csrc/
cp $ROOT/include/*.h $D/include/
make -C $D/strainscan_b200/csrc -j4 SS_DEFS="$DEFS" > $D/build.log 2>&1 || (tail -20 $D/build.log; exit 1)
cp $D/strainscan_b200/libstrainscan_b200.so $D/libstrainscan_b200.so
grep -A2 "ss_probe_kernelILb1ELb0" $D/strainscan_b200/csrc/ss_probe.ptxas.log | grep "Used" | head -2
echo "built $D/libstrainscan_b200.so"
