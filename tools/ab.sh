#!/bin/bash
# usage: tools/ab.sh <out-log> <rounds> <lib name in _abl/> ...   — alternating bench runs of alternative builds inside one gpurun call
out=$1; rounds=$2; shift 2
mkdir -p gpurun_out; : > gpurun_out/$out
for r in $(seq 1 $rounds); do
  for l in "$@"; do
    echo -n "$l " >> gpurun_out/$out
    JAERO_B200_LIB=$PWD/_abl/$l.so python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-e2e --no-saturation 2>>gpurun_out/$out.err | tail -1 >> gpurun_out/$out
  done
done
