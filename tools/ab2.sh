#!/bin/bash
# usage: tools/ab2.sh <out-log> <rounds> "<bench args A>" "<bench args B>" ...   — alternating bench configurations inside one gpurun call
out=$1; rounds=$2; shift 2
mkdir -p gpurun_out; : > gpurun_out/$out
for r in $(seq 1 $rounds); do
  for a in "$@"; do
    echo -n "[$a] " >> gpurun_out/$out
    python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-saturation $a 2>>gpurun_out/$out.err | tail -1 >> gpurun_out/$out
  done
done
