#!/bin/bash
# Sweep of one environment knob over both bench shapes.  Usage: gpurun -- 'bash scripts/gpu_sweep.sh VAR v1 v2 ...'
VAR=$1; shift
mkdir -p gpurun_out
free -g | head -2; nproc
for v in "$@"; do
  for shape in "" "--scale 0.1 --coverage 1000"; do
    env $VAR=$v python bench.py --steps 10 --warmup 3 --no-cpu $shape 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['config']['kernel_ms']; print('$VAR=$v', '$shape', 'tally %.4f fit %.4f step %.3f' % (k['tally'], k['fit'], d['ms_per_step']))"
  done
done
