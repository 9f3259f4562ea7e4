#!/bin/bash
# Quick GPU-box visit: parity tests, the C1 bench line, the C2-shaped bench line (a tenth of the genome at 1000x), and the
# active/elapsed SM cycles of the tally kernel at both shapes (load balance).  Usage: gpurun -- 'bash scripts/gpu_quick.sh <tag> [tests|notests]'
TAG=${1:-run}; TESTS=${2:-tests}
mkdir -p gpurun_out
if [ "$TESTS" = tests ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -8; fi
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python bench.py --steps 5 --warmup 3 --no-cpu --scale 0.1 --coverage 1000 > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err
python - <<PY
import json
for f in ["gpurun_out/bench_$TAG.json", "gpurun_out/bench_${TAG}_c2.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(d["ms_per_step"], d["config"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"])
    except Exception as e: print(f, "failed", e)
PY
M=sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_elapsed.max,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -k regex:tally_kernel -s 3 -c 1 --csv --log-file gpurun_out/balance_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:tally_kernel -s 3 -c 1 --csv --log-file gpurun_out/balance_${TAG}_c2.csv python bench.py --steps 1 --warmup 3 --no-cpu --scale 0.1 --coverage 1000 > /dev/null 2>&1
grep -h tally gpurun_out/balance_$TAG.csv gpurun_out/balance_${TAG}_c2.csv | cut -d, -f5,12- | cut -c1-200
