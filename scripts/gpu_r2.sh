#!/bin/bash
# One GPU-box visit of round 2.  Usage: gpurun --timeout 1800 -- 'bash scripts/gpu_r2.sh <tag> <what...>'
#   what: tests | stage | bench_small | bench | bench_c1 | timing | ncu
TAG=${1:-run}; shift
mkdir -p gpurun_out
for what in "$@"; do
  case $what in
    tests) python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ;;
    stage) BRQ_STAGE_TIMES=1 python scripts/stage_device_time.py 1.0 2>&1 | grep -v "^stage:" | tail -24 ;;
    bench_small) python bench.py --steps 5 --warmup 3 --scale 0.1 > gpurun_out/bench_${TAG}_small.json 2> gpurun_out/bench_${TAG}_small.err; tail -c 3000 gpurun_out/bench_${TAG}_small.json; tail -5 gpurun_out/bench_${TAG}_small.err ;;
    bench) python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 3500 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err ;;
    bench_c1) python bench.py --steps 10 --warmup 3 --config c1 > gpurun_out/bench_${TAG}_c1.json 2> gpurun_out/bench_${TAG}_c1.err; tail -c 3500 gpurun_out/bench_${TAG}_c1.json; tail -5 gpurun_out/bench_${TAG}_c1.err ;;
    timing) BRQ_TIMING=1 python bench.py --steps 2 --warmup 3 --no-cpu --no-bam > gpurun_out/timing_${TAG}.json 2> gpurun_out/timing_${TAG}.err; grep "evidence_export\|score:" gpurun_out/timing_${TAG}.err | tail -4 ;;
    ncu)   # the default workload (C2, full size): launch list of two steps, then one full capture of the tally kernel
      ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
          python bench.py --steps 2 --warmup 3 --no-cpu --no-bam > gpurun_out/ncu_list_$TAG.log 2>&1
      ncu --set full --clock-control none --import-source on -k "regex:tally_kernel" -s 3 -c 1 -f -o gpurun_out/prof_tally_$TAG \
          python bench.py --steps 1 --warmup 3 --no-cpu --no-bam > gpurun_out/ncu_full_$TAG.log 2>&1
      tail -3 gpurun_out/ncu_full_$TAG.log ;;
  esac
done
ls -la gpurun_out | tail -12
