#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list and one full capture of the
# scoring kernel.  Usage (here): gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh <tag> [tests|notests] [full|nofull]'
TAG=${1:-run}; TESTS=${2:-tests}; FULL=${3:-full}
mkdir -p gpurun_out
nvidia-smi -L
if [ "$TESTS" = tests ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -15; fi
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 2500 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_list_$TAG.log 2>&1
if [ "$FULL" = full ]; then
  ncu --set full --clock-control none --import-source on -k "regex:tally_kernel" -s 3 -c 1 -f -o gpurun_out/prof_tally_$TAG \
      python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
  tail -3 gpurun_out/ncu_full_$TAG.log
fi
ls -la gpurun_out
