#!/bin/bash
# A/B of library builds on one box: bash scripts/ab_variants.sh "<bench args>" v0 v1 ...  (libraries under breseq_b200/build/variants/)
ARGS=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  BRQ_LIB_PATH=$PWD/breseq_b200/build/variants/libbrq_$v.so python bench.py --no-cpu --no-bam $ARGS > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/ab_%s.json" % v) if l.startswith("{")][-1])
    c = d["config"]
    print(v, c["config"], c["genome_scale"], "step %.3f" % d["ms_per_step"], {k: round(x, 3) for k, x in c["kernel_ms_rank0"].items()})
except Exception as e:
    print(v, "failed", e, open("gpurun_out/ab_%s.err" % v).read()[-500:])
PY
done
