"""Wall time of staging the bench model on the GPU box, host vs device: python scripts/stage_device_time.py <scale> [coverage]
BRQ_STAGE_TIMES=1 prints the phases."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, breseq_b200 as bq
scale = float(sys.argv[1])
rs = [dict(bench.READ_SETS[0])]
if len(sys.argv) > 2:
    rs[0]["coverage"] = float(sys.argv[2])
for staging in ("device", "device", "host"):
    ctx = bq.Context(device=0)
    spec = bq.SynthSpec(seed=2, read_sets=rs, contig_lens=[int(bench.GENOME * scale)], contig_prefix="REL606_range0",
                        n_polymorphic=40, n_fixed=10, n_gaps=3)
    t0 = time.perf_counter()
    ctx.stage_synthetic(spec, read_file_sets=spec.read_file_sets(), staging=staging)
    ctx.sync()
    print("%s staging: stage_synthetic (read synthesis included) %.2f s" % (staging, time.perf_counter() - t0), flush=True)
    ctx.close()
