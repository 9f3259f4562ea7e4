"""Wall time of host staging (synthetic reads of the bench model), host only: python scripts/staging_time.py <scale> <threads>\nBRQ_STAGE_TIMES=1 prints the phases of stage()."""
import sys, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import bench, breseq_b200 as bq
scale = float(sys.argv[1]); th = int(sys.argv[2])
ctx = bq.Context(device=-1, threads=th)
spec = bq.SynthSpec(seed=2, read_sets=bench.READ_SETS, contig_lens=[int(bench.GENOME * scale)], contig_prefix="REL606_range0",
                    n_polymorphic=40, n_fixed=10, n_gaps=3)
t0 = time.perf_counter()
ctx.stage_synthetic(spec, read_file_sets=spec.read_file_sets())
print("threads %d: stage_synthetic %.2f s, records %d" % (th, time.perf_counter() - t0, ctx.stream()["n_score"]))
ctx.close()
