import sys, time
sys.path.insert(0, '/root/repo')
import bench, breseq_b200 as bq
ctx = bq.Context(device=0)
spec = bq.SynthSpec(seed=2, read_sets=bench.READ_SETS, contig_lens=[bench.GENOME], contig_prefix="REL606_range0", n_polymorphic=40, n_fixed=10, n_gaps=3)
ctx.stage_synthetic(spec, read_file_sets=spec.read_file_sets())
ctx.upload(); ctx.sync()
params = bq.Context.score_params(bench.MUTATION_CUTOFF, bench.POLYMORPHISM_CUTOFF, bench.PRECISION, bench.PLACES)
for it in range(8):
    t0 = time.perf_counter(); ctx.error_count(bench.COVARIATES); t1 = time.perf_counter(); ctx.derive_error_table(); t2 = time.perf_counter(); ctx.score_columns(params); t3 = time.perf_counter()
    k = ctx.kernel_ms()
    if it >= 3: print("wall ms: error_count %.3f derive %.3f score %.3f | kernels %s" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, {a: round(b, 3) for a, b in k.items()}))
