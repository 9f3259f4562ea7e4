"""Work-list statistics of the scoring pass on a C2-shaped stream: python scripts/fit_stats.py <scale> <coverage> <polymorphism cutoff>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, breseq_b200 as bq
scale, cov, pc = float(sys.argv[1]), float(sys.argv[2]), float(sys.argv[3])
rs = [dict(bench.READ_SETS[0])]; rs[0]["coverage"] = cov
ctx = bq.Context(device=0)
spec = bq.SynthSpec(seed=2, read_sets=rs, contig_lens=[int(bench.GENOME * scale)], contig_prefix="REL606", n_polymorphic=40, n_fixed=10, n_gaps=3)
ctx.stage_synthetic(spec, read_file_sets=spec.read_file_sets())
ctx.error_count(bench.COVARIATES); ctx.derive_error_table()
p = bq.Context.score_params(10.0, pc, 1e-6, 8)
for _ in range(3):
    ctx.score_columns(p)
print("kernel ms", ctx.kernel_ms())
gd = "/tmp/fit_stats.gd"; print("evidence", ctx.write_evidence(gd, [0.3 * cov], [0.0]))
cols, flagged = ctx.columns_download()
fit = (cols["bits"] & bq.CO_FIT) != 0
it = (cols["bits"] >> 16) & 0xFF
n = cols["n"]
print("slots %d, with records %d, fitted %d (%.1f %%), flagged %d, emit %d" % (len(cols), (n > 0).sum(), fit.sum(), 100.0 * fit.sum() / max(1, (n > 0).sum()), len(flagged), ((cols["bits"] & bq.CO_EMIT) != 0).sum()))
print("EM iterations of the fitted slots: mean %.1f, quantiles 50/90/99/max: %s" % (it[fit].mean(), np.percentile(it[fit], [50, 90, 99, 100])))
vs = cols["variant_score"][fit]
vs = vs[np.isfinite(vs)]
print("variant scores of fitted slots: quantiles 50/90/99/max %s; >= cutoff: %d" % (np.percentile(vs, [50, 90, 99, 100]), (vs >= pc).sum()))
print("depth n quantiles", np.percentile(n[fit], [1, 50, 99]))
ctx.close()
