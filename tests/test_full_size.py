"""BASELINE.json's full-size workload (configs[1]: 4.63 Mb, 100x pe150, 4.6e8 records) through the C ABI on a B200,
checked by size-independent properties: the oracle cannot run at this size in a test, checksums can.

  * histogram: sum of the covariate counts == valid observations in the staged histogram records
  * coverage:  sum over slots of unique / raw_redundant / n == records of each kind in the staged scoring stream
  * unique-only coverage histogram: total == columns without a redundant read, first moment == their records
  * idempotence: a second run over the resident stream gives bit-identical results (all 96 bytes of every slot)
  * scores: every column with scoring records has a finite consensus score and five finite log-likelihood sums whose
    best one names the called base; the planted fixed variants come out as RA rows
"""
import os
import sys

import numpy as np
import pytest

import breseq_b200 as bq

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (the workload definition lives with the benchmark)


@pytest.fixture(scope="module")
def full():
    ctx = bq.Context(device=0)
    spec = bq.SynthSpec(seed=2, read_sets=bench.READ_SETS, contig_lens=[bench.GENOME], contig_prefix="REL606_range0",
                        n_polymorphic=40, n_fixed=10, n_gaps=3)
    ctx.stage_synthetic(spec, read_file_sets=spec.read_file_sets())
    yield ctx, ctx.stream()
    ctx.close()


def test_full_size_checksums(full, tmp_path):
    ctx, s = full
    assert s["n_base"] == bench.GENOME and s["n_score"] > 4.5e8

    # ---- pass 1
    ctx.error_count(bench.COVARIATES)
    counts, cov = ctx.hist_download()
    h = s["hist_rec"]
    valid = int(((h >> 13) & 1).sum(dtype=np.int64) + ((h >> 27) & 1).sum(dtype=np.int64))
    assert int(counts.sum()) == valid
    off = s["hist_off"]
    red = (off[:-1] >> np.uint64(63)).astype(bool)
    depth = np.diff((off & np.uint64((1 << 63) - 1)).astype(np.int64))
    assert int(cov.sum()) == int((~red).sum())
    assert int((cov[0] * np.arange(cov.shape[1])).sum()) == int(depth[~red].sum())
    ctx.derive_error_table()

    # ---- pass 2
    params = bq.Context.score_params(bench.MUTATION_CUTOFF, bench.POLYMORPHISM_CUTOFF, bench.PRECISION, bench.PLACES)
    ctx.score_columns(params)
    cols, flagged = ctx.columns_download()
    cols = cols.copy()
    rec = s["score_rec"]
    g = s["geometry"]
    trash = g["n_st"] * g["n_q"] + 6
    pad = (trash >> 2) * 128 + (trash & 3) * 8
    kind = rec >> 30
    real = rec != pad
    n_red = int((kind == 3).sum(dtype=np.int64))
    n_unique = int(real.sum(dtype=np.int64)) - n_red
    n_scoring = int((((kind == 0) & real) | (kind == 2)).sum(dtype=np.int64))
    assert int(real.sum(dtype=np.int64)) == s["n_score"]
    assert int(cols["unique"].sum(dtype=np.int64)) == n_unique
    assert int(cols["raw_redundant"].sum(dtype=np.int64)) == n_red
    assert int(cols["n"].sum(dtype=np.int64)) == n_scoring
    top = (rec >> 13) & 1
    assert int(cols["unique"][:, 1].sum(dtype=np.int64)) == int((top[real & (kind != 3)]).sum(dtype=np.int64))

    m = cols["n"] > 0
    assert np.all(np.isfinite(cols["ll"][m])) and np.all(np.isfinite(cols["consensus_score"][m]))
    assert np.array_equal(np.argmax(cols["ll"][m], axis=1), (cols["bits"][m] & 7))
    assert np.all(np.isnan(cols["consensus_score"][~m]))

    # ---- idempotence over the resident stream
    ctx.score_columns(params)
    again, flagged2 = ctx.columns_download()
    assert np.array_equal(cols.view(np.uint8), again.view(np.uint8)), "a second run differs"
    assert sorted(flagged.tolist()) == sorted(flagged2.tolist())

    # ---- evidence: the ten planted fixed variants (and the gaps) surface
    gd = str(tmp_path / "ra_mc_evidence.gd")
    k = ctx.write_evidence(gd, [30.0], [0.0])
    assert k["RA"] >= 10 and k["MC"] >= 1
