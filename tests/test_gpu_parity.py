"""Parity of the CUDA path against the CPU oracle, through the C ABI (needs a B200).

Bars (BASELINE.json north_star): error counts and coverage bit-exact; identical set of RA calls and
consensus/polymorphism decisions; log-likelihood scores within 1e-9 relative; frequencies within 1e-6.
"""
import filecmp
import os

import numpy as np
import pytest

import breseq_b200 as bq
import helpers

pytestmark = pytest.mark.gpu

REL = 1e-9


# "bounded" is the product default: the EM fit runs only where the presence bound cannot rule an RA row out.
# "fit_all" forces the fit on every column so that every per-column diagnostic can be compared with the oracle.
@pytest.fixture(scope="module", params=[(n, m) for n in helpers.DATASETS for m in ("bounded", "fit_all")],
                ids=lambda p: "%s-%s" % p)
def run(request, datasets, tmp_path_factory):
    name, mode = request.param
    d = datasets[name]
    out = str(tmp_path_factory.mktemp("gpu_%s_%s" % (name, mode)))
    ctx = bq.Context(device=0)
    ctx.stage_bam(d["bam"], d["fasta"], **helpers.stage_kwargs(d))
    ctx.error_count(helpers.covariates(d))
    counts, cov = ctx.hist_download()
    ctx.derive_error_table()
    ctx.write_error_count_files(out, os.path.join(out, "error_rates.tab"), helpers.readfile_names(d))
    params = bq.Context.score_params(d["mutation_cutoff"], d["polymorphism_cutoff"], d["precision"], d["places"],
                                     fit_all_columns=(mode == "fit_all"))
    ctx.score_columns(params)
    cols, flagged = ctx.columns_download()
    n = len(d["contig_lens"])
    gd = os.path.join(out, "ra_mc_evidence.gd")
    stats = ctx.write_evidence(gd, [d["del_prop"]] * n, [d["del_seed"]] * n)
    o = helpers.oracle_columns(d["oracle_columns"])
    slot = helpers.oracle_slots(o, ctx.stream(), helpers.visit_slot0(helpers.contig_names(d), d["contig_lens"]))
    yield dict(d=d, out=out, counts=counts, cov=cov, cols=cols, flagged=flagged, gd=gd, stats=stats, o=o, g=cols[slot], ctx=ctx,
               mode=mode, params=params)
    ctx.close()


def test_error_counts_bit_exact(run):
    assert np.array_equal(run["counts"].astype(np.int64), helpers.oracle_counts(run["d"]["oracle_counts"]))


def test_error_rate_files_byte_identical(run):
    d = run["d"]
    assert filecmp.cmp(os.path.join(run["out"], "error_rates.tab"), d["oracle_rates"], shallow=False)
    for rf in helpers.readfile_names(d):
        name = "base_qual_error_prob.%s.tab" % rf
        assert filecmp.cmp(os.path.join(run["out"], name), os.path.join(d["oracle_dir"], name), shallow=False), name
    for g in range(len(d["contig_lens"])):
        name = "%d.unique_only_coverage_distribution.tab" % g
        assert filecmp.cmp(os.path.join(run["out"], name), os.path.join(d["oracle_dir"], name), shallow=False), name


def test_coverage_fit_from_the_device_histogram(run):
    """the fit between the passes takes the histogram pass 1 left in HBM: same doubles as from the file the reference reads back"""
    if run["mode"] != "bounded":
        pytest.skip("once per dataset")
    d, ctx = run["d"], run["ctx"]
    host = bq.Context(device=-1)
    for g in range(len(d["contig_lens"])):
        path = os.path.join(run["out"], "%d.unique_only_coverage_distribution.tab" % g)
        if not os.path.exists(path):
            continue
        for pr in (0.05 / np.sqrt(float(sum(d["contig_lens"]))), 0.01):
            a, b = ctx.fit_coverage_distribution(g, pr), host.fit_coverage_file(path, pr)
            assert {k: float(v).hex() for k, v in a.items()} == {k: float(v).hex() for k, v in b.items()}, (g, pr, a, b)


def test_coverage_bit_exact(run):
    g, o = run["g"], run["o"]
    assert np.array_equal(g["unique"], o["unique"].astype(np.uint32))
    assert np.array_equal(g["raw_redundant"], o["raw_redundant"].astype(np.uint32))
    assert np.array_equal(g["redundant"], o["redundant"])
    assert np.array_equal(g["n"], o["n"])
    assert np.array_equal(((g["bits"] & bq.CO_UNIQUE_ONLY) != 0), o["unique_only"].astype(bool))


def test_log_likelihoods_and_scores(run):
    g, o = run["g"], run["o"]
    m = o["n"] > 0
    rel = np.abs(g["ll"][m] - o["ll"][m]) / np.maximum(np.abs(o["ll"][m]), 1e-300)
    assert rel.max() < REL
    # scores are differences of O(100) sums: 1e-9 relative to the magnitudes that were subtracted
    scale = np.maximum(np.abs(o["ll"][m]).max(axis=1), 1.0)
    assert (np.abs(g["consensus_score"][m] - o["consensus_score"][m]) / scale).max() < REL
    assert np.all(np.isnan(g["consensus_score"][~m])) and np.all(np.isnan(o["consensus_score"][~m]))
    fit = (g["bits"] & bq.CO_FIT) != 0
    v = ~np.isnan(o["variant_score"])
    if run["mode"] == "fit_all":
        assert np.array_equal(fit, m)
    else:
        # a column the kernel did not fit has no scoring record, or provably cannot emit an RA row
        assert not np.any(fit & ~m)
        skipped = m & ~fit
        assert skipped.sum() > 0.5 * m.sum(), "the presence bound should settle most columns"
        assert not np.any(o["emitted"][skipped])
        vs = o["variant_score"][skipped & v]
        assert vs.size == 0 or vs.max() < run["d"]["polymorphism_cutoff"]
        assert np.all(np.isnan(g["variant_score"][~fit]))
    assert np.array_equal(v[fit], ~np.isnan(g["variant_score"][fit]))
    v &= fit
    scale_v = np.maximum(np.abs(o["log10_likelihood"][v]), 1.0)
    assert (np.abs(g["variant_score"][v] - o["variant_score"][v]) / scale_v).max() < 1e-7  # the EM stops at |df| < 1e-6


def test_calls_and_decisions_identical(run):
    g, o = run["g"], run["o"]
    bits = g["bits"]
    fit = (bits & bq.CO_FIT) != 0
    assert np.array_equal(bits & 7, o["best"])
    for name, shift in (("major", 3), ("minor", 6), ("variant", 9)):
        assert np.array_equal(((bits >> shift) & 7)[fit], o[name][fit]), name
        assert np.all(((bits >> shift) & 7)[~fit] == 5), name
    recheck = (bits & bq.CO_RECHECK) != 0
    pred = (bits & bq.CO_BASE_PREDICTED) != 0
    assert np.array_equal(pred[~recheck], o["base_predicted"][~recheck].astype(bool))
    emit = (bits & bq.CO_EMIT) != 0
    assert np.all(emit[o["emitted"] == 1]), "an oracle RA call was not flagged by the kernel"
    # EM iteration counts are reported for the slots that needed a fit; they follow the reference's
    assert np.array_equal(((bits >> 16) & 0xFF) > 0, fit)
    assert np.mean(((bits >> 16) & 0xFF)[fit] == o["iterations"][fit]) > 0.999


def test_genome_diff_identical(run):
    d = run["d"]
    mine, ora = helpers.parse_gd(run["gd"]), helpers.parse_gd(d["oracle_gd"])
    assert [(r["type"], r["spec"]) for r in mine] == [(r["type"], r["spec"]) for r in ora]
    for a, b in zip(mine, ora):
        assert a["id"] == b["id"]
        assert sorted(a["kv"]) == sorted(b["kv"])
        for k in a["kv"]:
            if a["kv"][k] == b["kv"][k]:
                continue
            if k in ("frequency", "major_frequency", "frequency_lower", "frequency_upper"):
                assert abs(float(a["kv"][k]) - float(b["kv"][k])) < 1e-6, (k, a["kv"][k], b["kv"][k])
            else:
                raise AssertionError("field %s differs: %s vs %s" % (k, a["kv"][k], b["kv"][k]))
    assert open(run["gd"]).read() == open(d["oracle_gd"]).read()


def test_entry_point_adapters(run, tmp_path):
    """breseq::error_count / identify_mutations argument lists, files on disk."""
    d = run["d"]
    out = str(tmp_path)
    rates = os.path.join(out, "error_rates.tab")
    bq.error_count(d["bam"], d["fasta"], out, helpers.readfile_names(d), True, True, False, 3, helpers.covariates(d),
                   read_file_sets=helpers.read_file_sets(d), error_rates_file_name=rates)
    assert filecmp.cmp(rates, d["oracle_rates"], shallow=False)
    n = len(d["contig_lens"])
    gd = os.path.join(out, "ra_mc_evidence.gd")
    bq.identify_mutations(d["bam"], d["fasta"], gd, [d["del_prop"]] * n, [d["del_seed"]] * n, d["mutation_cutoff"],
                          d["polymorphism_cutoff"], d["precision"], d["places"], False, error_rates_file_name=rates,
                          read_file_sets=helpers.read_file_sets(d))
    assert open(gd).read() == open(d["oracle_gd"]).read()


def test_idempotent_and_order_free(run):
    """Re-running the kernels on the resident stream gives bit-identical integer results."""
    ctx, d = run["ctx"], run["d"]
    ctx.error_count(helpers.covariates(d))
    counts, cov = ctx.hist_download()
    assert np.array_equal(counts, run["counts"]) and np.array_equal(cov, run["cov"])
    ctx.derive_error_table()
    ctx.score_columns(run["params"])
    cols, _ = ctx.columns_download()
    for f in ("unique", "raw_redundant", "n", "redundant"):
        assert np.array_equal(cols[f], run["cols"][f])
    # checksum property: scoring depth summed over slots == eligible records in the stream
    t = helpers.emulate_tally(ctx.stream())
    assert cols["n"].sum() == t["n"].sum() and cols["unique"].sum() == t["unique"].sum()


def test_read_pos_and_base_repeat_covariates(datasets, tmp_path):
    """Pass 1 with the optional covariates (8-byte histogram records, global-memory table): counts bit-exact."""
    d = datasets["multi"]
    cov = "read_set=3,obs_base,ref_base,quality=42,read_pos=150,base_repeat=5"
    out = str(tmp_path)
    ec, _ = helpers.cli_args(d, out)
    ec[ec.index("--covariates") + 1] = cov
    dump = os.path.join(out, "counts.tab")
    helpers.run_oracle(*ec, "--counts-dump", dump)
    ctx = bq.Context(device=0)
    ctx.stage_bam(d["bam"], d["fasta"], read_file_sets=helpers.read_file_sets(d), use_read_pos=True, use_base_repeat=True)
    assert ctx.stream()["hist_rec"].dtype == np.uint64
    ctx.error_count(cov)
    counts, _ = ctx.hist_download()
    assert np.array_equal(counts.astype(np.int64), helpers.oracle_counts(dump))
    # a stream staged without them refuses those covariates instead of counting wrongly
    ctx.stage_bam(d["bam"], d["fasta"], read_file_sets=helpers.read_file_sets(d))
    assert ctx.stream()["hist_rec"].dtype == np.uint32
    with pytest.raises(bq.BrqError):
        ctx.error_count(cov)
    ctx.close()


def test_compact_histogram_stream_with_other_covariates(datasets, tmp_path):
    """The 16-bit histogram records (csrc/brq_types.h) under covariate strings other than the default: counts bit-exact."""
    d = datasets["multi"]
    ctx = bq.Context(device=0)
    ctx.stage_bam(d["bam"], d["fasta"], read_file_sets=helpers.read_file_sets(d))
    s = ctx.stream()
    assert s["hist16"] is not None and len(s["hist16"]) > 0.9 * len(s["hist_rec"])
    # a quality axis of 64: the joint histogram no longer fits beside the table, every record takes the generic path
    cov = "read_set=3,obs_base,ref_base,quality=64"
    out = str(tmp_path)
    ec, _ = helpers.cli_args(d, out)
    ec[ec.index("--covariates") + 1] = cov
    dump = os.path.join(out, "counts.tab")
    helpers.run_oracle(*ec, "--counts-dump", dump)
    ctx.error_count(cov)
    counts, _ = ctx.hist_download()
    assert np.array_equal(counts.astype(np.int64), helpers.oracle_counts(dump))
    # no read_set covariate: one set plane in the joint histogram whatever the records' read sets are; the counts are
    # the default table summed over the read sets (index = read_set + 3 * (ref + 5 * obs + 25 * quality))
    ctx.error_count("obs_base,ref_base,quality=42")
    counts, _ = ctx.hist_download()
    full = helpers.oracle_counts(d["oracle_counts"])
    assert np.array_equal(counts.astype(np.int64), full.reshape(-1, 3).sum(axis=1))
    ctx.close()


class _DevArray:
    """A device buffer of the library as a CUDA array (int64) torch can wrap."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}


@pytest.mark.parametrize("name,world", [("lambda", 2), ("multi", 3), ("tiny", 2), ("ltee", 2)])
def test_sharded_run_merges_to_the_unsharded_evidence(name, world, datasets, tmp_path):
    """A run sharded by reference range (one context per shard, as one per GPU): the histograms are summed (the path's one
    collective), every shard scores its range, and the shards' evidence shares walked together give the unsharded
    ra_mc_evidence.gd byte for byte: MC / UN intervals and row ids cross shard boundaries."""
    import torch
    d = datasets[name]
    n = len(d["contig_lens"])
    prop, seed = [d["del_prop"]] * n, [d["del_seed"]] * n
    ctxs = [bq.Context(device=0) for _ in range(world)]
    try:
        for r, c in enumerate(ctxs):
            c.stage_bam(d["bam"], d["fasta"], shard=(r, world), **helpers.stage_kwargs(d))
            c.error_count(helpers.covariates(d))
            c.sync()
        parts = [torch.as_tensor(_DevArray(*c.hist_device()[:2]), device="cuda:0") for c in ctxs]
        total = torch.stack(parts).sum(dim=0)
        for p in parts:
            p.copy_(total)
        torch.cuda.synchronize()
        counts, _ = ctxs[-1].hist_download()
        assert np.array_equal(counts.astype(np.int64), helpers.oracle_counts(d["oracle_counts"]))
        params = bq.Context.score_params(d["mutation_cutoff"], d["polymorphism_cutoff"], d["precision"], d["places"])
        shares = []
        for c in ctxs:
            c.derive_error_table()
            c.score_columns(params)
            shares.append(c.evidence_export(prop))
        gd = str(tmp_path / "merged.gd")
        k = ctxs[0].write_evidence_merged(gd, shares[::-1], prop, seed)
        assert open(gd).read() == open(d["oracle_gd"]).read()
        assert k["RA"] + k["MC"] + k["UN"] == len(helpers.parse_gd(gd))
    finally:
        for c in ctxs:
            c.close()


def test_one_context_two_bams_through_the_adapter(datasets, tmp_path):
    """A context reused for a second BAM through brq_run_error_count (same covariates) must write the second BAM's files:
    the host copies of the histograms belong to the stream they were downloaded from."""
    a, b = datasets["tiny"], datasets["multi"]
    ctx = bq.Context(device=0)
    for d, sub in ((a, "a"), (b, "b")):
        out = str(tmp_path / sub)
        os.makedirs(out)
        cov = "read_set=3,obs_base,ref_base,quality=42"
        bq.error_count(d["bam"], d["fasta"], out, helpers.readfile_names(d), covariates=cov, read_file_sets=helpers.read_file_sets(d),
                       error_rates_file_name=os.path.join(out, "error_rates.tab"), ctx=ctx)
        assert helpers.covariates(d) == cov
        assert filecmp.cmp(os.path.join(out, "error_rates.tab"), d["oracle_rates"], shallow=False)
        for rf in helpers.readfile_names(d):
            name = "base_qual_error_prob.%s.tab" % rf
            assert filecmp.cmp(os.path.join(out, name), os.path.join(d["oracle_dir"], name), shallow=False), name
        for g in range(len(d["contig_lens"])):
            name = "%d.unique_only_coverage_distribution.tab" % g
            assert filecmp.cmp(os.path.join(out, name), os.path.join(d["oracle_dir"], name), shallow=False), name
    ctx.close()


def test_base_quality_cutoff_zero_is_a_value(datasets, tmp_path):
    """Settings::base_quality_cutoff = 0 (every quality scores) runs and matches the oracle run with the same cutoff."""
    d = datasets["tiny"]
    out = str(tmp_path)
    _, im = helpers.cli_args(d, out, rates=d["oracle_rates"], gd=os.path.join(out, "o.gd"))
    helpers.run_oracle(*im, "--base-quality-cutoff", "0")
    n = len(d["contig_lens"])
    bq.identify_mutations(d["bam"], d["fasta"], os.path.join(out, "p.gd"), [d["del_prop"]] * n, [d["del_seed"]] * n, d["mutation_cutoff"],
                          d["polymorphism_cutoff"], d["precision"], d["places"], error_rates_file_name=d["oracle_rates"],
                          read_file_sets=helpers.read_file_sets(d), base_quality_cutoff=0)
    assert open(os.path.join(out, "p.gd")).read() == open(os.path.join(out, "o.gd")).read()


# ---- how far the device's arithmetic is from the reference's, and what the decisions' slack has to cover
FIT_SLACK = 1e-6       # score_slots.cu: a fitted slot emits when variant_score >= cutoff - slack; emitting slots are re-evaluated on the host
SCREEN_MARGIN = 0.25   # score_slots.cu SCREEN_MARGIN


@pytest.mark.parametrize("name", ["deep", "lambda", "pop1000"])
def test_fit_parity_margins(name, datasets):
    """Every column fitted on the device (BRQ_SCORE_FIT_ALL_COLUMNS) against the oracle's fit: ABSOLUTE differences of the
    presence score and the frequencies, the columns whose EM stopped at another iteration listed one by one.  The slack of
    the emit decision must be ten times what is measured here; a column is never lost to rounding."""
    d = datasets[name]
    ctx = bq.Context(device=0)
    ctx.stage_bam(d["bam"], d["fasta"], **helpers.stage_kwargs(d))
    ctx.load_error_table(d["oracle_rates"])
    ctx.score_columns(bq.Context.score_params(d["mutation_cutoff"], d["polymorphism_cutoff"], d["precision"], d["places"], fit_all_columns=True))
    cols, _ = ctx.columns_download()
    o = helpers.oracle_columns(d["oracle_columns"])
    g = cols[helpers.oracle_slots(o, ctx.stream(), helpers.visit_slot0(helpers.contig_names(d), d["contig_lens"]))]
    ctx.close()
    fit = ((g["bits"] & bq.CO_FIT) != 0) & ~np.isnan(o["variant_score"]) & ~np.isnan(g["variant_score"])
    dv = np.abs(g["variant_score"][fit] - o["variant_score"][fit])
    it_g, it_o = ((g["bits"] >> 16) & 0xFF)[fit], o["iterations"][fit]
    same = it_g == it_o
    idx = np.flatnonzero(fit)
    print("\n%s: %d fitted columns with a variant; |d variant_score| max %.3g (same iteration count: %.3g); %d columns stop at another iteration"
          % (name, fit.sum(), dv.max(), dv[same].max(), (~same).sum()))
    for k in np.flatnonzero(~same)[:50]:
        i = idx[k]
        print("  tid %d pos %d ins %d: iterations device %d oracle %d, variant_score device %.12g oracle %.12g (cutoff %g)"
              % (o["tid"][i], o["pos1"][i], o["insert_count"][i], it_g[k], it_o[k], g["variant_score"][i], o["variant_score"][i], d["polymorphism_cutoff"]))
    assert dv[same].max() <= 1e-8, "device fit differs from the reference's beyond 1e-8 absolute"
    assert dv.max() <= FIT_SLACK / 10, "the emit slack (1e-6) must be ten times the largest difference"
    # a column that stops at another iteration is either far from the cutoff or flagged for the host's verdict
    near = np.abs(o["variant_score"][fit] - d["polymorphism_cutoff"]) < 10 * FIT_SLACK
    flagged = ((g["bits"] & (bq.CO_EMIT | bq.CO_RECHECK)) != 0)[fit]
    assert np.all(flagged[near & ~same])


@pytest.mark.parametrize("name", ["deep", "lambda", "pop1000", "multi"])
def test_presence_bounds_hold_on_every_column(name, datasets):
    """BRQ_SCORE_KEEP_BOUNDS: a slot the kernels did not fit reports the upper bound that settled it (the tally's, or the
    screen kernel's); it must be at or above the reference's presence score on EVERY such column, and under the cutoff by
    the decision's margin."""
    d = datasets[name]
    ctx = bq.Context(device=0)
    ctx.stage_bam(d["bam"], d["fasta"], **helpers.stage_kwargs(d))
    ctx.load_error_table(d["oracle_rates"])
    ctx.score_columns(bq.Context.score_params(d["mutation_cutoff"], d["polymorphism_cutoff"], d["precision"], d["places"], keep_bounds=True))
    cols, _ = ctx.columns_download()
    o = helpers.oracle_columns(d["oracle_columns"])
    g = cols[helpers.oracle_slots(o, ctx.stream(), helpers.visit_slot0(helpers.contig_names(d), d["contig_lens"]))]
    ctx.close()
    settled = ((g["bits"] & bq.CO_FIT) == 0) & (o["n"] > 0)
    assert settled.sum() > 0.5 * (o["n"] > 0).sum()
    bound = g["variant_score"][settled]
    assert np.all(np.isfinite(bound)) and bound.max() < d["polymorphism_cutoff"] - FIT_SLACK / 2
    v = ~np.isnan(o["variant_score"][settled])
    gap = bound[v] - o["variant_score"][settled][v]
    print("\n%s: %d settled columns, bound - reference score: min %.3g, median %.3g" % (name, settled.sum(), gap.min(), np.median(gap)))
    assert gap.min() >= 0.0, "a bound is below the reference's presence score"
    assert not np.any(o["emitted"][settled])
