"""Golden fixtures written by the REFERENCE'S OWN sources (tests/golden/make_golden.py) pin both the
CPU oracle and the CUDA path.

* not gpu: oracle/oracle_cli reproduces every golden file byte for byte; its per-column dump agrees
  with the reference's per-position debug file; the live reference build (when oracle/_ref/ref_cli is
  present, i.e. in the build container) still reproduces the committed files.
* gpu:     the CUDA path, through the C ABI, reproduces every golden file byte for byte.
"""
import filecmp
import hashlib
import os
import shutil

import numpy as np
import pytest

import breseq_b200 as bq
import helpers

NAMES = [n for n in helpers.DATASETS if not helpers.DATASETS[n].get("no_golden")]


def golden(name, f):
    return os.path.join(helpers.GOLDEN, name, f)


def sha256(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def assert_same_files(d, got_dir, want_dir, label):
    for f in helpers.pass_output_names(d):
        assert filecmp.cmp(os.path.join(got_dir, f), os.path.join(want_dir, f), shallow=False), "%s: %s differs" % (label, f)
    if d.get("big_table"):
        want = open(os.path.join(want_dir, "error_rates.tab.sha256")).read().split()[0]
        assert sha256(os.path.join(got_dir, "error_rates.tab")) == want, "%s: error_rates.tab differs" % label


@pytest.mark.parametrize("name", NAMES)
def test_generator_is_deterministic(name, datasets):
    """The committed outputs belong to exactly these inputs (same seed => same BAM, any thread count)."""
    d = datasets[name]
    want = dict(reversed(l.split()) for l in open(golden(name, "inputs.sha256")).read().strip().split("\n"))
    assert sha256(d["bam"]) == want["reference.bam"]
    assert sha256(d["fasta"]) == want["reference.fasta"]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference_files(name, datasets):
    d = datasets[name]
    assert_same_files(d, d["oracle_dir"], os.path.join(helpers.GOLDEN, name), "oracle vs reference golden")


@pytest.mark.parametrize("name", NAMES)
def test_preprocess_stage_matches_reference(name, datasets, tmp_path):
    """error_count(..., preprocess_stage = true), the stage 03 call (breseq_cmdline.cpp:1969): the oracle's and the product's
    no_pos_hash_per_position_pr (error_count.cpp:157-166, 191-194, 217-229) against what the reference build left in its
    Summary, to 17 significant digits; a run sharded by reference range adds up to the same numbers."""
    d = datasets[name]
    want = open(golden(name, helpers.PREPROCESS_TAB)).read()
    assert helpers.run_preprocess(helpers.ORACLE_CLI, d, str(tmp_path / "oracle")) == want
    assert helpers.product_preprocess_tab(d) == want
    assert helpers.product_preprocess_tab(d, shards=3) == want


def tiny_from_fixture(tmp):
    """The committed BAM itself (not a regenerated one)."""
    d = dict(helpers.DATASETS["tiny"])
    for f in ("reference.bam", "reference.fasta", "reference.fasta.fai"):
        shutil.copy(golden("tiny", f), os.path.join(tmp, f))
    d["bam"], d["fasta"] = os.path.join(tmp, "reference.bam"), os.path.join(tmp, "reference.fasta")
    return d


def test_oracle_on_committed_bam(built, tmp_path):
    d = tiny_from_fixture(str(tmp_path))
    out = str(tmp_path / "o")
    os.makedirs(out)
    ec, im = helpers.cli_args(d, out)
    helpers.run_oracle(*ec)
    helpers.run_oracle(*im, "--columns-out", os.path.join(out, "columns.bin"))
    assert_same_files(d, out, os.path.join(helpers.GOLDEN, "tiny"), "oracle on committed BAM")

    # per-column: the reference's per-position debug file (identify_mutations.cpp:1693-1733) prints,
    # for every (column, insert_count), the consensus score at 6 significant digits and the scoring
    # records per base and strand
    o = helpers.oracle_columns(os.path.join(out, "columns.bin"))
    rows = [l.split() for l in open(golden("tiny", "per_position_file.tab")) if l.strip()]
    assert len(rows) == len(o)
    order = np.lexsort((o["insert_count"], o["pos1"], np.argsort(np.argsort(helpers.contig_names(d)))[o["tid"]]))
    for row, c in zip(rows, o[order]):
        assert int(row[0]) == c["pos1"] and int(row[1]) == c["insert_count"]
        assert row[2] == "ACGT.N"[c["ref"]]
        want = "nan" if np.isnan(c["consensus_score"]) else "%g" % c["consensus_score"]
        assert row[3].lstrip("-") == want.lstrip("-") if want == "nan" else row[3] == want, (row[:4], want)
        per_base = [tuple(int(x) for x in row[5 + 2 * j].strip("()").split("/")) for j in range(5)]
        assert sum(b for b, t in per_base) + sum(t for b, t in per_base) == c["n"]


@pytest.mark.skipif(not os.path.exists(helpers.REF_CLI), reason="oracle/_ref/ref_cli is only built where /root/reference exists")
@pytest.mark.parametrize("name", NAMES)
def test_live_reference_matches_committed_golden(name, datasets, tmp_path):
    d = datasets[name]
    out = str(tmp_path / "ref")
    helpers.run_reference(d, out, per_position=False)
    assert_same_files(d, out, os.path.join(helpers.GOLDEN, name), "live reference vs committed golden")


def run_cuda(d, out, optional_outputs=False):
    os.makedirs(out, exist_ok=True)
    rates = os.path.join(out, "error_rates.tab")
    bq.error_count(d["bam"], d["fasta"], out, helpers.readfile_names(d), True, True, False, 3, helpers.covariates(d),
                   read_file_sets=helpers.read_file_sets(d), error_rates_file_name=rates)
    n = len(d["contig_lens"])
    bq.identify_mutations(d["bam"], d["fasta"], os.path.join(out, "ra_mc_evidence.gd"), [d["del_prop"]] * n, [d["del_seed"]] * n,
                          d["mutation_cutoff"], d["polymorphism_cutoff"], d["precision"], d["places"], optional_outputs,
                          error_rates_file_name=rates, read_file_sets=helpers.read_file_sets(d),
                          per_position_file_name=os.path.join(out, "per_position_file.tab") if optional_outputs else None,
                          coverage_tsv_pattern=os.path.join(out, "@.coverage.tsv") if optional_outputs else None)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_reproduces_reference_files(name, datasets, tmp_path):
    d = datasets[name]
    out = str(tmp_path / "cuda")
    run_cuda(d, out)
    assert_same_files(d, out, os.path.join(helpers.GOLDEN, name), "CUDA path vs reference golden")


@pytest.mark.gpu
def test_cuda_path_on_committed_bam(built, tmp_path):
    d = tiny_from_fixture(str(tmp_path))
    out = str(tmp_path / "cuda")
    run_cuda(d, out, optional_outputs=True)
    assert_same_files(d, out, os.path.join(helpers.GOLDEN, "tiny"), "CUDA path on committed BAM")
    # the per-position debug file (identify_mutations.cpp:1693-1733), byte for byte; insert sub-columns included
    assert filecmp.cmp(os.path.join(out, "per_position_file.tab"), golden("tiny", "per_position_file.tab"), shallow=False)
    # <seq>.coverage.tsv of a BAM with two read groups: the three aggregate columns and the same three per read group
    # (identify_mutations.cpp:858-862, 2046-2050)
    for c in helpers.contig_names(d):
        assert filecmp.cmp(os.path.join(out, c + ".coverage.tsv"), golden("tiny", c + ".coverage.tsv"), shallow=False), c


@pytest.mark.gpu
def test_cuda_optional_outputs_match_reference(datasets, tmp_path):
    """print_per_position_file and <seq>.coverage.tsv (--predict-copy-number) on the deep amplicon: what the reference wrote."""
    d = datasets["deep"]
    out = str(tmp_path / "cuda")
    run_cuda(d, out, optional_outputs=True)
    assert filecmp.cmp(os.path.join(out, "per_position_file.tab"), golden("deep", "per_position_file.tab"), shallow=False)
    tsv = helpers.contig_names(d)[0] + ".coverage.tsv"
    assert filecmp.cmp(os.path.join(out, tsv), golden("deep", tsv), shallow=False)


# ---- user evidence (Settings::user_evidence_genome_diff_file_name; identify_mutations.cpp:879, 1013-1020, 1346-1355, 1914-2019)
USER_GD = os.path.join(helpers.GOLDEN, "lambda", "user_evidence.gd")
USER_GOLDEN = os.path.join(helpers.GOLDEN, "lambda", "ra_mc_evidence.user_evidence.gd")


def test_oracle_reports_user_evidence_like_the_reference(datasets, tmp_path):
    """Eight user rows on the lambda dataset: one the data already reports (it only gains user_defined=1), absent alleles,
    forced insert sub-columns (insert_position 2 and 3 where no read has an insertion that long), a reference base of N."""
    d = datasets["lambda"]
    out = str(tmp_path)
    _, im = helpers.cli_args(d, out, rates=d["oracle_rates"], gd=os.path.join(out, "o.gd"))
    helpers.run_oracle(*im, "--user-evidence", USER_GD)
    assert open(os.path.join(out, "o.gd")).read() == open(USER_GOLDEN).read()


def test_user_evidence_forces_insert_sub_columns_in_host_staging(datasets):
    d = datasets["lambda"]
    plain, forced = bq.Context(device=-1), bq.Context(device=-1)
    plain.stage_bam(d["bam"], d["fasta"], **helpers.stage_kwargs(d))
    forced.stage_bam(d["bam"], d["fasta"], user_evidence_gd=USER_GD, **helpers.stage_kwargs(d))
    a, b = plain.stream(), forced.stream()
    have = {(int(p), int(k)) for p, k in zip(a["ins_parent"], a["ins_count"])}
    want = {(int(p), int(k)) for p, k in zip(b["ins_parent"], b["ins_count"])}
    assert have < want
    assert {(4999, 1), (4999, 2), (15183, 2), (15183, 3), (29999, 1)} <= want - have | have   # 0-based parent slot, level
    # a forced sub-column holds a '.' observation of every read that spans the column
    for p, k in want - have:
        s = int(b["n_base"]) + [i for i, (q, j) in enumerate(zip(b["ins_parent"], b["ins_count"])) if (int(q), int(j)) == (p, k)][0]
        assert b["score_cnt"][s] == b["score_cnt"][p] or b["score_cnt"][s] > 0
    plain.close()
    forced.close()


@pytest.mark.gpu
@pytest.mark.parametrize("staging", ["device", "host"])
def test_cuda_path_reports_user_evidence_like_the_reference(staging, datasets, tmp_path):
    d = datasets["lambda"]
    out = str(tmp_path)
    ctx = bq.Context(device=0)
    ctx.stage_bam(d["bam"], d["fasta"], user_evidence_gd=USER_GD, staging=staging, **helpers.stage_kwargs(d))
    ctx.load_error_table(os.path.join(helpers.GOLDEN, "lambda", "error_rates.tab"))
    ctx.score_columns(bq.Context.score_params(d["mutation_cutoff"], d["polymorphism_cutoff"], d["precision"], d["places"]))
    ctx.write_evidence(os.path.join(out, "ra_mc_evidence.gd"), [d["del_prop"]], [d["del_seed"]])
    assert open(os.path.join(out, "ra_mc_evidence.gd")).read() == open(USER_GOLDEN).read()
    ctx.close()
    # the one-call adapter, through the reference's own option name
    bq.identify_mutations(d["bam"], d["fasta"], os.path.join(out, "adapter.gd"), [d["del_prop"]], [d["del_seed"]], d["mutation_cutoff"],
                          d["polymorphism_cutoff"], d["precision"], d["places"], error_rates_file_name=os.path.join(helpers.GOLDEN, "lambda", "error_rates.tab"),
                          read_file_sets=helpers.read_file_sets(d), user_evidence_genome_diff_file_name=USER_GD)
    assert open(os.path.join(out, "adapter.gd")).read() == open(USER_GOLDEN).read()


@pytest.mark.gpu
def test_user_evidence_across_targets_and_shards(datasets, tmp_path):
    """Several targets: entries consumed target by target in visit order, an entry for a position no column has (it blocks
    the rest of the list, as in the reference), and the same run cut into three shards."""
    d = datasets["multi"]
    names = helpers.contig_names(d)
    user = str(tmp_path / "user.gd")
    with open(user, "w") as f:
        f.write("#=GENOME_DIFF\t1.0\n")
        rows = [(names[0], 100, 0, "A", "C"), (names[0], 100, 1, ".", "G"), (names[0], 3000, 0, "T", "G"), (names[1], 50, 0, "C", "T"),
                (names[1], 3499, 2, ".", "A"), (names[1], 999999, 0, "A", "C"), (names[2], 10, 0, "A", "C")]
        for i, r in enumerate(rows):
            f.write("RA\t%d\t.\t%s\t%d\t%d\t%s\t%s\n" % ((i + 1,) + r))
    n = len(names)
    out = str(tmp_path)
    _, im = helpers.cli_args(d, out, rates=d["oracle_rates"], gd=os.path.join(out, "o.gd"))
    helpers.run_oracle(*im, "--user-evidence", user)
    want = open(os.path.join(out, "o.gd")).read()
    assert want.count("user_defined=1") == 5   # the entry past its target's end is never consumed and blocks the one behind it
    params = bq.Context.score_params(d["mutation_cutoff"], d["polymorphism_cutoff"], d["precision"], d["places"])
    ctx = bq.Context(device=0)
    ctx.stage_bam(d["bam"], d["fasta"], user_evidence_gd=user, **helpers.stage_kwargs(d))
    ctx.load_error_table(d["oracle_rates"])
    ctx.score_columns(params)
    ctx.write_evidence(os.path.join(out, "one.gd"), [d["del_prop"]] * n, [d["del_seed"]] * n)
    assert open(os.path.join(out, "one.gd")).read() == want
    ctx.close()
    shares = []
    for rank in range(3):
        c = bq.Context(device=0)
        c.stage_bam(d["bam"], d["fasta"], user_evidence_gd=user, shard=(rank, 3), **helpers.stage_kwargs(d))
        c.load_error_table(d["oracle_rates"])
        c.score_columns(params)
        shares.append(c.evidence_export([d["del_prop"]] * n))
        c.close()
    m = bq.Context(device=-1)
    m.write_evidence_merged(os.path.join(out, "merged.gd"), shares, [d["del_prop"]] * n, [d["del_seed"]] * n)
    m.close()
    assert open(os.path.join(out, "merged.gd")).read() == want


# ---- covariates with ref_pos: the per-position count table (error_count.cpp:105-111, 193-198, 803-846)
def _sha(path):
    import hashlib
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


@pytest.mark.parametrize("name", ["tiny", "ltee"])
def test_per_position_count_table_matches_reference(name, datasets, tmp_path):
    """`ref_pos` in the covariate string: every position's non-empty bins, counted on the host from the staged histogram records
    (ltee: read_pos and base_repeat covariates, eight-byte records); the reference's file is too long to commit, its checksum is"""
    d = datasets[name]
    want = open(golden(name, "error_counts.per_position.sha256")).read().split()[0]
    ctx = bq.Context(device=-1)
    ctx.stage_bam(d["bam"], d["fasta"], **helpers.stage_kwargs(d))
    out = str(tmp_path / "error_counts.tab")
    ctx.write_per_position_counts("ref_pos," + helpers.covariates(d), out)
    assert _sha(out) == want
    ctx.close()


@pytest.mark.gpu
def test_error_count_with_ref_pos_writes_the_per_position_table(datasets, tmp_path):
    """through the entry point: brq_run_error_count with ref_pos writes error_counts.tab (from the device-built stream) and the
    coverage distributions, and no error rates, as the reference does"""
    d = datasets["tiny"]
    out = str(tmp_path / "cuda")
    os.makedirs(out)
    bq.error_count(d["bam"], d["fasta"], out, helpers.readfile_names(d), True, True, False, 3, "ref_pos," + helpers.covariates(d),
                   read_file_sets=helpers.read_file_sets(d), error_rates_file_name=os.path.join(out, "error_rates.tab"))
    assert _sha(os.path.join(out, "error_counts.tab")) == open(golden("tiny", "error_counts.per_position.sha256")).read().split()[0]
    assert not os.path.exists(os.path.join(out, "error_rates.tab"))
    for g in range(len(d["contig_lens"])):
        nm = "%d.unique_only_coverage_distribution.tab" % g
        assert filecmp.cmp(os.path.join(out, nm), golden("tiny", nm), shallow=False)
