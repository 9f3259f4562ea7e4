"""BAM2COV's table from the CUDA path (brq_write_coverage_table: coverage_tile_kernel over the reads staged in HBM) against the
tables the reference build's coverage_output::table wrote (tests/golden/<name>/coverage_table.<k>.tab): byte for byte."""
import filecmp
import os

import pytest

import breseq_b200 as bq
import helpers
from test_coverage_table import requests

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", [n for n in helpers.DATASETS if not helpers.DATASETS[n].get("no_golden")])
def test_coverage_tables_byte_identical(name, datasets, tmp_path):
    d = datasets[name]
    ctx = bq.Context(device=0)
    ctx.stage_bam(d["bam"], d["fasta"], staging="device", **helpers.stage_kwargs(d))
    for table, region, resolution, total_only, fmt, per_rg in requests(name):
        out = str(tmp_path / table)
        ctx.write_coverage_table(region, out, int(resolution), total_only == "1", fmt == "csv", per_rg == "1")
        assert filecmp.cmp(out, os.path.join(helpers.GOLDEN, name, table), shallow=False), (name, table, region)
    # the walk does not disturb the passes: they still run on the same staged stream
    ctx.error_count(helpers.covariates(d))
    ctx.close()


def test_coverage_table_needs_device_staging_and_a_region_inside_it(datasets, tmp_path):
    d = datasets["multi"]
    names = helpers.contig_names(d)
    ctx = bq.Context(device=0)
    ctx.stage_bam(d["bam"], d["fasta"], staging="host", **helpers.stage_kwargs(d))
    with pytest.raises(bq.BrqError):
        ctx.write_coverage_table(names[0] + ":1-100", str(tmp_path / "a.tab"))
    ctx.stage_bam(d["bam"], d["fasta"], staging="device", seq_ids=[names[0]], **helpers.stage_kwargs(d))
    ctx.write_coverage_table(names[0] + ":1-100", str(tmp_path / "b.tab"))
    with pytest.raises(bq.BrqError):
        ctx.write_coverage_table(names[1] + ":1-100", str(tmp_path / "c.tab"))   # not a visited target
    with pytest.raises(bq.BrqError):
        ctx.write_coverage_table("nosuch:1-100", str(tmp_path / "d.tab"))
    ctx.close()


def test_hand_derived_coverage_rows_on_the_device(tmp_path):
    """the known-answer BAM of test_coverage_table.py (deletion, reference skip, clips, both strands, redundant and unmapped
    reads) through the CUDA path"""
    from test_coverage_table import kat_expected_text, kat_inputs
    bam, fasta = kat_inputs(tmp_path)
    ctx = bq.Context(device=0)
    ctx.stage_bam(bam, fasta, staging="device")
    out = str(tmp_path / "gpu.tab")
    ctx.write_coverage_table("chr:1-14", out)
    assert "".join(l for l in open(out) if not l.startswith("#")) == kat_expected_text()
    ctx.write_coverage_table("chr:1-19", out, total_only=True)   # the tail past the last read: zero rows to the region's end
    rows = [l.split("\t") for l in open(out) if not l.startswith("#")][1:]
    assert [r[0] for r in rows] == [str(i) for i in range(1, 20)] and rows[-1][2:] == ["0", "0", "0\n"]
    ctx.close()


@pytest.mark.parametrize("name", ["multi", "ltee"])
def test_read_group_columns_of_the_coverage_tsv_add_up(name, datasets, tmp_path):
    """<seq>.coverage.tsv of a BAM with several read groups: the per-read-group columns come from the coverage walk, the
    aggregate ones from the tally kernel; two independent counts of the same pileup, so every row has to add up (the golden
    of `tiny` pins the bytes; this covers the larger datasets)."""
    from test_golden import run_cuda
    d = datasets[name]
    out = str(tmp_path / "cuda")
    run_cuda(d, out, optional_outputs=True)
    rows = 0
    for c in helpers.contig_names(d):
        lines = open(os.path.join(out, c + ".coverage.tsv")).read().splitlines()
        head = lines[0].split("\t")
        n_rg = (len(head) - 5) // 3
        assert n_rg >= 2 and head[5] == "RG-0_unique_cov"
        for l in lines[1:]:
            f = l.split("\t")
            unique, redundant = float(f[2]), float(f[3])
            assert unique == sum(float(f[5 + 3 * g]) for g in range(n_rg)), l
            assert abs(redundant - sum(float(f[6 + 3 * g]) for g in range(n_rg))) < 1e-4 * max(1.0, redundant), l
            rows += 1
    assert rows == sum(d["contig_lens"])


def test_reference_suite_tables_on_the_device(tmp_path):
    """the two BAM2COV tables of the reference's own test suite, their BAMs rebuilt from the tables (test_coverage_table.py),
    through the CUDA path: byte for byte"""
    from test_coverage_table import REBUILT_TABLES, REFERENCE_AVERAGE, THINNED_TABLES, rebuilt_inputs, thinned_inputs
    for table, fasta_fixture in REBUILT_TABLES:
        sub = tmp_path / table
        sub.mkdir()
        bam, fasta, region, want = rebuilt_inputs(table, fasta_fixture, sub)
        ctx = bq.Context(device=0)
        ctx.stage_bam(bam, fasta, staging="device")
        out = str(sub / "gpu.tab")
        ctx.write_coverage_table(region, out, 600, False, False, True)
        assert open(out).read() == open(want).read(), table
        ctx.close()
    for table, csv in THINNED_TABLES:   # BAM2COV -a at the default resolution, as a table and as CSV
        sub = tmp_path / table
        sub.mkdir()
        bam, fasta, region, want = thinned_inputs(table, csv, sub)
        ctx = bq.Context(device=0)
        ctx.stage_bam(bam, fasta, staging="device")
        out = str(sub / "gpu.out")
        ctx.write_coverage_table(region, out, 600, False, csv, False, reference_average=REFERENCE_AVERAGE)
        assert open(out).read() == open(want).read(), table
        ctx.close()
