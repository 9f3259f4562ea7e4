"""The reference's OWN test suite as the pin for the two steps behind pass 2 (tests/golden/make_reference_tests_golden.py took the
rows from /root/reference/tests/<test>/expected.gd, the annotated GenomeDiff each of those breseq runs has to reproduce):

* the RA rows there carry what test_RA_evidence (identify_mutations.cpp:687-749) left in them -- with prediction= and the three
  reject fields removed, brq_test_ra_evidence under the test's settings has to put exactly those back, and drop none of the rows;
* the SNP / DEL / INS / SUB rows whose evidence is RA rows are what predictRAtoSNPorDELorINSorSUB (mutation_predictor.cpp:1955-2211)
  made of them -- brq_predict_ra_mutations has to make the same ones (type, evidence, position, columns; in polymorphism mode
  also frequency= and insert_position=, which a consensus-mode run only adds in later steps).
Host only."""
import json
import os

import pytest

import breseq_b200 as bq
import helpers

DIR = os.path.join(helpers.GOLDEN, "reference_tests")
TESTS = json.load(open(os.path.join(DIR, "tests.json")))
FILTER_FIELDS = ("prediction", "reject", "consensus_reject", "polymorphism_reject")
MUTATION_COLUMNS = {"SNP": 6, "DEL": 6, "INS": 6, "SUB": 7}


@pytest.fixture(scope="module")
def ctx(built):
    c = bq.Context(device=-1)
    yield c
    c.close()


def fasta_for(test, tmp_path):
    t = TESTS[test]
    if not t["fasta"]:   # a filter-only test: one base per sequence the rows name (the filter does not look at it)
        dummy = tmp_path / "dummy.fasta"
        dummy.write_text("".join(">%s\nN\n" % name for name in sorted({r.split("\t")[3] for r in rows_of(test) if r.startswith("RA\t")})))
        return str(dummy)
    path = os.path.join(DIR, t["fasta"])
    if not t["rename"]:
        return path
    text = open(path).read().split("\n", 1)
    renamed = tmp_path / "renamed.fasta"
    renamed.write_text(">%s\n%s" % (t["rename"], text[1]))
    return str(renamed)


def rows_of(test):
    return [line.rstrip("\n") for line in open(os.path.join(DIR, test + ".gd"))]


@pytest.mark.parametrize("test", sorted(TESTS))
def test_filter_puts_back_what_the_reference_run_wrote(ctx, test, tmp_path):
    want = [r for r in rows_of(test) if r.startswith("RA\t")]
    stripped = []
    for r in want:
        c = r.split("\t")
        stripped.append("\t".join(c[:8] + [f for f in c[8:] if f.split("=")[0] not in FILTER_FIELDS]))
    assert stripped != want
    gd = tmp_path / "in.gd"
    gd.write_text("#=GENOME_DIFF\t1.0\n" + "\n".join(stripped) + "\n")
    out = tmp_path / "out.gd"
    t = TESTS[test]
    counts = ctx.test_RA_evidence(str(gd), fasta_for(test, tmp_path), str(out), t["polymorphism_prediction"], **t["settings"])
    got = [line.rstrip("\n") for line in open(out) if line.startswith("RA\t")]
    assert got == want
    assert counts["deleted"] == 0 and counts["rows"] == len(want)


@pytest.mark.parametrize("test", sorted(t for t in TESTS if TESTS[t]["fasta"]))
def test_prediction_makes_the_reference_run_s_mutations(ctx, test, tmp_path):
    rows = rows_of(test)
    t = TESTS[test]
    poly = t["polymorphism_prediction"]
    want = [r.split("\t") for r in rows if r.split("\t")[0] in MUTATION_COLUMNS]
    assert want
    gd = tmp_path / "in.gd"
    gd.write_text("#=GENOME_DIFF\t1.0\n" + "\n".join(r for r in rows if r[:3] in ("RA\t", "MC\t")) + "\n")
    out = tmp_path / "out.gd"
    ctx.predict_ra_mutations(str(gd), fasta_for(test, tmp_path), str(out), poly)
    got = [line.rstrip("\n").split("\t") for line in open(out) if line.split("\t")[0] in MUTATION_COLUMNS]

    def comparable(c):   # without the id: the reference run numbered its mutations after other steps had added theirs
        n = MUTATION_COLUMNS[c[0]]
        return tuple(c[:1] + c[2:n] + (c[n:] if poly else []))
    assert sorted(map(comparable, got)) == sorted(map(comparable, want))
    # the order among them is the GenomeDiff order in both
    assert [comparable(c) for c in got] == [comparable(c) for c in want]


def test_fisher_strand_p_values_of_the_reference_suite(built):
    """fisher_strand_p_value as pass 2's finalisation computes it (finalize.cpp: fisher_2x2; stats.cpp:2144-2171) for every distinct
    pair of strand counts in the RA rows of the reference's test suite: the printed value, digit for digit."""
    rows = [line.split() for line in list(open(os.path.join(DIR, "fisher_kat.tsv")))[1:]]
    assert len(rows) > 2000
    for major, minor, want in rows:
        a, b = map(int, major.split("/"))
        c, d = map(int, minor.split("/"))
        assert "%.5e" % bq.fisher_strand_p_value(c, d, a, b) == want, (major, minor)


# ---- the suite's user-evidence file (tests/lambda_polymorphism_user_evidence/user_evidence.gd: RA rows that duplicate real evidence,
# rows for alleles no read shows, inserted columns up to insert_position 10, a JC row and comment lines in between) on reads
# simulated over the real lambda sequence, C0's read model (se35, ~107x, polymorphism mode)
USER_GD = os.path.join(DIR, "lambda_user_evidence.input.gd")
USER_WANT = os.path.join(DIR, "lambda_user_evidence.ra_mc_evidence.gd")
USER_DATASET = dict(seed=1, contig_lens=[48502], prefix="unused", read_sets=[dict(name="lambda", paired=False, read_len=35, coverage=107.0)],
                    n_polymorphic=60, n_fixed=10, n_gaps=1, mutation_cutoff=10.0, polymorphism_cutoff=2.0, precision=1e-6, places=8,
                    del_prop=12.0, del_seed=0.0)


def user_evidence_inputs(outdir):
    d = dict(USER_DATASET)
    os.makedirs(outdir, exist_ok=True)
    d["bam"], d["fasta"] = os.path.join(outdir, "reference.bam"), os.path.join(outdir, "reference.fasta")
    spec = bq.SynthSpec(seed=d["seed"], read_sets=d["read_sets"], fasta=os.path.join(DIR, "lambda.fasta"), n_polymorphic=d["n_polymorphic"],
                        n_fixed=d["n_fixed"], n_gaps=d["n_gaps"])
    ctx = bq.Context(device=-1)
    ctx.synth_write(spec, d["bam"], d["fasta"])
    ctx.close()
    return d


def run_both_passes(cli, d, outdir, user_gd):
    import subprocess
    ec, im = helpers.cli_args(d, outdir)
    for args in (ec, im + ["--user-evidence", user_gd]):
        subprocess.run([cli] + [str(a) for a in args], check=True, capture_output=True, text=True)
    return open(os.path.join(outdir, "ra_mc_evidence.gd")).read()


def test_suite_user_evidence_file_through_the_oracle(built, tmp_path):
    """The oracle reads the suite's file like the reference build did (golden: lambda_user_evidence.ra_mc_evidence.gd): the JC row
    and the comment lines are passed over, every RA row comes out -- marked on a row the data reports anyway, as a row of its own
    for an allele no read shows, with forced sub-columns up to insert_position 10."""
    d = user_evidence_inputs(str(tmp_path / "in"))
    os.makedirs(str(tmp_path / "oracle"))
    got = run_both_passes(helpers.ORACLE_CLI, d, str(tmp_path / "oracle"), USER_GD)
    assert got == open(USER_WANT).read()
    rows = [line.split("\t") for line in got.splitlines() if line.startswith("RA\t") and "user_defined=1" in line]
    asked = [line.split("\t")[3:8] for line in open(USER_GD) if line.startswith("RA\t")]
    assert sorted(r[3:8] for r in rows) == sorted([a[0], a[1], a[2], a[3], a[4].strip()] for a in asked)
    if os.path.exists(helpers.REF_CLI):   # where the reference build is at hand: the golden is still what it writes
        os.makedirs(str(tmp_path / "ref"))
        assert run_both_passes(helpers.REF_CLI, d, str(tmp_path / "ref"), USER_GD) == got


def test_suite_user_evidence_file_forces_its_sub_columns_in_host_staging(built, tmp_path):
    d = user_evidence_inputs(str(tmp_path / "in"))
    forced = bq.Context(device=-1)
    forced.stage_bam(d["bam"], d["fasta"], user_evidence_gd=USER_GD, read_file_sets=[("lambda", 1)])
    s = forced.stream()
    have = {(int(p), int(k)) for p, k in zip(s["ins_parent"], s["ins_count"])}
    assert {(100, k) for k in range(1, 11)} <= have and (133, 1) in have     # rows at 101 / insert 10 and 134 / insert 1 (0-based parents)
    forced.close()
