"""REAL reads with their real base qualities through the path: the five lambda read files of the reference's own test suite
(/root/reference/tests/data/lambda/lambda_mixed_population.*.fastq.gz, 200 000 Illumina reads of 35 bases with N runs and
quality-2 tails), placed on the lambda sequence without gaps by tests/tools/ungapped_align.cpp (bowtie2, which breseq aligns with,
is not in this image), ~145 000 reads, ~105x.

* the oracle writes, byte for byte, what the reference build wrote for this BAM (tests/golden/real_lambda/: the evidence file and
  the hashes of the other three, made by tests/golden/make_real_lambda_golden.py) -- and still does where the build is at hand;
* host staging carries exactly the records the oracle counts;
* against the suite's own expected.gd for these reads (tests/lambda_polymorphism, aligned with bowtie2 and passed through the
  stages in front of the pileup, so not the same alignments): every consensus substitution it lists is among the rows found here,
  and the rows both have agree on the variant frequency.
Needs the read files, i.e. /root/reference: skipped where that is absent (the GPU box)."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import breseq_b200 as bq
import helpers

READS_DIR = "/root/reference/tests/data/lambda"
READ_FILES = [os.path.join(READS_DIR, "lambda_mixed_population.%s.fastq.gz" % x) for x in "AB345"]
GOLD = os.path.join(helpers.GOLDEN, "real_lambda")
LAMBDA_FASTA = os.path.join(helpers.GOLDEN, "reference_tests", "lambda.fasta")
DATASET = dict(seed=0, contig_lens=[48502], prefix="unused", read_sets=[dict(name="lambda", paired=False, read_len=35, coverage=105.0)],
               mutation_cutoff=10.0, polymorphism_cutoff=2.0, precision=1e-6, places=8, del_prop=12.0, del_seed=0.0)

pytestmark = pytest.mark.skipif(not all(os.path.exists(f) for f in READ_FILES), reason="the reference suite's lambda read files are not here")


def build_inputs(outdir):
    """BAM + FASTA of the placed reads; returns the dataset dict the helpers take."""
    import minibam
    os.makedirs(outdir, exist_ok=True)
    tool = os.path.join(outdir, "ungapped_align")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", tool, os.path.join(helpers.ROOT, "tests", "tools", "ungapped_align.cpp"), "-lz"], check=True)
    tsv = os.path.join(outdir, "placed.tsv")
    subprocess.run([tool, LAMBDA_FASTA, tsv] + READ_FILES, check=True, capture_output=True)
    name, seq = open(LAMBDA_FASTA).read().split("\n", 1)
    seq = seq.replace("\n", "")
    reads = []
    for line in open(tsv):
        n, flag, pos, bases, quals = line.rstrip("\n").split("\t")
        reads.append(dict(name=n, tid=0, pos=int(pos), cigar="%dM" % len(bases), seq=bases, qual=[ord(c) - 33 for c in quals], flag=int(flag),
                          tags={"RG": "lambda", "X1": 1}))
    reads.sort(key=lambda r: r["pos"])   # stable: file order within a position
    d = dict(DATASET)
    d["bam"], d["fasta"] = os.path.join(outdir, "reference.bam"), os.path.join(outdir, "reference.fasta")
    minibam.write(d["bam"], d["fasta"], [(name[1:], seq)], reads, read_groups=["lambda"])
    d["n_reads"] = len(reads)
    return d


def run_passes(cli, d, outdir, extra_ec=(), extra_im=()):
    os.makedirs(outdir, exist_ok=True)
    ec, im = helpers.cli_args(d, outdir)
    for args in (ec + list(extra_ec), im + list(extra_im)):
        subprocess.run([cli] + [str(a) for a in args], check=True, capture_output=True, text=True)


@pytest.fixture(scope="module")
def real(built, tmp_path_factory):
    root = str(tmp_path_factory.mktemp("real_lambda"))
    d = build_inputs(os.path.join(root, "in"))
    odir = os.path.join(root, "oracle")
    d["oracle_dir"], d["oracle_counts"], d["oracle_columns"] = odir, os.path.join(odir, "error_counts.tab"), os.path.join(odir, "columns.bin")
    d["oracle_rates"], d["oracle_gd"] = os.path.join(odir, "error_rates.tab"), os.path.join(odir, "ra_mc_evidence.gd")
    run_passes(helpers.ORACLE_CLI, d, odir, ["--counts-dump", d["oracle_counts"]], ["--columns-out", d["oracle_columns"]])
    return d


def sha256(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def test_oracle_writes_what_the_reference_build_wrote(real, tmp_path):
    assert sha256(real["bam"]) == open(os.path.join(GOLD, "inputs.sha256")).read().split()[0], "the placed reads are not the ones the golden was made from"
    assert real["n_reads"] > 140000
    assert open(real["oracle_gd"]).read() == open(os.path.join(GOLD, "ra_mc_evidence.gd")).read()
    want = dict(line.split()[::-1] for line in open(os.path.join(GOLD, "outputs.sha256")))
    for f in helpers.pass_output_names(real):
        assert sha256(os.path.join(real["oracle_dir"], f)) == want[f], f
    if os.path.exists(helpers.REF_CLI):
        out = str(tmp_path / "ref")
        run_passes(helpers.REF_CLI, real, out)
        for f in helpers.pass_output_names(real):
            assert sha256(os.path.join(out, f)) == want[f], "live reference build: " + f


def test_host_staging_carries_the_oracle_s_records(real):
    import test_staging
    ctx = bq.Context(device=-1)
    ctx.stage_bam(real["bam"], real["fasta"], **helpers.stage_kwargs(real))
    staged = (real, ctx, ctx.stream())
    test_staging.test_histogram_stream_matches_oracle_counts(staged)
    test_staging.test_unique_only_coverage_matches_oracle(staged)
    test_staging.test_score_stream_tallies_match_oracle(staged)
    ctx.close()


def test_rows_agree_with_the_suite_s_expected_gd(real):
    def rows(path):
        out = {}
        for line in open(path):
            c = line.rstrip("\n").split("\t")
            if c[0] == "RA":
                out[tuple(c[4:8])] = dict(f.split("=", 1) for f in c[8:] if "=" in f)
        return out
    want = rows(os.path.join(helpers.GOLDEN, "reference_tests", "lambda_polymorphism.gd"))
    got = rows(real["oracle_gd"])
    consensus_substitutions = {k for k, v in want.items() if v.get("prediction") == "consensus" and "." not in k[2:]}
    assert len(consensus_substitutions) >= 20 and consensus_substitutions <= set(got)   # (indels need gapped alignments)
    shared = set(want) & set(got)
    assert len(shared) >= 35
    diff = sorted(abs(float(want[k]["frequency"]) - float(got[k]["frequency"])) for k in shared)
    assert diff[len(diff) // 2] < 0.02 and diff[-1] < 0.25


def test_device_expander_logic_on_the_real_reads(real, tmp_path):
    """csrc/expand_core.h (what the kernels of csrc/expand.cu run per read and per (read, column)), executed serially on the CPU by
    tests/expand_check.cpp over the real reads: every stream equals host staging's bit for bit."""
    exe = str(tmp_path / "expand_check")
    csrc = os.path.join(helpers.ROOT, "breseq_b200", "csrc")
    srcs = [os.path.join(helpers.ROOT, "tests", "expand_check.cpp")] + [os.path.join(csrc, f) for f in
                                                                         ("staging.cpp", "synth.cpp", "bam_io.cpp", "inflate.cpp", "expand_plan.cpp")]
    subprocess.run(["g++", "-O2", "-std=c++17", "-w", "-o", exe] + srcs + ["-lz", "-lpthread"], check=True)
    p = subprocess.run([exe, "--bam", real["bam"], real["fasta"], "lambda"], capture_output=True, text=True)
    assert p.returncode == 0 and "equal" in p.stdout and "DIFFERENT" not in p.stdout, p.stdout + p.stderr
