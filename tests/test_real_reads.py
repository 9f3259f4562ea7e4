"""REAL reads with their real base qualities through the path: read files of the reference's own test suite -- the five lambda
files (/root/reference/tests/data/lambda/lambda_mixed_population.*.fastq.gz, 200 000 Illumina reads of 35 bases with N runs and
quality-2 tails; ~145 000 placed, ~105x) and the paired 150-base reads of its tmv_plasmid tests (9193 pairs, binned qualities; two
read files, ~200x on 10.4 kb) -- placed on their reference without gaps by tests/tools/ungapped_align.cpp (bowtie2, which breseq
aligns with, is not in this image).

* the oracle writes, byte for byte, what the reference build wrote for this BAM (tests/golden/real_lambda/: the evidence file and
  the hashes of the other three, made by tests/golden/make_real_reads_golden.py) -- and still does where the build is at hand;
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

DATA = "/root/reference/tests/data"
FIXTURES = os.path.join(helpers.GOLDEN, "reference_tests")
# name -> reference FASTA (fixture), read files (one per read file of the run: two = a paired set, second file flagged 128), the
# read set's name, run settings (lambda: polymorphism mode like the suite's lambda_polymorphism; tmv: consensus mode like its tmv tests)
REAL = {
    "lambda": dict(fasta=os.path.join(FIXTURES, "lambda.fasta"),
                   files=[[os.path.join(DATA, "lambda", "lambda_mixed_population.%s.fastq.gz" % x) for x in "AB345"]], set_name="lambda",
                   settings=dict(mutation_cutoff=10.0, polymorphism_cutoff=2.0, precision=1e-6, places=8, del_prop=12.0, del_seed=0.0), min_reads=140000),
    # 9193 pairs of 150-base reads (binned qualities) on a 10.4 kb plasmid that carries two overlapping deletions
    "tmv": dict(fasta=os.path.join(FIXTURES, "tmv_plasmid.fasta"),
                files=[[os.path.join(DATA, "tmv_plasmid", "D3-9_1P.fastq.gz")], [os.path.join(DATA, "tmv_plasmid", "D3-9_2P.fastq.gz")]], set_name="D3-9",
                settings=dict(mutation_cutoff=10.0, polymorphism_cutoff=10.0, precision=1e-6, places=3, del_prop=20.0, del_seed=0.0), min_reads=13000),
}

pytestmark = pytest.mark.skipif(not all(os.path.exists(f) for r in REAL.values() for group in r["files"] for f in group),
                                reason="the reference suite's read files are not here")


def gold_dir(name):
    return os.path.join(helpers.GOLDEN, "real_" + name)


def build_inputs(name, outdir):
    """BAM + FASTA of the placed reads; returns the dataset dict the helpers take."""
    import minibam
    r = REAL[name]
    os.makedirs(outdir, exist_ok=True)
    tool = os.path.join(outdir, "ungapped_align")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", tool, os.path.join(helpers.ROOT, "tests", "tools", "ungapped_align.cpp"), "-lz"], check=True)
    contig, seq = open(r["fasta"]).read().split("\n", 1)
    seq = seq.replace("\n", "")
    paired = len(r["files"]) == 2
    reads = []
    for k, group in enumerate(r["files"]):
        tsv = os.path.join(outdir, "placed%d.tsv" % k)
        subprocess.run([tool, r["fasta"], tsv] + group, check=True, capture_output=True)
        for line in open(tsv):
            n, flag, pos, bases, quals = line.rstrip("\n").split("\t")
            flag = int(flag) | ((1 | (64 if k == 0 else 128)) if paired else 0)
            reads.append(dict(name=n.split()[0], tid=0, pos=int(pos), cigar="%dM" % len(bases), seq=bases, qual=[ord(c) - 33 for c in quals], flag=flag,
                              tags={"RG": r["set_name"], "X1": 1}))
    reads.sort(key=lambda x: x["pos"])   # stable: file order within a position
    d = dict(seed=0, contig_lens=[len(seq)], prefix="unused", read_sets=[dict(name=r["set_name"], paired=paired, read_len=len(reads[0]["seq"]), coverage=100.0)])
    d.update(r["settings"])
    d["bam"], d["fasta"] = os.path.join(outdir, "reference.bam"), os.path.join(outdir, "reference.fasta")
    minibam.write(d["bam"], d["fasta"], [(contig[1:], seq)], reads, read_groups=[r["set_name"]])
    d["n_reads"] = len(reads)
    return d


def run_passes(cli, d, outdir, extra_ec=(), extra_im=()):
    os.makedirs(outdir, exist_ok=True)
    ec, im = helpers.cli_args(d, outdir)
    for args in (ec + list(extra_ec), im + list(extra_im)):
        subprocess.run([cli] + [str(a) for a in args], check=True, capture_output=True, text=True)


@pytest.fixture(scope="module", params=sorted(REAL))
def real(request, built, tmp_path_factory):
    root = str(tmp_path_factory.mktemp("real_" + request.param))
    d = build_inputs(request.param, os.path.join(root, "in"))
    d["real_name"] = request.param
    odir = os.path.join(root, "oracle")
    d["oracle_dir"], d["oracle_counts"], d["oracle_columns"] = odir, os.path.join(odir, "error_counts.tab"), os.path.join(odir, "columns.bin")
    d["oracle_rates"], d["oracle_gd"] = os.path.join(odir, "error_rates.tab"), os.path.join(odir, "ra_mc_evidence.gd")
    run_passes(helpers.ORACLE_CLI, d, odir, ["--counts-dump", d["oracle_counts"]], ["--columns-out", d["oracle_columns"]])
    return d


def sha256(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def test_oracle_writes_what_the_reference_build_wrote(real, tmp_path):
    gold = gold_dir(real["real_name"])
    assert sha256(real["bam"]) == open(os.path.join(gold, "inputs.sha256")).read().split()[0], "the placed reads are not the ones the golden was made from"
    assert real["n_reads"] > REAL[real["real_name"]]["min_reads"]
    assert open(real["oracle_gd"]).read() == open(os.path.join(gold, "ra_mc_evidence.gd")).read()
    want = dict(line.split()[::-1] for line in open(os.path.join(gold, "outputs.sha256")))
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
    if real["real_name"] != "lambda":
        pytest.skip("the lambda reads only")
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


@pytest.fixture(scope="module")
def expand_check(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("expand_check") / "expand_check")
    csrc = os.path.join(helpers.ROOT, "breseq_b200", "csrc")
    srcs = [os.path.join(helpers.ROOT, "tests", "expand_check.cpp")] + [os.path.join(csrc, f) for f in
                                                                         ("staging.cpp", "synth.cpp", "bam_io.cpp", "inflate.cpp", "expand_plan.cpp")]
    subprocess.run(["g++", "-O2", "-std=c++17", "-w", "-o", exe] + srcs + ["-lz", "-lpthread"], check=True)
    return exe


def test_device_expander_logic_on_the_real_reads(real, expand_check):
    """csrc/expand_core.h (what the kernels of csrc/expand.cu run per read and per (read, column)), executed serially on the CPU by
    tests/expand_check.cpp over the real reads: every stream equals host staging's bit for bit."""
    set_name = REAL[real["real_name"]]["set_name"]
    p = subprocess.run([expand_check, "--bam", real["bam"], real["fasta"], set_name] + (["paired"] if len(REAL[real["real_name"]]["files"]) == 2 else []),
                       capture_output=True, text=True)
    assert p.returncode == 0 and "equal" in p.stdout and "DIFFERENT" not in p.stdout, p.stdout + p.stderr


COVERAGE_REQUESTS = [("whole", None, 600, 0, 0), ("window", (2001, 2600), 0, 0, 1)]   # name, span (None = all), resolution, total_only, per_read_group


def coverage_request_args(d, span):
    contig = open(d["fasta"]).readline()[1:].split()[0]
    lo, hi = span or (1, d["contig_lens"][0])
    return "%s:%d-%d" % (contig, lo, min(hi, d["contig_lens"][0]))


@pytest.fixture(scope="module")
def coverage_check(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("covcheck") / "coverage_check")
    csrc = os.path.join(helpers.ROOT, "breseq_b200", "csrc")
    srcs = [os.path.join(helpers.ROOT, "tests", "coverage_check.cpp")] + [os.path.join(csrc, f) for f in
                                                                           ("staging.cpp", "bam_io.cpp", "inflate.cpp", "expand_plan.cpp", "coverage_table.cpp")]
    subprocess.run(["g++", "-O2", "-std=c++17", "-w", "-I/usr/local/cuda/include", "-o", exe] + srcs + ["-lz", "-lpthread"], check=True)
    return exe


def test_bam2cov_tables_of_the_real_reads(real, coverage_check, tmp_path):
    """BAM2COV's table over the real reads -- the whole sequence thinned to 600 rows, a window at full resolution with the
    per-read-group columns -- from the walk the device kernel wraps and the product's writer: the files the reference build's
    coverage_output::table wrote (hashes in tests/golden/real_<name>/coverage_tables.sha256)."""
    want = dict(line.split()[::-1] for line in open(os.path.join(gold_dir(real["real_name"]), "coverage_tables.sha256")))
    for name, span, resolution, total_only, per_rg in COVERAGE_REQUESTS:
        region = coverage_request_args(real, span)
        out = str(tmp_path / (name + ".tab"))
        p = subprocess.run([coverage_check, real["bam"], real["fasta"], region, str(resolution), str(total_only), "0", out, str(per_rg)], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        assert sha256(out) == want[name], (real["real_name"], name)
        if os.path.exists(helpers.REF_CLI):
            ref_out = str(tmp_path / (name + ".ref.tab"))
            subprocess.run([helpers.REF_CLI, "coverage_table", "--bam", real["bam"], "--fasta", real["fasta"], "--region", region, "--resolution", str(resolution),
                            "--total-only", str(total_only), "--format", "tsv", "--per-read-group", str(per_rg), "--table", ref_out], check=True, cwd=str(tmp_path),
                           capture_output=True)
            assert sha256(ref_out) == want[name], "live reference build: " + name


def ra_chain_reference(d, polymorphism_prediction, workdir):
    """(filtered, predicted) texts of the reference build for the dataset's evidence file, mode defaults, minus the #=TITLE lines."""
    os.makedirs(workdir, exist_ok=True)
    gd = os.path.join(gold_dir(d["real_name"]), "ra_mc_evidence.gd")
    mode = ["--polymorphism-prediction"] if polymorphism_prediction else []
    ctx = bq.Context(device=-1)
    knobs = []
    for k, v in ctx.ra_filter_defaults(polymorphism_prediction).items():
        if k != "polymorphism_prediction":
            knobs += ["--" + k, repr(v)]
    ctx.close()
    filtered, predicted = os.path.join(workdir, "filtered.gd"), os.path.join(workdir, "predicted.gd")
    subprocess.run([helpers.REF_CLI, "test_ra", "--fasta", d["fasta"], "--gd-in", gd, "--gd-out", filtered, "--out", workdir] + mode + knobs,
                   check=True, capture_output=True, cwd=workdir)
    subprocess.run([helpers.REF_CLI, "predict_ra", "--fasta", d["fasta"], "--gd-in", filtered, "--gd-out", predicted, "--out", workdir] + mode,
                   check=True, capture_output=True, cwd=workdir)
    strip = lambda path: "".join(line for line in open(path) if not line.startswith("#=TITLE"))
    return strip(filtered), strip(predicted)


def test_filter_and_prediction_on_the_real_evidence(real, tmp_path):
    """The evidence of the real reads on through the Output stage's RA filter and the RA step of mutation prediction (the mode of
    the run, default thresholds): the reference build's files (hashes in tests/golden/real_<name>/ra_chain.sha256)."""
    poly = real["polymorphism_cutoff"] < real["mutation_cutoff"]
    want = dict(line.split()[::-1] for line in open(os.path.join(gold_dir(real["real_name"]), "ra_chain.sha256")))
    ctx = bq.Context(device=-1)
    filtered, predicted = str(tmp_path / "filtered.gd"), str(tmp_path / "predicted.gd")
    ctx.test_RA_evidence(real["oracle_gd"], real["fasta"], filtered, poly)
    n = ctx.predict_ra_mutations(filtered, real["fasta"], predicted, poly)
    ctx.close()
    assert sha256(filtered) == want["filtered"] and sha256(predicted) == want["predicted"]
    assert n["SNP"] >= 1
    if os.path.exists(helpers.REF_CLI):
        ref_filtered, ref_predicted = ra_chain_reference(real, poly, str(tmp_path / "ref"))
        assert open(filtered).read() == ref_filtered and open(predicted).read() == ref_predicted
