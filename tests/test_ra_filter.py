"""The Output stage's filter over pass 2's RA rows (brq_test_ra_evidence, csrc/ra_filter.cpp) against what the reference's own
test_RA_evidence (identify_mutations.cpp:687-749) wrote for the same files and settings (tests/golden/make_ra_filter_golden.py).
Host only: no device involved."""
import hashlib
import json
import os
import re

import pytest

import breseq_b200 as bq
import helpers

GOLD = os.path.join(helpers.GOLDEN, "ra_filter")
SETS = json.load(open(os.path.join(GOLD, "option_sets.json")))
SWITCHES = [(0, 0), (1, 0), (0, 1)]   # (targeted_sequencing, call_mutations_overlapping_missing_coverage), as in the generator
EXPECTED = [line.rstrip("\n").split("\t") for line in list(open(os.path.join(GOLD, "expected.tsv")))[1:]]


@pytest.fixture(scope="module")
def ctx(built):
    c = bq.Context(device=-1)
    yield c
    c.close()


def filtered(ctx, gd_in, fasta, option_set, tmp_path):
    out = str(tmp_path / "filtered.gd")
    counts = ctx.test_RA_evidence(gd_in, fasta, out, option_set["polymorphism_prediction"], **option_set["settings"])
    return open(out).read(), counts


@pytest.mark.parametrize("k", range(len(SETS)))
def test_corner_cases_like_the_reference(ctx, k, tmp_path):
    text, counts = filtered(ctx, os.path.join(GOLD, "edge.gd"), os.path.join(GOLD, "edge.fasta"), SETS[k], tmp_path)
    assert text == open(os.path.join(GOLD, "edge.%d.gd" % k)).read()
    kept = [line for line in text.splitlines() if line.startswith("RA\t")]
    assert counts["rows"] == 17 and counts["rows"] - counts["deleted"] == len(kept)
    assert counts["consensus"] + counts["polymorphism"] + counts["rejected_kept"] >= len(kept)   # a consensus call of the reference base is counted, then dropped
    # what is not an RA row passes through untouched
    assert [line for line in text.splitlines() if not line.startswith("RA\t")] == \
        [line for line in open(os.path.join(GOLD, "edge.gd")).read().splitlines() if not line.startswith("RA\t")]


@pytest.mark.parametrize("k", range(len(SETS)))
def test_evidence_without_recorded_bounds_like_the_reference(ctx, k, tmp_path):
    """Rows of an older breseq carry no frequency_lower / frequency_upper: Clopper-Pearson bounds from total_cov
    (identify_mutations.cpp:215-228), the point estimate where there is no depth."""
    text, counts = filtered(ctx, os.path.join(GOLD, "legacy.gd"), os.path.join(GOLD, "edge.fasta"), SETS[k], tmp_path)
    assert text == open(os.path.join(GOLD, "legacy.%d.gd" % k)).read()


def test_binomial_bounds_bit_for_bit(built):
    """binomial_frequency_lower_bound / _upper_bound (stats.cpp:2394-2414, Cephes incbi / ndtri / incbet underneath) against
    what the reference build printed with %.17g: exact, degenerate arguments included."""
    rows = [line.split("\t") for line in list(open(os.path.join(GOLD, "binomial_bounds.tsv")))[1:]]
    assert len(rows) > 1000
    for k, n, alpha, lower, upper in rows:
        lo, hi = bq.binomial_frequency_bounds(float(k), float(n), float(alpha))
        assert ("%.17g" % lo, "%.17g" % hi) == (lower, upper.strip()), (k, n, alpha)
    lo, hi = bq.binomial_frequency_bounds(5, 5)
    assert lo == 0.05 ** (1 / 5) and hi == 1.0              # all reads agree: alpha^(1/k)
    assert bq.binomial_frequency_bounds(0, 8) == (0.0, 1.0 - 0.05 ** (1 / 8))


def test_every_reject_reason_is_exercised():
    """The golden files are only worth what they reach: every reason the filter can give shows up in them."""
    seen = set()
    for k in range(len(SETS)):
        for m in re.finditer(r"reject=([A-Z_,]+)", open(os.path.join(GOLD, "edge.%d.gd" % k)).read()):
            seen.update(m.group(1).split(","))
    assert seen >= {"SCORE_CUTOFF", "FREQUENCY_CUTOFF", "VARIANT_STRAND_COVERAGE", "TOTAL_STRAND_COVERAGE", "VARIANT_COVERAGE", "TOTAL_COVERAGE",
                    "INDEL_HOMOPOLYMER", "SURROUNDING_HOMOPOLYMER", "POLYMORPHIC_INDEL", "KS_BASE_QUALITY", "FISHER_STRAND", "EXISTING"}


def test_dataset_evidence_like_the_reference(ctx, datasets, tmp_path):
    """The reference's own pass-2 files of the test datasets through every option set: the hash of what the reference kept."""
    for name, gd, k, kept, digest, *predicted in EXPECTED:
        text, counts = filtered(ctx, os.path.join(helpers.GOLDEN, name, gd), datasets[name]["fasta"], SETS[int(k)], tmp_path)
        assert hashlib.sha256(text.encode()).hexdigest() == digest, (name, gd, k)
        assert counts["rows"] - counts["deleted"] == int(kept)
        # ... and on into the RA step of mutation prediction, under its three switches
        for (targeted, over_mc), want in zip(SWITCHES, predicted):
            out = str(tmp_path / "predicted.gd")
            n = ctx.predict_ra_mutations(str(tmp_path / "filtered.gd"), datasets[name]["fasta"], out, SETS[int(k)]["polymorphism_prediction"], targeted, over_mc)
            got = open(out).read()
            assert hashlib.sha256(got.encode()).hexdigest() == want, (name, gd, k, targeted, over_mc)
            assert n["SNP"] + n["DEL"] + n["INS"] + n["SUB"] == sum(1 for line in got.splitlines() if line[:4] in ("SNP\t", "DEL\t", "INS\t", "SUB\t"))


@pytest.mark.parametrize("k", range(len(SETS)))
def test_corner_cases_become_the_reference_s_mutations(ctx, k, tmp_path):
    out = str(tmp_path / "predicted.gd")
    n = ctx.predict_ra_mutations(os.path.join(GOLD, "edge.%d.gd" % k), os.path.join(GOLD, "edge.fasta"), out, SETS[k]["polymorphism_prediction"])
    assert open(out).read() == open(os.path.join(GOLD, "edge.%d.predicted.gd" % k)).read()
    row_1_stayed = "\nRA\t1\t" in open(os.path.join(GOLD, "edge.%d.gd" % k)).read()
    assert n["ra_marked_deleted"] == int(row_1_stayed)    # row 1 sits in the MC row at 1-2


def test_neighbouring_calls_join(ctx, tmp_path):
    """Hand-made: two neighbouring consensus SNPs become one SUB, a run of deleted bases one DEL, inserted columns 1 and 2 one INS;
    in polymorphism mode a polymorphic neighbour stays on its own; an insertion that starts at column 2 is not called in
    consensus mode (mutation_predictor.cpp:2140-2146)."""
    fasta = tmp_path / "j.fasta"
    fasta.write_text(">j\nACGTACGTACGTACGTACGT\n")
    common = "frequency=1.0e+00\tmajor_cov=9/9\tminor_cov=0/0\ttotal_cov=9/9\tscore=50.0\tprediction=consensus"
    def ra(i, pos, ins, ref, new, extra=common):
        return "RA\t%d\t.\tj\t%d\t%d\t%s\t%s\tmajor_base=%s\tminor_base=%s\t%s" % (i, pos, ins, ref, new, new, ref, extra)
    poly = common.replace("frequency=1.0e+00", "frequency=4.0e-01").replace("prediction=consensus", "prediction=polymorphism")
    rows = [ra(1, 2, 0, "C", "T"), ra(2, 3, 0, "G", "A"),                      # -> SUB 2 (size 2, TA)
            ra(3, 6, 0, "C", "."), ra(4, 7, 0, "G", "."), ra(5, 8, 0, "T", "."),   # -> DEL 6 size 3
            ra(6, 10, 1, ".", "G"), ra(7, 10, 2, ".", "G"),                    # -> INS 10 GG
            ra(8, 13, 2, ".", "A"),                                            # starts at column 2: dropped in consensus mode
            ra(9, 15, 0, "G", "C"), ra(10, 16, 0, "T", "C", poly)]             # consensus + polymorphic neighbour
    gd = tmp_path / "j.gd"
    gd.write_text("#=GENOME_DIFF\t1.0\n" + "\n".join(rows) + "\n")
    out = tmp_path / "o.gd"
    n = ctx.predict_ra_mutations(str(gd), str(fasta), str(out))
    muts = [line.split("\t") for line in out.read_text().splitlines() if line[:2] not in ("RA", "#=")]
    assert [m[:1] + m[2:] for m in muts] == [["SUB", "1,2", "j", "2", "2", "TA"], ["DEL", "3,4,5", "j", "6", "3"], ["INS", "6,7", "j", "10", "GG"],
                                            ["SNP", "9", "j", "15", "C"]]
    assert [m[1] for m in muts] == ["11", "12", "13", "14"] and n == {"SNP": 1, "DEL": 1, "INS": 1, "SUB": 1, "ra_marked_deleted": 0}
    n = ctx.predict_ra_mutations(str(gd), str(fasta), str(out), polymorphism_prediction=True)
    muts = [line.split("\t") for line in out.read_text().splitlines() if line[:2] not in ("RA", "#=")]
    assert [m[:1] + m[2:] for m in muts] == [
        ["SUB", "1,2", "j", "2", "2", "TA", "frequency=1"], ["DEL", "3,4,5", "j", "6", "3", "frequency=1"],
        ["INS", "6,7", "j", "10", "GG", "frequency=1", "insert_position=1"], ["INS", "8", "j", "13", "A", "frequency=1", "insert_position=2"],
        ["SNP", "9", "j", "15", "C", "frequency=1"], ["SNP", "10", "j", "16", "C", "frequency=4.0e-01"]]


def test_modes_differ_where_the_reference_says(ctx, tmp_path):
    """Consensus mode drops a row that answers neither question, polymorphism mode keeps it with reject=; user_defined rows always stay."""
    gd, fa = os.path.join(GOLD, "edge.gd"), os.path.join(GOLD, "edge.fasta")
    cons, _ = filtered(ctx, gd, fa, SETS[0], tmp_path)
    poly, _ = filtered(ctx, gd, fa, SETS[1], tmp_path)
    ids = lambda text: {line.split("\t")[1] for line in text.splitlines() if line.startswith("RA\t")}
    assert "13" not in ids(cons) and "13" in ids(poly)
    row13 = [line for line in poly.splitlines() if line.startswith("RA\t13\t")][0]
    assert "\treject=" in row13 and "\tprediction=polymorphism" in row13 and "polymorphism_reject" not in row13
    assert "9" in ids(cons) and "12" not in ids(cons) and "12" not in ids(poly)


def test_defaults_are_the_modes_of_settings_cpp(ctx):
    c, p = ctx.ra_filter_defaults(False), ctx.ra_filter_defaults(True)
    assert (c["consensus_frequency_cutoff"], c["polymorphism_frequency_cutoff"], c["polymorphism_log10_e_value_cutoff"]) == (0.5, 0.1, 10.0)   # settings.cpp:918-948
    assert (p["consensus_frequency_cutoff"], p["polymorphism_frequency_cutoff"], p["polymorphism_log10_e_value_cutoff"]) == (0.95, 0.05, 2.0)  # settings.cpp:862-896
    for o in (c, p):
        assert o["polymorphism_minimum_variant_coverage_each_strand"] == 2 and o["polymorphism_fisher_strand_p_value_cutoff"] == 0.05
        assert o["polymorphism_ks_quality_p_value_cutoff"] == 0 and o["mutation_log10_e_value_cutoff"] == 10


def test_errors_are_loud(ctx, tmp_path):
    fa = os.path.join(GOLD, "edge.fasta")
    with pytest.raises(bq.BrqError, match="cannot open"):
        ctx.test_RA_evidence(str(tmp_path / "absent.gd"), fa, str(tmp_path / "o.gd"))
    bad = tmp_path / "bad.gd"
    bad.write_text("#=GENOME_DIFF\t1.0\nRA\t1\t.\tedge\t5\t0\tT\tA\tfrequency=1\tmajor_base=A\tmajor_cov=5/5\ttotal_cov=5/5\n")
    with pytest.raises(bq.BrqError, match="score"):
        ctx.test_RA_evidence(str(bad), fa, str(tmp_path / "o.gd"))
    bad.write_text("#=GENOME_DIFF\t1.0\nRA\t1\t.\telsewhere\t5\t0\tT\t.\tfrequency=1\tfrequency_lower=0.9\tfrequency_upper=1\tscore=20\tmajor_base=.\tmajor_cov=5/5\ttotal_cov=5/5\n")
    with pytest.raises(bq.BrqError, match="elsewhere"):
        ctx.test_RA_evidence(str(bad), fa, str(tmp_path / "o.gd"), consensus_reject_indel_homopolymer_length=3)
    with pytest.raises(bq.BrqError, match="no member"):
        ctx.test_RA_evidence(str(bad), fa, str(tmp_path / "o.gd"), not_a_setting=1)


@pytest.mark.skipif(not os.path.exists(helpers.REF_CLI), reason="oracle/_ref/ref_cli (the reference build) is not here")
@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_rows_against_the_reference_build(ctx, seed, tmp_path):
    """Rows and thresholds drawn at random (runs of equal bases in the sequence, thin strands, p-values and scores on either side
    of the cutoffs, user_defined and pre-rejected rows), live through ref_cli test_ra and through the library."""
    import random
    import subprocess
    rng = random.Random(seed)
    seq = "".join(rng.choice("ACGT") * rng.choice([1, 1, 1, 2, 3, 5, 7]) for _ in range(120))
    fasta = tmp_path / "random.fasta"
    fasta.write_text(">r\n%s\n" % seq)
    rows, rid = [], 0
    for pos in sorted(rng.sample(range(1, len(seq) + 1), 90)):
        for ins in ([0] if rng.random() < 0.7 else [1, 2]):
            rid += 1
            ref = "." if ins else seq[pos - 1]
            new = rng.choice("ACGT") if ins else rng.choice("ACGT.".replace(ref, ""))
            major, minor = (new, ref) if rng.random() < 0.5 else (ref, new)
            f = rng.random()
            lo, hi = max(0.0, f - rng.random() * 0.3), min(1.0, f + rng.random() * 0.3)
            cov = lambda: "%d/%d" % (rng.choice([0, 1, 2, 3, 8, 20, 40]), rng.choice([0, 1, 2, 5, 9, 25, 40]))
            kv = {"frequency": "%.8e" % f, "frequency_lower": "%.8e" % lo, "frequency_upper": "%.8e" % hi, "major_base": major,
                  "minor_base": minor, "major_cov": cov(), "minor_cov": cov(), "total_cov": cov(),
                  "fisher_strand_p_value": "%.5e" % rng.choice([1.0, 0.5, 0.04, 1e-5]), "ks_quality_p_value": "%.5e" % rng.choice([1.0, 0.3, 0.02]),
                  "score": rng.choice(["NA", "%.1f" % (rng.random() * 60), "1.5", "12.0"])}
            if rng.random() < 0.1:
                kv["user_defined"] = "1"
            if rng.random() < 0.1:
                kv["reject"] = rng.choice(["EXISTING", "SCORE_CUTOFF", "A,B"])
            if rng.random() < 0.1:
                del kv["score"]
                kv[rng.choice(["consensus_score", "polymorphism_score"])] = "%.1f" % (rng.random() * 40)
            rows.append("\t".join(["RA", str(rid), ".", "r", str(pos), str(ins), ref, new] + ["%s=%s" % (k, kv[k]) for k in sorted(kv)]))
    for k in range(4):   # missing coverage over some of the rows
        start = rng.randrange(1, len(seq) - 20)
        rid += 1
        rows.append("MC\t%d\t.\tr\t%d\t%d\t0\t0\tleft_inside_cov=0\tleft_outside_cov=9\tright_inside_cov=0\tright_outside_cov=9" % (rid, start, start + rng.randrange(0, 12)))
    rows.sort(key=lambda line: (line[:2] != "RA", int(line.split("\t")[4]), int(line.split("\t")[5])))   # RA rows, then MC rows, by position: the GenomeDiff order
    n_ra = sum(1 for line in rows if line.startswith("RA\t"))
    gd = tmp_path / "random.gd"
    gd.write_text("#=GENOME_DIFF\t1.0\n" + "\n".join(rows) + "\n")
    for trial in range(4):
        poly = rng.random() < 0.5
        settings = {"mutation_log10_e_value_cutoff": rng.choice([10.0, 30.0]), "polymorphism_log10_e_value_cutoff": rng.choice([2.0, 10.0]),
                    "consensus_frequency_cutoff": rng.choice([0.0, 0.5, 0.95]), "polymorphism_frequency_cutoff": rng.choice([0.0, 0.05, 0.2]),
                    "polymorphism_fisher_strand_p_value_cutoff": rng.choice([0.0, 0.05]), "polymorphism_ks_quality_p_value_cutoff": rng.choice([0.0, 0.05]),
                    "polymorphism_no_indels": rng.choice([0, 1])}
        for name in ("consensus_minimum_variant_coverage", "consensus_minimum_total_coverage", "consensus_minimum_variant_coverage_each_strand",
                     "consensus_minimum_total_coverage_each_strand", "polymorphism_minimum_variant_coverage", "polymorphism_minimum_total_coverage",
                     "polymorphism_minimum_variant_coverage_each_strand", "polymorphism_minimum_total_coverage_each_strand"):
            settings[name] = rng.choice([0, 0, 2, 10])
        for name in ("consensus_reject_indel_homopolymer_length", "polymorphism_reject_indel_homopolymer_length",
                     "consensus_reject_surrounding_homopolymer_length", "polymorphism_reject_surrounding_homopolymer_length"):
            settings[name] = rng.choice([0, 2, 4, 6])
        own, counts = filtered(ctx, str(gd), str(fasta), {"polymorphism_prediction": poly, "settings": settings}, tmp_path)
        out = tmp_path / "ref.gd"
        args = [helpers.REF_CLI, "test_ra", "--fasta", str(fasta), "--gd-in", str(gd), "--gd-out", str(out), "--out", str(tmp_path)]
        args += ["--polymorphism-prediction"] if poly else []
        for k, v in settings.items():
            args += ["--" + k, repr(v)]
        subprocess.run(args, check=True, capture_output=True, cwd=str(tmp_path))
        ref = "".join(line for line in open(out) if not line.startswith("#=TITLE"))
        assert own == ref, (seed, trial, settings)
        assert counts["rows"] == n_ra
        # on into the RA step of mutation prediction
        targeted, over_mc = rng.choice([(0, 0), (0, 0), (1, 0), (0, 1)])
        predicted = str(tmp_path / "predicted.gd")
        ctx.predict_ra_mutations(str(tmp_path / "filtered.gd"), str(fasta), predicted, poly, targeted, over_mc)
        args = [helpers.REF_CLI, "predict_ra", "--fasta", str(fasta), "--gd-in", str(tmp_path / "filtered.gd"), "--gd-out", str(out), "--out", str(tmp_path),
                "--targeted-sequencing", str(targeted), "--call-mutations-overlapping-missing-coverage", str(over_mc)]
        args += ["--polymorphism-prediction"] if poly else []
        subprocess.run(args, check=True, capture_output=True, cwd=str(tmp_path))
        ref = "".join(line for line in open(out) if not line.startswith("#=TITLE"))
        assert open(predicted).read() == ref, (seed, trial, "predict", targeted, over_mc)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted({row[0] for row in EXPECTED}))
def test_from_the_bam_to_mutation_rows(name, datasets, tmp_path):
    """The whole chain on the CUDA path -- error_count, identify_mutations, the RA filter, the RA step of mutation prediction --
    ends in the files the reference's chain ends in (both modes of the filter, the default switches of the prediction)."""
    from test_golden import run_cuda
    d = datasets[name]
    out = str(tmp_path / "cuda")
    run_cuda(d, out)
    c = bq.Context(device=-1)
    for _, gd, k, kept, digest, predicted, *_ in [row for row in EXPECTED if row[0] == name and row[1] == "ra_mc_evidence.gd" and row[2] in ("0", "1")]:
        poly = SETS[int(k)]["polymorphism_prediction"]
        counts = c.test_RA_evidence(os.path.join(out, gd), d["fasta"], str(tmp_path / "filtered.gd"), poly)
        assert hashlib.sha256(open(tmp_path / "filtered.gd", "rb").read()).hexdigest() == digest and counts["rows"] - counts["deleted"] == int(kept)
        c.predict_ra_mutations(str(tmp_path / "filtered.gd"), d["fasta"], str(tmp_path / "predicted.gd"), poly)
        assert hashlib.sha256(open(tmp_path / "predicted.gd", "rb").read()).hexdigest() == predicted
    c.close()
