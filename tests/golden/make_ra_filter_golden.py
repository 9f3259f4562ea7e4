#!/usr/bin/env python
"""Golden outputs of the Output stage's RA filter from the REFERENCE'S OWN code (oracle/_ref/ref_cli test_ra: cGenomeDiff::read,
test_RA_evidence (identify_mutations.cpp:687-749), cGenomeDiff::write, compiled unmodified from /root/reference).

    python tests/golden/make_ra_filter_golden.py        # after g.build(); needs /root/reference

Inputs: every committed ra_mc_evidence.gd of tests/golden/<dataset>/ (the reference's own pass-2 output on the test datasets)
and tests/golden/ra_filter/edge.gd + edge.fasta, a hand-written file of the corner cases (indels at either end of a run and of
the sequence, substitutions joining runs, a row that already carries reject=, user_defined rows, legacy score fields, score=NA,
a consensus call of the reference base, a row failing both questions), and legacy.gd, the same rows without frequency_lower /
frequency_upper (evidence of an older breseq: the filter rebuilds Clopper-Pearson bounds from total_cov; binomial_bounds.tsv pins
those bounds themselves).  Each goes through the option sets of
tests/golden/ra_filter/option_sets.json (members of breseq::Settings by name over the mode's defaults).  The reference's writer
adds a #=TITLE line from the file name, which is dropped.  Every filtered file then goes through ref_cli predict_ra
(MutationPredictor::predictRAtoSNPorDELorINSorSUB, mutation_predictor.cpp:1955-2211) three times: as is, as a targeted-sequencing
run, and with mutations called over missing coverage.  Results: tests/golden/ra_filter/edge.<set>.gd and
edge.<set>.predicted.gd in full and, for the datasets, tests/golden/ra_filter/expected.tsv (dataset, file, set, rows kept, sha256
of the filtered file and of the three predicted files)."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import breseq_b200 as bq  # noqa: E402

OUT = os.path.join(HERE, "ra_filter")
DATASET_FILES = [("tiny", "ra_mc_evidence.gd"), ("ltee", "ra_mc_evidence.gd"), ("deep", "ra_mc_evidence.gd"), ("lambda", "ra_mc_evidence.gd"),
                 ("lambda", "ra_mc_evidence.user_evidence.gd"), ("multi", "ra_mc_evidence.gd")]


def option_sets():
    return json.load(open(os.path.join(OUT, "option_sets.json")))


def reference_filter(ctx, gd_in, fasta, option_set, tmp):
    """The filtered file as the reference writes it, minus its #=TITLE line."""
    full = ctx.ra_filter_defaults(option_set["polymorphism_prediction"])   # every member goes on the command line
    full.update(option_set["settings"])
    out = os.path.join(tmp, "filtered.gd")
    args = [helpers.REF_CLI, "test_ra", "--fasta", fasta, "--gd-in", gd_in, "--gd-out", out, "--out", tmp]
    if option_set["polymorphism_prediction"]:
        args.append("--polymorphism-prediction")
    for k, v in full.items():
        if k != "polymorphism_prediction":
            args += ["--" + k, repr(v)]
    subprocess.run(args, check=True, capture_output=True, cwd=tmp)
    return "".join(line for line in open(out) if not line.startswith("#=TITLE"))


SWITCHES = [(0, 0), (1, 0), (0, 1)]   # (targeted_sequencing, call_mutations_overlapping_missing_coverage)


def reference_predict(filtered_text, fasta, polymorphism_prediction, targeted, over_mc, tmp):
    """predict_ra on a filtered file: the mutation rows in front of the evidence rows, minus the #=TITLE line."""
    gd_in, out = os.path.join(tmp, "to_predict.gd"), os.path.join(tmp, "predicted.gd")
    open(gd_in, "w").write(filtered_text)
    args = [helpers.REF_CLI, "predict_ra", "--fasta", fasta, "--gd-in", gd_in, "--gd-out", out, "--out", tmp,
            "--targeted-sequencing", str(targeted), "--call-mutations-overlapping-missing-coverage", str(over_mc)]
    if polymorphism_prediction:
        args.append("--polymorphism-prediction")
    subprocess.run(args, check=True, capture_output=True, cwd=tmp)
    return "".join(line for line in open(out) if not line.startswith("#=TITLE"))


def write_binomial_bounds():
    """binomial_frequency_lower_bound / _upper_bound (stats.cpp:2394-2414) of the reference build for a fixed list of (k, n, alpha)."""
    import random
    rng = random.Random(7)
    pairs = [(f * n, n) for n in (1, 2, 3, 5, 10, 29, 100, 292, 1000, 1e5) for f in (0, 0.001, 0.02, 0.24, 0.48, 0.5, 0.9, 0.982, 0.999, 1.0)]
    for _ in range(300):
        n = rng.choice([rng.randint(1, 50), rng.randint(1, 3000), rng.random() * 200])
        pairs.append((rng.random() * n if rng.random() < 0.7 else float(rng.randint(0, int(n))), n))
    pairs += [(0.5, 0.5), (0.2, 1.0), (1e-9, 10), (10 - 1e-9, 10), (-1, 5), (7, 5), (3, 0), (0.3, 0.9)]
    with open(os.path.join(OUT, "binomial_bounds.tsv"), "w") as fh:
        fh.write("k\tn\talpha\tlower\tupper\n")
        for alpha in (0.05, 0.5, 0.001):
            p = subprocess.run([helpers.REF_CLI, "binomial_bounds", "--alpha", repr(alpha)], input="".join("%r %r\n" % kn for kn in pairs),
                               capture_output=True, text=True, check=True)
            for (k, n), line in zip(pairs, p.stdout.strip().splitlines()):
                fh.write("%r\t%r\t%r\t%s\n" % (k, n, alpha, line))


def main():
    if not os.path.exists(helpers.REF_CLI):
        sys.exit("oracle/_ref/ref_cli is missing: run oracle/ref_build.sh where /root/reference exists")
    ctx = bq.Context(device=-1)
    sets = option_sets()
    with tempfile.TemporaryDirectory() as tmp:
        for k, s in enumerate(sets):
            text = reference_filter(ctx, os.path.join(OUT, "edge.gd"), os.path.join(OUT, "edge.fasta"), s, tmp)
            open(os.path.join(OUT, "edge.%d.gd" % k), "w").write(text)
            predicted = reference_predict(text, os.path.join(OUT, "edge.fasta"), s["polymorphism_prediction"], 0, 0, tmp)
            open(os.path.join(OUT, "edge.%d.predicted.gd" % k), "w").write(predicted)
            # the same rows as an older breseq wrote them, without frequency_lower / frequency_upper: Clopper-Pearson bounds
            text = reference_filter(ctx, os.path.join(OUT, "legacy.gd"), os.path.join(OUT, "edge.fasta"), s, tmp)
            open(os.path.join(OUT, "legacy.%d.gd" % k), "w").write(text)
        write_binomial_bounds()
        rows = []
        fasta = {}
        for name, gd in DATASET_FILES:
            if name not in fasta:
                fasta[name] = helpers.generate_inputs(name, os.path.join(tmp, name))["fasta"]
            for k, s in enumerate(sets):
                text = reference_filter(ctx, os.path.join(HERE, name, gd), fasta[name], s, tmp)
                kept = sum(1 for line in text.splitlines() if line.startswith("RA\t"))
                digests = [hashlib.sha256(text.encode()).hexdigest()]
                for targeted, over_mc in SWITCHES:
                    predicted = reference_predict(text, fasta[name], s["polymorphism_prediction"], targeted, over_mc, tmp)
                    digests.append(hashlib.sha256(predicted.encode()).hexdigest())
                rows.append("%s\t%s\t%d\t%d\t%s" % (name, gd, k, kept, "\t".join(digests)))
    with open(os.path.join(OUT, "expected.tsv"), "w") as fh:
        fh.write("dataset\tfile\toption_set\tra_rows_kept\tsha256\tsha256_predicted\tsha256_predicted_targeted\tsha256_predicted_over_mc\n" + "\n".join(rows) + "\n")
    print(len(sets), "option sets,", len(rows), "dataset outputs")


if __name__ == "__main__":
    main()
