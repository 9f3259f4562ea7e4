#!/usr/bin/env python
"""Golden outputs of the Output stage's RA filter from the REFERENCE'S OWN code (oracle/_ref/ref_cli test_ra: cGenomeDiff::read,
test_RA_evidence (identify_mutations.cpp:687-749), cGenomeDiff::write, compiled unmodified from /root/reference).

    python tests/golden/make_ra_filter_golden.py        # after g.build(); needs /root/reference

Inputs: every committed ra_mc_evidence.gd of tests/golden/<dataset>/ (the reference's own pass-2 output on the test datasets)
and tests/golden/ra_filter/edge.gd + edge.fasta, a hand-written file of the corner cases (indels at either end of a run and of
the sequence, substitutions joining runs, a row that already carries reject=, user_defined rows, legacy score fields, score=NA,
a consensus call of the reference base, a row failing both questions).  Each goes through the option sets of
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
