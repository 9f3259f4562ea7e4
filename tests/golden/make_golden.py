#!/usr/bin/env python
"""Regenerate tests/golden/ from the REFERENCE'S OWN sources (oracle/_ref/ref_cli).

Run in the build container, where /root/reference exists:

    python -c 'import __graft_entry__ as g; g.build()'     # libbrq.so, oracle_cli and oracle/_ref/ref_cli
    python tests/golden/make_golden.py

For every dataset in tests/helpers.py:DATASETS the inputs (BAM + FASTA) are written by the product's
seeded generator and both reference entry points are run on them through oracle/ref_driver.cpp:
breseq::error_count() (error_count.cpp:50-68) and breseq::identify_mutations()
(identify_mutations.cpp:48-88), compiled unmodified from /root/reference/src/breseq against the
htslib shim.  What they wrote is committed here:

    <name>/error_rates.tab, base_qual_error_prob.*.tab, *.unique_only_coverage_distribution.tab,
    <name>/ra_mc_evidence.gd, <name>/inputs.sha256 (so generator drift is detected, not silently absorbed)
    <name>/preprocess_error_count.tab  (error_count(..., preprocess_stage = true): Summary::preprocess_error_count per seq id)
    tiny/reference.bam, tiny/reference.fasta(.fai), tiny/per_position_file.tab   (inputs kept for the smallest case)
    tiny/<seq>.coverage.tsv   (two read groups: the --predict-copy-number table with its per-read-group columns)
    tiny/, ltee/error_counts.per_position.sha256   (covariates with ref_pos: checksum of the reference's per-position error_counts.tab)

The fixtures pin (a) oracle/oracle.cpp, (b) the CUDA path, against the reference's real arithmetic
and file writers.  The BAM decode / pileup layer under the reference is still oracle/hts_shim.
"""
import hashlib
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402


def sha256(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    if not os.path.exists(helpers.REF_CLI):
        sys.exit("oracle/_ref/ref_cli is missing: run oracle/ref_build.sh where /root/reference exists")
    for name in helpers.DATASETS:
        if helpers.DATASETS[name].get("no_golden"):
            continue
        gdir = os.path.join(HERE, name)
        os.makedirs(gdir, exist_ok=True)   # (coverage_table.* are make_coverage_table_golden.py's: left alone)
        with tempfile.TemporaryDirectory() as tmp:
            d = helpers.generate_inputs(name, tmp)
            out = os.path.join(tmp, "ref")
            helpers.run_reference(d, out, coverage_tsv=(name in ("deep", "tiny")))
            for f in helpers.pass_output_names(d):
                shutil.copy(os.path.join(out, f), os.path.join(gdir, f))
            with open(os.path.join(gdir, "inputs.sha256"), "w") as fh:
                fh.write("%s  reference.bam\n%s  reference.fasta\n" % (sha256(d["bam"]), sha256(d["fasta"])))
            if d.get("big_table"):  # too many rows to commit: the checksum of what the reference wrote
                with open(os.path.join(gdir, "error_rates.tab.sha256"), "w") as fh:
                    fh.write("%s  error_rates.tab\n" % sha256(os.path.join(out, "error_rates.tab")))
            with open(os.path.join(gdir, helpers.PREPROCESS_TAB), "w") as fh:  # the stage 03 call (preprocess_stage = true)
                fh.write(helpers.run_preprocess(helpers.REF_CLI, d, os.path.join(tmp, "ref_preprocess")))
            if name == "lambda":  # user evidence (Settings::user_evidence_genome_diff_file_name): the committed input list, what the reference reports for it
                import subprocess
                user = os.path.join(HERE, "..", "user_evidence_lambda.gd")
                shutil.copy(user, os.path.join(gdir, "user_evidence.gd"))
                _, im = helpers.cli_args(d, out, gd=os.path.join(out, "user.gd"))
                subprocess.run([helpers.REF_CLI] + [str(a) for a in im] + ["--user-evidence", user], check=True, capture_output=True)
                shutil.copy(os.path.join(out, "user.gd"), os.path.join(gdir, "ra_mc_evidence.user_evidence.gd"))
            if name in ("tiny", "ltee"):  # the per-position count table of a covariate string with ref_pos (too long to commit: its checksum)
                import subprocess
                pp = os.path.join(tmp, "ref_per_position")
                os.makedirs(pp)
                ec, _ = helpers.cli_args(d, pp)
                ec[ec.index("--covariates") + 1] = "ref_pos," + helpers.covariates(d)
                subprocess.run([helpers.REF_CLI] + [str(a) for a in ec], check=True, capture_output=True)
                with open(os.path.join(gdir, "error_counts.per_position.sha256"), "w") as fh:
                    fh.write("%s  error_counts.tab\n" % sha256(os.path.join(pp, "error_counts.tab")))
            if name == "tiny":
                for f in ("reference.bam", "reference.fasta", "reference.fasta.fai"):
                    shutil.copy(os.path.join(tmp, f), os.path.join(gdir, f))
                shutil.copy(os.path.join(out, "per_position_file.tab"), os.path.join(gdir, "per_position_file.tab"))
                for c in helpers.contig_names(d):  # two read groups: <seq>.coverage.tsv with its per-read-group columns
                    shutil.copy(os.path.join(out, c + ".coverage.tsv"), os.path.join(gdir, c + ".coverage.tsv"))
            if name == "deep":  # one read group: the optional outputs of pass 2 for a whole (small) dataset
                shutil.copy(os.path.join(out, "per_position_file.tab"), os.path.join(gdir, "per_position_file.tab"))
                tsv = helpers.contig_names(d)[0] + ".coverage.tsv"
                shutil.copy(os.path.join(out, tsv), os.path.join(gdir, tsv))
        print("golden/%s: %d files" % (name, len(os.listdir(gdir))))


if __name__ == "__main__":
    main()
