#!/usr/bin/env python
"""Golden outputs of the REFERENCE BUILD (oracle/_ref/ref_cli) for the real reads of the reference's test suite (lambda, tmv plasmid), placed
without gaps (tests/test_real_reads.py builds the BAMs): tests/golden/real_<name>/ra_mc_evidence.gd, outputs.sha256 (all four
files of the two passes) and inputs.sha256 (the BAM).

    python tests/golden/make_real_reads_golden.py        # after g.build(); needs /root/reference"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import test_real_reads as t  # noqa: E402


def main():
    if not os.path.exists(helpers.REF_CLI):
        sys.exit("oracle/_ref/ref_cli is missing: run oracle/ref_build.sh where /root/reference exists")
    for name in sorted(t.REAL):
        gold = t.gold_dir(name)
        os.makedirs(gold, exist_ok=True)
        with tempfile.TemporaryDirectory() as tmp:
            d = t.build_inputs(name, os.path.join(tmp, "in"))
            out = os.path.join(tmp, "ref")
            t.run_passes(helpers.REF_CLI, d, out)
            shutil.copy(os.path.join(out, "ra_mc_evidence.gd"), os.path.join(gold, "ra_mc_evidence.gd"))
            # the coverage histogram of real reads: one more input of the coverage-fit goldens (make_coverage_fit_golden.py globs it)
            shutil.copy(os.path.join(out, "0.unique_only_coverage_distribution.tab"), os.path.join(gold, "0.unique_only_coverage_distribution.tab"))
            with open(os.path.join(gold, "outputs.sha256"), "w") as fh:
                for f in helpers.pass_output_names(d):
                    fh.write("%s  %s\n" % (t.sha256(os.path.join(out, f)), f))
            with open(os.path.join(gold, "coverage_tables.sha256"), "w") as fh:   # BAM2COV's tables of the same BAM (coverage_output::table)
                for cname, span, resolution, total_only, per_rg in t.COVERAGE_REQUESTS:
                    table = os.path.join(tmp, cname + ".tab")
                    subprocess.run([helpers.REF_CLI, "coverage_table", "--bam", d["bam"], "--fasta", d["fasta"], "--region", t.coverage_request_args(d, span),
                                    "--resolution", str(resolution), "--total-only", str(total_only), "--format", "tsv", "--per-read-group", str(per_rg),
                                    "--table", table], check=True, cwd=tmp, capture_output=True)
                    fh.write("%s  %s\n" % (t.sha256(table), cname))
            d["real_name"] = name
            import hashlib
            filtered, predicted = t.ra_chain_reference(d, d["polymorphism_cutoff"] < d["mutation_cutoff"], os.path.join(tmp, "chain"))
            with open(os.path.join(gold, "ra_chain.sha256"), "w") as fh:   # the Output stage's RA filter and the RA step of mutation prediction
                fh.write("%s  filtered\n%s  predicted\n" % (hashlib.sha256(filtered.encode()).hexdigest(), hashlib.sha256(predicted.encode()).hexdigest()))
            with open(os.path.join(gold, "inputs.sha256"), "w") as fh:
                fh.write("%s  reference.bam (%d reads)\n" % (t.sha256(d["bam"]), d["n_reads"]))
        print("real_" + name + ":", d["n_reads"], "reads")


if __name__ == "__main__":
    main()
