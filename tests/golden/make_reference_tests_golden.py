#!/usr/bin/env python
"""Fixtures from the REFERENCE'S OWN TEST SUITE (/root/reference/tests/<test>/expected.gd: the annotated GenomeDiff a full breseq
run of that test has to reproduce) for the two steps behind pass 2 that this library restates: the Output stage's RA filter
(test_RA_evidence, identify_mutations.cpp:687-749) and the RA step of mutation prediction (mutation_predictor.cpp:1955-2211).

    python tests/golden/make_reference_tests_golden.py        # needs /root/reference

An expected.gd holds the RA rows as they left the filter (prediction=, consensus_reject=, polymorphism_reject=, reject=), the MC
rows, and the mutations made from them.  Per test this script writes tests/golden/reference_tests/<test>.gd with
  * the RA rows (minus annotation keys added later by the run -- gene_*, locus_tag*, snp_type, aa_*, codon_* ...: opaque to both
    steps; kept in full for lambda_polymorphism so that pass-through of arbitrary keys is exercised),
  * the MC rows,
  * the SNP / DEL / INS / SUB rows whose evidence is RA rows only, reduced to type, id, evidence, their columns and frequency= /
    insert_position=,
and the reference sequences as FASTA (from the tests' GenBank files, ORIGIN sections only); fisher_kat.tsv holds every distinct
(major_cov, minor_cov) -> fisher_strand_p_value of the whole suite's RA rows, known answers for pass 2's strand-bias test.  tests/golden/reference_tests/tests.json
lists, per test, the mode and the thresholds its command line sets (testcmd.sh) and which FASTA it reads.
tests/test_reference_suite.py strips the four filter fields from the RA rows, runs the filter and expects the rows back as they
were; it runs the prediction on the rows as they are and expects the mutation rows."""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "reference_tests")
REF_TESTS = "/root/reference/tests"
DATA = os.path.join(REF_TESTS, "data")

LAMBDA = ["lambda/lambda.gbk"]
LAMBDA_SPLIT = ["lambda/lambda.1-2.gbk", "lambda/lambda.3.gbk", "lambda/lambda.4.gbk", "lambda/lambda.5.gbk"]
# test -> (polymorphism_prediction, Settings members its command line changes, GenBank files, FASTA fixture, rename the one sequence)
TESTS = {
    "lambda_polymorphism": (True, {}, LAMBDA, "lambda.fasta", None),
    "lambda_polymorphism_ignore_low_quality": (True, {"polymorphism_frequency_cutoff": 0.2, "polymorphism_reject_indel_homopolymer_length": 0,
                                                      "polymorphism_reject_surrounding_homopolymer_length": 2}, LAMBDA, "lambda.fasta", None),
    "lambda_polymorphism_mapping_quality_mismatches": (True, {}, LAMBDA, "lambda.fasta", None),
    "lambda_polymorphism_soft_clipping": (True, {}, LAMBDA, "lambda.fasta", None),
    "lambda_polymorphism_user_evidence": (True, {}, LAMBDA, "lambda.fasta", None),
    "lambda_polymorphism_aligned_sam": (True, {}, LAMBDA, "lambda.fasta", None),
    "lambda_polymorphism_aligned_sam_user_evidence": (True, {}, LAMBDA, "lambda.fasta", None),
    "lambda_mixed_pop": (False, {}, LAMBDA, "lambda.fasta", "NC_001416.1"),          # --genbank-field-for-seq-id VERSION
    "lambda_mult_ref_read": (False, {}, LAMBDA_SPLIT, "lambda_split.fasta", None),
    "lambda_mult_ref_read_polymorphism": (True, {}, LAMBDA_SPLIT, "lambda_split.fasta", None),
    "bull_1": (False, {}, ["bull/bull_1.gbk"], "bull_1.fasta", None),
    "bull_2": (False, {}, ["bull/bull_2.gbk"], "bull_2.fasta", None),
    "REL606_tiled_reads_continuation": (True, {}, ["REL606/REL606.fragment.gbk"], "REL606_fragment.fasta", None),
    "lambda_polymorphism_no_junction_bad_orfs": (True, {}, ["lambda/lambda_bad_orfs.gbk"], "lambda.fasta", None),   # the same sequence
    # the filter only (no prediction): the reference sequence of the run is a download that is not here, or the run has
    # further references; the filter does not read the sequence under these settings
    "lambda_contig_ref": (False, {}, None, None, None),
    "lambda_mixed_pop_bad_contigs": (False, {}, None, None, None),
    "lambda_mixed_pop_cn_evidence": (False, {}, None, None, None),
    "lambda_mixed_pop_cn_no_coverage": (False, {}, None, None, None),
    "lambda_mixed_pop_custom_bowtie2": (False, {}, None, None, None),
    "lambda_mixed_pop_names_with_spaces": (False, {}, None, None, None),
    "lambda_short_sequence_repeats": (False, {}, None, None, None),
    "tmv_plasmid_circular_deletion": (False, {}, None, None, None),
    "tmv_plasmid_circular_deletion_end_only": (False, {}, None, None, None),
    "tmv_plasmid_circular_deletion_start_only": (False, {}, None, None, None),
    "tmv_plasmid_missing_pairs": (False, {}, None, None, None),
    "long_ltee_clone": (False, {}, None, None, None),
    "long_ltee_ara_m3_32k_mp2800": (False, {}, None, None, None),
    "long_ltee_ara_p1_50k_pe101": (False, {}, None, None, None),
    "long_ltee_ara_m3_38k_se36": (False, {}, None, None, None),
    "long_ltee_ara_m1_40k_pe36": (False, {}, None, None, None),
    "long_ltee_ara_p3_30k_pe150": (False, {}, None, None, None),
    "long_ltee_ara_p6_40k_se36": (False, {}, None, None, None),
}
MAX_ROWS_FILTER_ONLY = 150   # of a filter-only test, every k-th RA row (the filter takes each row on its own)
KEEP_ANNOTATION = {"lambda_polymorphism"}
ANNOTATION = re.compile(r"^(gene_|genes_|locus_tag|aa_|codon_|snp_type|transl_table|mutation_category|multiple_polymorphic)")
MUTATION_COLUMNS = {"SNP": 6, "DEL": 6, "INS": 6, "SUB": 7}


def genbank_sequences(paths):
    out = []
    for p in paths:
        name, seq, on = None, [], False
        for line in open(os.path.join(DATA, p)):
            if line.startswith("LOCUS"):
                name, seq, on = line.split()[1], [], False
            elif line.startswith("ORIGIN"):
                on = True
            elif line.startswith("//"):
                out.append((name, "".join(seq).upper()))
                on = False
            elif on:
                seq.append(re.sub(r"[^A-Za-z]", "", line))
    return out


def main():
    if not os.path.isdir(REF_TESTS):
        sys.exit("/root/reference/tests is not here")
    os.makedirs(OUT, exist_ok=True)
    index, written = {}, set()
    for test, (poly, settings, gbk, fasta, rename) in TESTS.items():
        if fasta and fasta not in written:
            with open(os.path.join(OUT, fasta), "w") as fh:
                for name, seq in genbank_sequences(gbk):
                    fh.write(">%s\n" % name + "".join(seq[i:i + 70] + "\n" for i in range(0, len(seq), 70)))
            written.add(fasta)
        rows = [line.rstrip("\n") for line in open(os.path.join(REF_TESTS, test, "expected.gd"))]
        if fasta and test == "lambda_polymorphism_no_junction_bad_orfs":
            assert genbank_sequences(gbk) == genbank_sequences(LAMBDA)
        if not fasta:
            ra = [r for r in rows if r.startswith("RA\t")]
            stride = max(1, -(-len(ra) // MAX_ROWS_FILTER_ONLY))
            rows = ra[::stride]
        ra_ids = {r.split("\t")[1] for r in rows if r.startswith("RA\t")} if fasta else set()
        keep = ["#=GENOME_DIFF\t1.0"]
        for r in rows:
            c = r.split("\t")
            if c[0] in MUTATION_COLUMNS and all(e in ra_ids for e in c[2].split(",")):
                n = MUTATION_COLUMNS[c[0]]
                keep.append("\t".join(c[:n] + [f for f in c[n:] if f.split("=")[0] in ("frequency", "insert_position")]))
        for r in rows:
            c = r.split("\t")
            if c[0] == "RA":
                if test not in KEEP_ANNOTATION:
                    c = c[:8] + [f for f in c[8:] if not ANNOTATION.match(f)]
                keep.append("\t".join(c))
        keep += [r for r in rows if r.startswith("MC\t")]
        open(os.path.join(OUT, test + ".gd"), "w").write("\n".join(keep) + "\n")
        index[test] = {"polymorphism_prediction": poly, "settings": settings, "fasta": fasta, "rename": rename}
        print(test, sum(1 for r in keep if r.startswith("RA\t")), "RA rows")
    # one more reference sequence, for the real reads of tests/test_real_reads.py
    with open(os.path.join(OUT, "tmv_plasmid.fasta"), "w") as fh:
        for name, seq in genbank_sequences(["tmv_plasmid/tmv-plasmid.gbk"]):
            fh.write(">%s\n" % name + "".join(seq[i:i + 70] + "\n" for i in range(0, len(seq), 70)))
    # known answers for the strand-bias test of pass 2's finalisation: every distinct (major_cov, minor_cov) -> fisher_strand_p_value
    # among the RA rows of ALL the suite's expected.gd files (the long_ltee_* runs included)
    import glob
    kat = {}
    for path in sorted(glob.glob(os.path.join(REF_TESTS, "*", "expected.gd"))):
        for line in open(path):
            if line.startswith("RA\t"):
                kv = dict(f.split("=", 1) for f in line.rstrip("\n").split("\t")[8:] if "=" in f)
                if all(k in kv for k in ("fisher_strand_p_value", "major_cov", "minor_cov")):
                    kat[(kv["major_cov"], kv["minor_cov"])] = kv["fisher_strand_p_value"]
    with open(os.path.join(OUT, "fisher_kat.tsv"), "w") as fh:
        fh.write("major_cov\tminor_cov\tfisher_strand_p_value\n" + "".join("%s\t%s\t%s\n" % (k[0], k[1], v) for k, v in sorted(kat.items())))
    print(len(kat), "Fisher known answers")
    # the two per-read-group BAM2COV tables of the suite, as they are (tests/test_coverage_table.py rebuilds reads out of them)
    import shutil
    os.makedirs(os.path.join(OUT, "bam2cov"), exist_ok=True)
    for name in ("no_read_groups", "multiple_read_groups"):
        shutil.copy(os.path.join(REF_TESTS, "bam2cov_per_read_group", "expected.%s.tab" % name), os.path.join(OUT, "bam2cov", "per_read_group.%s.tab" % name))
    # ... and its two `BAM2COV -a` tables at the default resolution (summary.json there: coverage_average 330.7552)
    shutil.copy(os.path.join(REF_TESTS, "bam2cov", "expected.tab"), os.path.join(OUT, "bam2cov", "show_average.tab"))
    shutil.copy(os.path.join(REF_TESTS, "bam2cov_csv", "expected.csv"), os.path.join(OUT, "bam2cov", "show_average.csv"))
    # the suite's user-evidence file, and what the reference build (oracle/_ref/ref_cli) reports for it on reads simulated over the
    # real lambda sequence (tests/test_reference_suite.py holds the read model): the golden of the oracle's and the library's parsers
    shutil.copy(os.path.join(REF_TESTS, "lambda_polymorphism_user_evidence", "user_evidence.gd"), os.path.join(OUT, "lambda_user_evidence.input.gd"))
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import tempfile
    import helpers
    import test_reference_suite as suite
    if os.path.exists(helpers.REF_CLI):
        with tempfile.TemporaryDirectory() as tmp:
            d = suite.user_evidence_inputs(os.path.join(tmp, "in"))
            os.makedirs(os.path.join(tmp, "ref"))
            open(suite.USER_WANT, "w").write(suite.run_both_passes(helpers.REF_CLI, d, os.path.join(tmp, "ref"), suite.USER_GD))
    else:
        print("oracle/_ref/ref_cli is missing: lambda_user_evidence.ra_mc_evidence.gd left as it is")
    with open(os.path.join(OUT, "tests.json"), "w") as fh:
        fh.write(json.dumps(index, indent=1) + "\n")


if __name__ == "__main__":
    main()
