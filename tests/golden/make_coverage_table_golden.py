#!/usr/bin/env python
"""Golden BAM2COV tables from the REFERENCE'S OWN code (oracle/_ref/ref_cli coverage_table, which calls coverage_output::table,
coverage_output.cpp:190-283, compiled unmodified from /root/reference).

    python tests/golden/make_coverage_table_golden.py        # after g.build(); needs /root/reference

For every test dataset (tests/helpers.py: inputs written by the product's seeded generator; inputs.sha256 pins them) a few
(region, resolution, total_only, format) requests: the whole first sequence thinned to about 600 rows, a window with commas in
its coordinates at full resolution, the tail of the last sequence as totals in CSV, a single position, and two tables with
the per-read-group column sets.  The tables land in
tests/golden/<name>/coverage_table.<k>.tab and the requests in tests/golden/<name>/coverage_tables.tsv."""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402


def requests_for(d):
    names, lens = helpers.contig_names(d), d["contig_lens"]
    first, last, n0, n1 = names[0], names[-1], lens[0], lens[-1]
    lo = min(101, n0)
    hi = min(n0, lo + 1199)
    return [("%s:1-%d" % (first, n0), 600, 0, "tsv"),
            ("%s:%s-%s" % (first, "{:,}".format(lo), "{:,}".format(hi)), 0, 0, "tsv"),
            ("%s:%d-%d" % (last, max(1, n1 - 199), n1), 0, 1, "csv"),
            ("%s:%d" % (first, min(50, n0)), 0, 0, "tsv"),
            ("%s:1-%d" % (last, n1), 37, 1, "tsv"),
            ("%s:%d-%d" % (first, min(201, n0), min(n0, 500)), 0, 0, "tsv", 1),      # per read group, all columns
            ("%s:1-%d" % (last, n1), 90, 1, "csv", 1)]                              # per read group, totals, thinned


def main():
    if not os.path.exists(helpers.REF_CLI):
        sys.exit("oracle/_ref/ref_cli is missing: run oracle/ref_build.sh where /root/reference exists")
    for name in helpers.DATASETS:
        if helpers.DATASETS[name].get("no_golden"):
            continue
        gdir = os.path.join(HERE, name)
        with tempfile.TemporaryDirectory() as tmp:
            d = helpers.generate_inputs(name, tmp)
            rows = []
            for k, req in enumerate(requests_for(d)):
                region, resolution, total_only, fmt = req[:4]
                per_rg = req[4] if len(req) > 4 else 0
                out = os.path.join(gdir, "coverage_table.%d.tab" % k)
                subprocess.run([helpers.REF_CLI, "coverage_table", "--bam", d["bam"], "--fasta", d["fasta"], "--region", region,
                                "--resolution", str(resolution), "--total-only", str(total_only), "--format", fmt, "--per-read-group", str(per_rg),
                                "--table", out],
                               check=True, cwd=tmp)
                rows.append("\t".join([os.path.basename(out), region, str(resolution), str(total_only), fmt, str(per_rg)]))
            with open(os.path.join(gdir, "coverage_tables.tsv"), "w") as fh:
                fh.write("table\tregion\tresolution\ttotal_only\tformat\tper_read_group\n" + "\n".join(rows) + "\n")
        print(name, len(rows))


if __name__ == "__main__":
    main()
