#!/usr/bin/env python
"""Golden vectors of the coverage fit from the REFERENCE'S OWN code (oracle/_ref/ref_cli fit_coverage, which calls
CoverageDistribution::fit, coverage_distribution.cpp:422-498, compiled unmodified from /root/reference).

Run in the build container, after `python -c 'import __graft_entry__ as g; g.build()'`:

    python tests/golden/make_coverage_fit_golden.py

Inputs: the unique-only coverage distributions the reference wrote for the test datasets (tests/golden/<name>/), and a few
seeded synthetic histograms written to tests/golden/coverage_fit/ that reach the branches those do not (coarse bins for very
deep coverage, a Poisson-like histogram whose size parameter runs away, a deletion spike, histograms too small to smooth, an
empty one, coverage so low that the sequence counts as missing).  Output: tests/golden/coverage_fit/expected.tsv, one row per
(histogram, probability cutoff) with the reference's doubles at 17 significant digits.

CoverageDistribution::fit also draws a plot through `gnuplot`, which this image does not have: a do-nothing stand-in is put on
PATH for the reference binary (the plot is not part of the fixtures).
"""
import glob
import os
import stat
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "ref_cli")
OUT = os.path.join(HERE, "coverage_fit")
FIELDS = ["average", "variance", "relative_variance", "nb_fit_size", "nb_fit_mu", "deletion_coverage_propagation_cutoff"]
CUTOFFS = [0.05 / np.sqrt(4629812.0), 0.05 / np.sqrt(48502.0), 0.01, 1e-9]


def write_hist(name, counts):
    """counts[i] = positions of coverage i; rows 1 .. last non-zero, like error_count writes them (error_count.cpp:239-253)"""
    last = max([i for i, c in enumerate(counts) if c] + [0])
    with open(os.path.join(OUT, name), "w") as fh:
        fh.write("coverage\tn\n")
        for i in range(1, last + 1):
            fh.write("%d\t%d\n" % (i, counts[i]))


def synthetic():
    rng = np.random.default_rng(20260101)

    def hist_of(samples, size=None):
        return np.bincount(samples.astype(np.int64), minlength=size or 0).tolist()

    def nb(mu, size, n):
        return rng.negative_binomial(size, size / (size + mu), n)

    write_hist("deep5000_nb.tab", hist_of(nb(5000.0, 30.0, 2_000_000)))                 # window wider than 2000: coarse bins
    write_hist("poisson1000.tab", hist_of(rng.poisson(1000.0, 4_600_000)))               # C2-like: size runs to infinity
    write_hist("nb100_deletion_spike.tab", hist_of(np.concatenate([nb(100.0, 12.0, 500_000), rng.poisson(1.5, 40_000)])))
    write_hist("bimodal.tab", hist_of(np.concatenate([nb(60.0, 20.0, 300_000), nb(180.0, 25.0, 120_000)])))
    write_hist("low_coverage.tab", hist_of(rng.poisson(2.0, 50_000)))                    # mean <= 3: cutoff -1
    write_hist("four_bins.tab", [0, 5, 9, 4, 2])                                          # N < 5: no smoothing
    write_hist("one_position.tab", [0, 0, 0, 1])
    write_hist("flat.tab", [0] + [10] * 40)
    with open(os.path.join(OUT, "empty.tab"), "w") as fh:
        fh.write("coverage\tn\n")


def main():
    if not os.path.exists(REF_CLI):
        sys.exit("oracle/_ref/ref_cli is missing: run oracle/ref_build.sh where /root/reference exists")
    os.makedirs(OUT, exist_ok=True)
    synthetic()
    inputs = sorted(glob.glob(os.path.join(HERE, "*", "*.unique_only_coverage_distribution.tab"))) + sorted(glob.glob(os.path.join(OUT, "*.tab")))
    with tempfile.TemporaryDirectory() as tmp:
        fake = os.path.join(tmp, "bin")
        os.makedirs(fake)
        with open(os.path.join(fake, "gnuplot"), "w") as fh:
            fh.write("#!/bin/sh\nexit 0\n")
        os.chmod(os.path.join(fake, "gnuplot"), stat.S_IRWXU)
        env = dict(os.environ, PATH=fake + os.pathsep + os.environ["PATH"])
        rows = []
        for path in inputs:
            for pr in CUTOFFS:
                r = subprocess.run([REF_CLI, "fit_coverage", "--distribution", path, "--pr-cutoff", "%.17g" % pr, "--out", tmp],
                                   capture_output=True, text=True, env=env, cwd=tmp)
                got = dict(line.split("\t") for line in r.stdout.splitlines() if "\t" in line)
                if r.returncode != 0 or any(f not in got for f in FIELDS):
                    sys.exit("ref_cli fit_coverage failed on %s:\n%s%s" % (path, r.stdout, r.stderr))
                rows.append([os.path.relpath(path, HERE), "%.17g" % pr] + [got[f] for f in FIELDS])
    with open(os.path.join(OUT, "expected.tsv"), "w") as fh:
        fh.write("\t".join(["histogram", "pr_cutoff"] + FIELDS) + "\n")
        for r in rows:
            fh.write("\t".join(r) + "\n")
    print("wrote %d rows" % len(rows))


if __name__ == "__main__":
    main()
