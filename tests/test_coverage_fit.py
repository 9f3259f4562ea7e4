"""The coverage fit between the passes: brq_fit_coverage_file / brq_fit_coverage_distribution against the
reference's own CoverageDistribution::fit (coverage_distribution.cpp:115-400, 422-498).

tests/golden/coverage_fit/expected.tsv holds what the reference build computed (make_coverage_fit_golden.py) for the coverage
distributions of the test datasets and for synthetic histograms that reach the other branches.  The bar is the one of byte
work: every double identical (the fit is a chain of comparisons between nearly equal objective values: a last-bit difference
in one of them can end in another restart's optimum)."""
import os

import pytest

import breseq_b200 as bq
from helpers import GOLDEN

FIELDS = [("average", "average"), ("variance", "variance"), ("relative_variance", "relative_variance"),
          ("nb_fit_size", "nbinom_size_parameter"), ("nb_fit_mu", "nbinom_mean_parameter"),
          ("deletion_coverage_propagation_cutoff", "deletion_coverage_propagation_cutoff")]


def expected_rows():
    path = os.path.join(GOLDEN, "coverage_fit", "expected.tsv")
    lines = open(path).read().splitlines()
    head = lines[0].split("\t")
    return [dict(zip(head, l.split("\t"))) for l in lines[1:]]


def test_coverage_fit_matches_the_reference_bit_for_bit():
    ctx = bq.Context(device=-1)  # host arithmetic only
    rows = expected_rows()
    assert len(rows) >= 60
    bad = []
    for r in rows:
        got = ctx.fit_coverage_file(os.path.join(GOLDEN, r["histogram"]), float(r["pr_cutoff"]))
        for ref_name, name in FIELDS:
            if float(r[ref_name]).hex() != float(got[name]).hex():
                bad.append((r["histogram"], r["pr_cutoff"], ref_name, r[ref_name], repr(got[name])))
    assert not bad, bad[:10]


def test_coverage_fit_branches_are_covered():
    """the golden rows reach: a real fit, coarse bins, a runaway size parameter, no fit (fallbacks), a missing sequence"""
    rows = {(r["histogram"], r["pr_cutoff"]): r for r in expected_rows()}
    by_hist = {}
    for (h, _), r in rows.items():
        by_hist.setdefault(os.path.basename(h), r)
    assert float(by_hist["deep5000_nb.tab"]["nb_fit_mu"]) > 4000          # coarse bins, mean scaled back
    assert float(by_hist["poisson1000.tab"]["nb_fit_size"]) > 1e6          # Poisson-like: the size parameter runs away
    assert float(by_hist["low_coverage.tab"]["deletion_coverage_propagation_cutoff"]) == -1.0
    assert float(by_hist["empty.tab"]["deletion_coverage_propagation_cutoff"]) == -1.0
    assert float(by_hist["one_position.tab"]["nb_fit_mu"]) == 0.0          # no window: no fit
    assert float(by_hist["nb100_deletion_spike.tab"]["nb_fit_mu"]) > 90


def test_fit_without_error_count_fails():
    ctx = bq.Context(device=-1)
    with pytest.raises(bq.BrqError):
        ctx.fit_coverage_distribution(0, 0.01)
    with pytest.raises(bq.BrqError):
        ctx.fit_coverage_file("/nonexistent/coverage.tab", 0.01)
