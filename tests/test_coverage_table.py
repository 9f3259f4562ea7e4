"""BAM2COV's per-position coverage table (coverage_output.cpp:190-283, 307-470).

The goldens (tests/golden/<name>/coverage_table.<k>.tab, requests in coverage_tables.tsv) were written by the reference build's
own coverage_output::table (make_coverage_table_golden.py).  Here, without a GPU: the per-column function the device kernel
wraps (csrc/expand_core.h: coverage_lane) run serially by tests/coverage_check.cpp, with the product's table writer, must
reproduce every file byte for byte.  tests/test_gpu_coverage_table.py asks the same of the CUDA path through the C ABI."""
import filecmp
import os
import subprocess

import pytest

import helpers

ROOT = helpers.ROOT


def requests(name):
    lines = open(os.path.join(helpers.GOLDEN, name, "coverage_tables.tsv")).read().splitlines()[1:]
    return [l.split("\t") for l in lines]


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("covcheck") / "coverage_check")
    csrc = os.path.join(ROOT, "breseq_b200", "csrc")
    srcs = [os.path.join(ROOT, "tests", "coverage_check.cpp")] + [os.path.join(csrc, f) for f in
                                                                   ("staging.cpp", "bam_io.cpp", "expand_plan.cpp", "coverage_table.cpp")]
    subprocess.run(["g++", "-O2", "-std=c++17", "-w", "-I/usr/local/cuda/include", "-o", exe] + srcs + ["-lz", "-lpthread"], check=True)
    return exe


@pytest.mark.parametrize("name", [n for n in helpers.DATASETS if not helpers.DATASETS[n].get("no_golden")])
def test_coverage_walk_reproduces_the_reference_tables(name, checker, datasets, tmp_path):
    d = datasets[name]
    reqs = requests(name)
    assert len(reqs) == 5
    for table, region, resolution, total_only, fmt in reqs:
        out = str(tmp_path / table)
        p = subprocess.run([checker, d["bam"], d["fasta"], region, resolution, total_only, "1" if fmt == "csv" else "0", out],
                           capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        assert filecmp.cmp(out, os.path.join(helpers.GOLDEN, name, table), shallow=False), (name, table, region)


def test_bad_regions_are_refused(checker, datasets, tmp_path):
    d = datasets["tiny"]
    first = helpers.contig_names(d)[0]
    for region in ("nosuchseq:1-10", first, first + ":0-10", first + ":10-5", first + ":1-99999999", first + ":1-2-3"):
        p = subprocess.run([checker, d["bam"], d["fasta"], region, "0", "0", "0", str(tmp_path / "x.tab")], capture_output=True, text=True)
        assert p.returncode == 1 and "coverage_check:" in p.stderr, region
