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
                                                                   ("staging.cpp", "bam_io.cpp", "inflate.cpp", "expand_plan.cpp", "coverage_table.cpp")]
    subprocess.run(["g++", "-O2", "-std=c++17", "-w", "-I/usr/local/cuda/include", "-o", exe] + srcs + ["-lz", "-lpthread"], check=True)
    return exe


@pytest.mark.parametrize("name", [n for n in helpers.DATASETS if not helpers.DATASETS[n].get("no_golden")])
def test_coverage_walk_reproduces_the_reference_tables(name, checker, datasets, tmp_path):
    d = datasets[name]
    reqs = requests(name)
    assert len(reqs) == 7
    for table, region, resolution, total_only, fmt, per_rg in reqs:
        out = str(tmp_path / table)
        p = subprocess.run([checker, d["bam"], d["fasta"], region, resolution, total_only, "1" if fmt == "csv" else "0", out, per_rg],
                           capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        assert filecmp.cmp(out, os.path.join(helpers.GOLDEN, name, table), shallow=False), (name, table, region)


def test_bad_regions_are_refused(checker, datasets, tmp_path):
    d = datasets["tiny"]
    first = helpers.contig_names(d)[0]
    for region in ("nosuchseq:1-10", first, first + ":0-10", first + ":10-5", first + ":1-99999999", first + ":1-2-3"):
        p = subprocess.run([checker, d["bam"], d["fasta"], region, "0", "0", "0", str(tmp_path / "x.tab")], capture_output=True, text=True)
        assert p.returncode == 1 and "coverage_check:" in p.stderr, region


# ---- known answers worked out by hand from coverage_output.cpp:307-470 -----------------------------------------------------
# One BAM with the awkward cases side by side on a 19-base reference.  Columns (1-based) and what each read adds:
#   r1  forward  2M1D2M at 1   unique, aligned at 1 2 4 5 (3 is deleted: no coverage there); begins at 1 (top)
#   r2  reverse  2S3M1S at 5   unique, aligned at 5 6 7; a reversed read begins at its LAST base: query index 5 of 6 is soft
#                              clipped, so no aligned position is its first base: no begin count
#   r3  reverse  4M at 3       unique, aligned at 3..6; begins at 6 (bottom)
#   r4  forward  4M at 2, X1=3 redundant: 1/3 at 2..5, raw count 1
#   r5  forward  2M2N2M at 8   unique, aligned at 8 9 12 13 (10 11 are a reference skip); begins at 8
#   r6  forward  3H4M at 10    unique, aligned at 10..13; hard clips are not part of the read: begins at 10
#   r7  reverse  4M at 10, X1=2 redundant on the bottom strand: 0.5 at 10..13
#   r8  unmapped flag          never seen
KAT_REF = "ACGTACGTACGTTTGACCA"
KAT_READS = [dict(tid=0, pos=0, cigar="2M1D2M", seq="ACTA", qual=[30] * 4),
             dict(tid=0, pos=1, cigar="4M", seq="CGTA", qual=[30] * 4, tags={"X1": 3}),
             dict(tid=0, pos=2, cigar="4M", seq="GTAC", qual=[30] * 4, flag=16),
             dict(tid=0, pos=4, cigar="2S3M1S", seq="TTACGA", qual=[30] * 6, flag=16),
             dict(tid=0, pos=7, cigar="2M2N2M", seq="TATT", qual=[30] * 4),
             dict(tid=0, pos=9, cigar="3H4M", seq="CGTT", qual=[30] * 4),
             dict(tid=0, pos=9, cigar="4M", seq="CGTT", qual=[30] * 4, flag=16, tags={"X1": 2}),
             dict(tid=0, pos=12, cigar="4M", seq="TTGA", qual=[30] * 4, flag=4)]
#            pos: unique_top unique_bot red_top red_bot raw_top raw_bot begin_top begin_bot
KAT_ROWS = {1: (1, 0, "0", "0", 0, 0, 1, 0), 2: (1, 0, "0.333333", "0", 1, 0, 0, 0), 3: (0, 1, "0.333333", "0", 1, 0, 0, 0),
            4: (1, 1, "0.333333", "0", 1, 0, 0, 0), 5: (1, 2, "0.333333", "0", 1, 0, 0, 0), 6: (0, 2, "0", "0", 0, 0, 0, 1),
            7: (0, 1, "0", "0", 0, 0, 0, 0), 8: (1, 0, "0", "0", 0, 0, 1, 0), 9: (1, 0, "0", "0", 0, 0, 0, 0),
            10: (1, 0, "0", "0.5", 0, 1, 1, 0), 11: (1, 0, "0", "0.5", 0, 1, 0, 0), 12: (2, 0, "0", "0.5", 0, 1, 0, 0),
            13: (2, 0, "0", "0.5", 0, 1, 0, 0), 14: (0, 0, "0", "0", 0, 0, 0, 0)}


def kat_inputs(tmp_path):
    import minibam
    bam, fasta = str(tmp_path / "kat.bam"), str(tmp_path / "kat.fasta")
    minibam.write(bam, fasta, [("chr", KAT_REF)], KAT_READS)
    return bam, fasta


def kat_expected_text():
    lines = ["position\tref_base\tunique_top_cov\tunique_bot_cov\tredundant_top_cov\tredundant_bot_cov\traw_redundant_top_cov\t"
             "raw_redundant_bot_cov\tunique_top_begin\tunique_bot_begin"]
    for pos in range(1, 15):
        lines.append("\t".join([str(pos), KAT_REF[pos - 1]] + [str(x) for x in KAT_ROWS[pos]]))
    return "\n".join(lines) + "\n"


def test_hand_derived_coverage_rows(checker, tmp_path):
    """deletion, reference skip, soft and hard clips, both strands, redundant and unmapped reads: the walk, and the reference
    build itself where it is at hand, against rows worked out by hand"""
    bam, fasta = kat_inputs(tmp_path)
    mine = str(tmp_path / "mine.tab")
    p = subprocess.run([checker, bam, fasta, "chr:1-14", "0", "0", "0", mine], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    body = "".join(l for l in open(mine) if not l.startswith("#"))
    assert body == kat_expected_text()
    if os.path.exists(helpers.REF_CLI):
        ref = str(tmp_path / "ref.tab")
        subprocess.run([helpers.REF_CLI, "coverage_table", "--bam", bam, "--fasta", fasta, "--region", "chr:1-14", "--table", ref],
                       check=True, cwd=str(tmp_path))
        assert open(ref).read() == open(mine).read()


# ---- the reference's own BAM2COV fixture, with its missing input rebuilt from the table itself ------------------------------------
def reads_from_table(rows, seq, first, last):
    """A set of reads whose coverage IS the table.  Per strand the read-begin column gives one end of every read (a forward read
    begins at its leftmost base, a reversed one at its rightmost: coverage_output.cpp:395-447); the other ends are placed so that
    as few reads as possible are open (the tightest monotone envelope of the coverage column), and where the coverage dips below
    the number of open reads, that many of them carry a deleted base there -- which BAM2COV does not count.  Open reads are closed
    oldest first; deletions go to the newest ones.  Reads reach 20 bases past either end of the window."""
    positions = list(range(first, last + 1))
    lo, hi = first - 20, last + 20
    out = []
    for strand, cov_col, begin_col in ((0, 0, 6), (1, 1, 7)):
        cov = [rows[p][cov_col] for p in positions]
        begin = [rows[p][begin_col] for p in positions]
        n = len(positions)
        if strand == 0:   # opens known (begin), closes chosen: open[i] = R[i] + B[i], R non-increasing and as small as the coverage allows
            B = [0] * n
            for i in range(1, n):
                B[i] = B[i - 1] + begin[i]
            R = [0] * n
            for i in range(n - 1, -1, -1):
                R[i] = max(cov[i] - B[i], R[i + 1] if i + 1 < n else 0)
            opened = [R[0]] + begin[1:]
            closed_after = [R[i] - R[i + 1] for i in range(n - 1)] + [0]          # reads whose last base is positions[i]
            first_lefts = [lo] * (R[0] - begin[0]) + [first] * begin[0]
        else:             # closes known (begin = right ends), opens chosen: open[i] = S[i] - C[i], S non-decreasing and minimal
            closed_after = begin[:]
            C = [0] * n
            for i in range(1, n):
                C[i] = C[i - 1] + closed_after[i - 1]
            S = [0] * n
            for i in range(n):
                S[i] = max(cov[i] + C[i], S[i - 1] if i else 0)
            opened = [S[0]] + [S[i] - S[i - 1] for i in range(1, n)]
            first_lefts = [lo] * S[0]
        open_reads = []   # [left, deleted positions], oldest first
        for i, p in enumerate(positions):
            open_reads += [[left, []] for left in first_lefts] if i == 0 else [[p, []] for _ in range(opened[i])]
            deleted = len(open_reads) - cov[i]
            carriers = [r for r in open_reads[closed_after[i]:] if r[0] < p]      # not ending here, not starting here
            assert 0 <= deleted <= len(carriers) and closed_after[i] <= len(open_reads), (p, strand)
            for r in carriers[len(carriers) - deleted:]:
                r[1].append(p)
            for left, dels in open_reads[:closed_after[i]]:
                out.append((left, p, strand, dels))
            open_reads = open_reads[closed_after[i]:]
        out += [(left, hi, strand, dels) for left, dels in open_reads]
    out.sort(key=lambda r: (r[0], r[1], r[2]))
    reads = []
    for left, right, strand, dels in out:
        cigar, bases, run_start, p = "", "", left, left
        while p <= right:
            if p in dels:
                q = p
                while q + 1 in dels:
                    q += 1
                cigar += "%dM%dD" % (p - run_start, q - p + 1)
                bases += seq[run_start - 1:p - 1]
                run_start = p = q + 1
            else:
                p += 1
        cigar += "%dM" % (right + 1 - run_start)
        bases += seq[run_start - 1:right]
        reads.append(dict(tid=0, pos=left - 1, cigar=cigar, seq=bases, qual=[30] * len(bases), flag=16 * strand))
    return reads


def fasta_records(path):
    out = []
    for block in open(path).read().split(">")[1:]:
        name, seq = block.split("\n", 1)
        out.append((name.split()[0], seq.replace("\n", "")))
    return out


REBUILT_TABLES = [("per_read_group.no_read_groups.tab", "bull_1.fasta"), ("per_read_group.multiple_read_groups.tab", "lambda_split.fasta")]


def rebuilt_inputs(table, fasta_fixture, tmp_path):
    """(bam, fasta, region, path of the table the reference suite expects) with the BAM rebuilt from that table."""
    import minibam
    want = os.path.join(helpers.GOLDEN, "reference_tests", "bam2cov", table)
    lines = [line.rstrip("\n").split("\t") for line in open(want)]
    n_groups = sum(1 for c in lines[0] if c.endswith("_unique_top_cov") and c.startswith("RG-"))
    rows = [dict() for _ in range(n_groups)]
    for c in lines[1:]:
        if c[0].isdigit():
            assert [int(x) for x in c[2:10]] == [sum(int(c[10 + 8 * g + k]) for g in range(n_groups)) for k in range(8)]   # groups add up
            for g in range(n_groups):
                v = [int(x) for x in c[10 + 8 * g:18 + 8 * g]]
                assert v[2:6] == [0, 0, 0, 0]                        # no redundant reads in these windows
                rows[g][int(c[0])] = v
    first, last = min(rows[0]), max(rows[0])
    contigs = fasta_records(os.path.join(helpers.GOLDEN, "reference_tests", fasta_fixture))
    name, seq = contigs[0]
    groups = ["rg%d" % g for g in range(n_groups)] if n_groups > 1 else []
    reads = []
    for g in range(n_groups):
        for r in reads_from_table(rows[g], seq, first, last):
            if groups:
                r["tags"] = {"RG": groups[g]}
            reads.append(r)
    reads.sort(key=lambda r: r["pos"])
    assert any("D" in r["cigar"] for r in reads) or table != "per_read_group.no_read_groups.tab"   # the dip at 665 needs deletions
    bam, fasta = str(tmp_path / "rebuilt.bam"), str(tmp_path / "rebuilt.fasta")
    minibam.write(bam, fasta, contigs, reads, read_groups=groups)
    return bam, fasta, "%s:%d-%d" % (name, first, last), want


@pytest.mark.parametrize("table, fasta_fixture", REBUILT_TABLES)
def test_reference_suite_table_from_reads_rebuilt_out_of_it(table, fasta_fixture, checker, tmp_path):
    """/root/reference/tests/bam2cov_per_read_group/expected.*.tab (`breseq BAM2COV --format TSV --per-read-group`: rachael:657-767 of
    bull_1.bam, which has no @RG line, and NC_001416-0:8000-8110 of lambda_mult_ref_read's BAM with four read groups, two of them
    empty; neither BAM is in the checkout): reads rebuilt from the coverage and begin columns of each read group, put through the
    walk and the table writer, give the file back byte for byte -- header, the RG-<n> column sets, every row (the dips where
    reads have a base deleted included), the averages."""
    bam, fasta, region, want = rebuilt_inputs(table, fasta_fixture, tmp_path)
    out = str(tmp_path / "table.tab")
    p = subprocess.run([checker, bam, fasta, region, "600", "0", "0", out, "1"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert open(out).read() == open(want).read()


def fill_between_printed_rows(rows):
    """A thinned table prints every k-th position.  The rows in between are free: each gets the read begins (forward strand) or the
    read ends (reversed strand) that let the coverage move from one printed row to the next without deletions."""
    printed = sorted(rows)
    full = {printed[0]: rows[printed[0]]}
    for a, b in zip(printed, printed[1:]):
        for q in range(a + 1, b):
            full[q] = list(full[q - 1])
            full[q][6] = full[q][7] = 0
        if b - a > 1:
            q = b - 1
            rise = rows[b][0] - rows[b][6] - rows[a][0]            # forward: what begins at b is printed; the rest began before b
            if rise > 0:
                full[q][0] += rise
                full[q][6] = rise
            # reversed: reads that end at a leave behind a; what else is missing at b ended (= "began") on a row in between
            fall = (rows[a][1] - rows[a][7]) - rows[b][1]
            for k in range(a + 1, b):
                full[k][1] = rows[a][1] - rows[a][7]
            if fall > 0:
                full[q][7] = fall
        full[b] = rows[b]
    return full


THINNED_TABLES = [("show_average.tab", False), ("show_average.csv", True)]
REFERENCE_AVERAGE = 330.7552   # references.reference[rachael].coverage_average in the suite's summary.json


def thinned_inputs(table, csv, tmp_path):
    import minibam
    want = os.path.join(helpers.GOLDEN, "reference_tests", "bam2cov", table)
    rows = {}
    for line in open(want):
        c = line.rstrip("\n").split("," if csv else "\t")
        if c[0].isdigit():
            assert c[4:8] == ["0", "0", "0", "0"]
            rows[int(c[0])] = [int(c[2]), int(c[3]), 0, 0, 0, 0, int(c[8]), int(c[9])]
    first, last = min(rows), max(rows)
    assert (first, last, len(rows)) == (657, 2167, 756)
    full = fill_between_printed_rows(rows)
    name, seq = fasta_records(os.path.join(helpers.GOLDEN, "reference_tests", "bull_1.fasta"))[0]
    bam, fasta = str(tmp_path / "rebuilt.bam"), str(tmp_path / "rebuilt.fasta")
    minibam.write(bam, fasta, [(name, seq)], reads_from_table(full, seq, first, last))
    return bam, fasta, "%s:%d-%d" % (name, first, last), want


@pytest.mark.parametrize("table, csv", THINNED_TABLES)
def test_reference_suite_thinned_tables_with_the_reference_average(table, csv, checker, tmp_path):
    """/root/reference/tests/bam2cov/expected.tab and bam2cov_csv/expected.csv (`BAM2COV -a -r rachael:657-2167` at the default
    resolution of 600: every second position of 1511, as a table and as CSV; -a adds the fit average of summary.json, 330.7552).
    Reads rebuilt from the printed rows (the rows in between are free), through the walk and the writer: byte for byte, i.e. the
    thinning rule, the averages over the printed rows only, number_of_positions, the -a line, the CSV form."""
    bam, fasta, region, want = thinned_inputs(table, csv, tmp_path)
    out = str(tmp_path / "table.out")
    p = subprocess.run([checker, bam, fasta, region, "600", "0", "1" if csv else "0", out, "0", repr(REFERENCE_AVERAGE)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert open(out).read() == open(want).read()
