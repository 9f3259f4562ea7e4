"""Known-answer cases worked out BY HAND from the reference's source (not from the oracle):
what error_count bins a read with a given CIGAR increments, and which coverage each column gets.
Both the oracle (htslib-shim pileup) and the product's read-centric staging must reproduce them.

Bin naming follows error_count.cpp:854-986: (ref, obs, quality) on the READ strand; '..' is the
"next base also aligned" observation, 'X.' a one-base deletion, '.X' a one-base insertion.
"""
import os

import numpy as np
import pytest

import breseq_b200 as bq
import helpers
import minibam

Q = 42
B = {"A": 0, "C": 1, "G": 2, "T": 3, ".": 4}


def bin_index(ref, obs, q, n_sets=1, read_set=0):
    return read_set + B[ref] * n_sets + B[obs] * 5 * n_sets + q * 25 * n_sets


def expected(bins):
    c = np.zeros(25 * Q, dtype=np.int64)
    for ref, obs, q in bins:
        c[bin_index(ref, obs, q)] += 1
    return c


REF = "ACGTACGTACGTTTGACCA"

CASES = {
    # forward 4M: four matches, three '..' with the quality of the NEXT base
    "forward_match": (dict(tid=0, pos=0, cigar="4M", seq="ACGT", qual=[30, 31, 32, 33]),
                      [("A", "A", 30), ("C", "C", 31), ("G", "G", 32), ("T", "T", 33), (".", ".", 31), (".", ".", 32), (".", ".", 33)],
                      {1: 4}),
    # reverse strand: both bases complemented; '..' uses the quality of THIS base (mqpos = q + 1 - 1)
    "reverse_match": (dict(tid=0, pos=0, cigar="4M", seq="ACGT", qual=[30, 31, 32, 33], flag=16),
                      [("T", "T", 30), ("G", "G", 31), ("C", "C", 32), ("A", "A", 33), (".", ".", 30), (".", ".", 31), (".", ".", 32)],
                      {1: 4}),
    # mismatch at the second base: reference C read as T
    "mismatch": (dict(tid=0, pos=0, cigar="3M", seq="ATG", qual=[20, 21, 22]),
                 [("A", "A", 20), ("C", "T", 21), ("G", "G", 22), (".", ".", 21), (".", ".", 22)], {1: 3}),
    # 2M1D2M: the deleted column is skipped by pass 1; the base before it reports (next ref base, '.')
    # with the quality of the first read base after the deletion
    "one_base_deletion": (dict(tid=0, pos=0, cigar="2M1D2M", seq="ACTA", qual=[30, 31, 32, 33]),
                          [("A", "A", 30), ("C", "C", 31), ("T", "T", 32), ("A", "A", 33), (".", ".", 31), ("G", ".", 32), (".", ".", 33)],
                          {1: 4}),
    # a two-base deletion produces NO second observation at the base before it (indel == -2)
    "two_base_deletion": (dict(tid=0, pos=0, cigar="2M2D2M", seq="ACAC", qual=[30, 31, 32, 33]),
                          [("A", "A", 30), ("C", "C", 31), ("A", "A", 32), ("C", "C", 33), (".", ".", 31), (".", ".", 33)], {1: 4}),
    # 2M1I2M: the base before the insertion reports ('.', inserted base) with the inserted base's quality
    "one_base_insertion": (dict(tid=0, pos=0, cigar="2M1I2M", seq="ACTGT", qual=[30, 31, 32, 33, 34]),
                           [("A", "A", 30), ("C", "C", 31), ("G", "G", 33), ("T", "T", 34), (".", ".", 31), (".", "T", 32), (".", ".", 34)],
                           {1: 4}),
    # soft clips shift the query index and bound the '..' observation at the last ALIGNED base
    "soft_clips": (dict(tid=0, pos=4, cigar="2S3M1S", seq="TTACGA", qual=[10, 11, 30, 31, 32, 12]),
                   [("A", "A", 30), ("C", "C", 31), ("G", "G", 32), (".", ".", 31), (".", ".", 32)], {1: 3}),
    # an N in the read: no substitution observation there, and no '..' pointing at it
    "n_in_read": (dict(tid=0, pos=0, cigar="4M", seq="ANGT", qual=[30, 2, 32, 33]),
                  [("A", "A", 30), ("G", "G", 32), ("T", "T", 33), (".", ".", 32), (".", ".", 33)], {1: 4}),
    # a padding operation inside an insertion run is skipped: 2M1I1P1I2M reports indel = +2 at the base before it, which is
    # not a one-base insertion, so that base has NO second observation (error_count.cpp:963: indel == +1 only)
    "pad_inside_insertion": (dict(tid=0, pos=0, cigar="2M1I1P1I2M", seq="ACTTGT", qual=[30, 31, 32, 33, 34, 35]),
                             [("A", "A", 30), ("C", "C", 31), ("G", "G", 34), ("T", "T", 35), (".", ".", 31), (".", ".", 35)], {1: 4}),
    # a reference skip (N) is a deleted column for the pileup (is_del): pass 1 skips it; the base before it sees indel == 0
    # (only a D after a non-D operation is reported), so it reports '..' with the quality of the next read base
    "reference_skip": (dict(tid=0, pos=0, cigar="2M2N2M", seq="ACAC", qual=[30, 31, 32, 33]),
                       [("A", "A", 30), ("C", "C", 31), ("A", "A", 32), ("C", "C", 33), (".", ".", 31), (".", ".", 32), (".", ".", 33)], {1: 4}),
    # hard clips consume nothing: the same bins as forward_match
    "hard_clips": (dict(tid=0, pos=0, cigar="3H4M2H", seq="ACGT", qual=[30, 31, 32, 33]),
                   [("A", "A", 30), ("C", "C", 31), ("G", "G", 32), ("T", "T", 33), (".", ".", 31), (".", ".", 32), (".", ".", 33)],
                   {1: 4}),
}
# reads the pileup engine keeps although samtools' defaults would drop them: bam_plp_push only tests BAM_FUNMAP
# (oracle/hts_shim/hts_shim.cpp header; csrc/expand_core.h PILEUP_FLAG_MASK)
for _flag, _name in ((256, "secondary"), (512, "qc_fail"), (1024, "duplicate")):
    CASES["flagged_" + _name] = (dict(tid=0, pos=0, cigar="4M", seq="ACGT", qual=[30, 31, 32, 33], flag=_flag),
                                 [("A", "A", 30), ("C", "C", 31), ("G", "G", 32), ("T", "T", 33), (".", ".", 31), (".", ".", 32), (".", ".", 33)],
                                 {1: 4})
# ... and the one it drops
CASES["flagged_unmapped"] = (dict(tid=0, pos=0, cigar="4M", seq="ACGT", qual=[30, 31, 32, 33], flag=4), [], {})


def run_case(tmp_path, name, reads):
    bam, fasta = str(tmp_path / (name + ".bam")), str(tmp_path / (name + ".fasta"))
    minibam.write(bam, fasta, [("chr", REF)], reads)
    out = str(tmp_path / (name + "_oracle"))
    os.makedirs(out, exist_ok=True)
    counts = os.path.join(out, "counts.tab")
    helpers.run_oracle("error_count", "--bam", bam, "--fasta", fasta, "--out", out, "--covariates",
                       "read_set=1,obs_base,ref_base,quality=%d" % Q, "--readfiles", "r", "--counts-dump", counts)
    ora = helpers.oracle_counts(counts)
    cov = {}
    for l in open(os.path.join(out, "0.unique_only_coverage_distribution.tab")).read().strip().split("\n")[1:]:
        j, c = l.split("\t")
        if int(c):
            cov[int(j)] = int(c)
    ctx = bq.Context(device=-1)
    ctx.stage_bam(bam, fasta)
    s = ctx.stream()
    mine = helpers.emulate_hist(s["hist_rec"], 1, Q)
    mine_cov = helpers.emulate_coverage_hist(s["hist_off"])
    ctx.close()
    return ora, cov, mine, {j: int(c) for j, c in enumerate(mine_cov) if j > 0 and c}


@pytest.mark.parametrize("name", sorted(CASES))
def test_single_read_known_answers(built, tmp_path, name):
    read, bins, cov = CASES[name]
    ora, ora_cov, mine, mine_cov = run_case(tmp_path, name, [read])
    want = expected(bins)
    assert np.array_equal(ora, want), "oracle disagrees with the hand-derived bins"
    assert np.array_equal(mine, want), "staging disagrees with the hand-derived bins"
    assert ora_cov == cov and mine_cov == cov


def test_redundant_read_excludes_columns_from_coverage(built, tmp_path):
    reads = [dict(tid=0, pos=0, cigar="4M", seq="ACGT", qual=[30] * 4),
             dict(tid=0, pos=2, cigar="4M", seq="GTAC", qual=[30] * 4, tags={"X1": 3})]
    ora, ora_cov, mine, mine_cov = run_case(tmp_path, "redundant", reads)
    # only the unique read is counted; columns 3-4 (1-based) carry a redundant read and are left out
    # of the coverage histogram, columns 5-6 have no unique read at all but a redundant one
    assert ora.sum() == 7 and np.array_equal(ora, mine)
    assert ora_cov == {1: 2} and mine_cov == {1: 2}


def test_pass2_coverage_on_deleted_and_inserted_columns(built, tmp_path):
    """identify_mutations counts a deletion-spanning read as unique coverage (no is_del skip before
    identify_mutations.cpp:1591) and opens one sub-column per inserted base of a unique read."""
    reads = [dict(tid=0, pos=0, cigar="2M1D3M", seq="ACTAC", qual=[30, 31, 32, 33, 34]),
             dict(tid=0, pos=0, cigar="3M2I2M", seq="ACGAATA", qual=[30] * 7, flag=16)]
    bam, fasta = str(tmp_path / "p2.bam"), str(tmp_path / "p2.fasta")
    minibam.write(bam, fasta, [("chr", REF)], reads)
    out = str(tmp_path / "p2_oracle")
    os.makedirs(out, exist_ok=True)
    helpers.run_oracle("error_count", "--bam", bam, "--fasta", fasta, "--out", out, "--covariates",
                       "read_set=1,obs_base,ref_base,quality=%d" % Q, "--readfiles", "r")
    cols = os.path.join(out, "cols.bin")
    helpers.run_oracle("identify_mutations", "--bam", bam, "--fasta", fasta, "--error-rates", os.path.join(out, "error_rates.tab"),
                       "--gd", os.path.join(out, "o.gd"), "--del-prop", "0", "--del-seed", "0", "--columns-out", cols)
    o = helpers.oracle_columns(cols)
    rows = {(int(r["pos1"]), int(r["insert_count"])): r for r in o}
    assert len(o) == len(REF) + 2 and (3, 1) in rows and (3, 2) in rows
    # column 3: read 1 is deleted there (top strand), read 2 has a G (bottom strand); both count
    assert list(rows[(3, 0)]["unique"]) == [1.0, 1.0] and rows[(3, 0)]["n"] == 2
    # the sub-columns see '.' from read 1 and the inserted bases from read 2
    assert list(rows[(3, 1)]["unique"]) == [1.0, 1.0] and rows[(3, 2)]["n"] == 2
    ctx = bq.Context(device=-1)
    ctx.stage_bam(bam, fasta)
    s = ctx.stream()
    slot = helpers.oracle_slots(o, s, np.zeros(1, dtype=np.int64))
    t = helpers.emulate_tally(s)
    assert np.array_equal(t["unique"][slot], o["unique"].astype(np.int64))
    assert np.array_equal(t["n"][slot], o["n"].astype(np.int64))
    # observed bases at column 3: '.' (4) from the deleted read, G (2) from the other
    r = bq.decode_score_records(s)
    assert sorted(r["obs"][r["slot"] == 2].tolist()) == [2, 4]
    ctx.close()
