"""Host staging against the oracle, without a GPU: the staged streams must carry exactly what the
reference counts.  The numpy re-statements in helpers.py play the kernels' part."""
import os

import numpy as np
import pytest

import breseq_b200 as bq
import helpers


@pytest.fixture(scope="module", params=list(helpers.DATASETS))
def staged(request, datasets):
    d = datasets[request.param]
    ctx = bq.Context(device=-1)
    ctx.stage_bam(d["bam"], d["fasta"], **helpers.stage_kwargs(d))
    yield d, ctx, ctx.stream()
    ctx.close()


def test_histogram_stream_matches_oracle_counts(staged):
    d, ctx, s = staged
    if d.get("covariates"):
        pytest.skip("the numpy emulation knows the default covariate layout only; the GPU suite checks these counts")
    mine = helpers.emulate_hist(s["hist_rec"], len(helpers.readfile_names(d)), 42)
    ora = helpers.oracle_counts(d["oracle_counts"])
    assert mine.sum() == ora.sum()
    assert np.array_equal(mine, ora)


def test_compact_histogram_stream_is_the_same_records(staged):
    """hist16 + hist_exc (what the device reads) hold exactly the records of hist_rec, as a multiset."""
    d, ctx, s = staged
    if s["hist_rec"].dtype == np.uint64:
        assert s["hist16"] is None  # 8-byte records (read_pos / base_repeat) have no compact form
        return
    assert s["hist16"] is not None and len(s["hist16"]) + len(s["hist_exc"]) == len(s["hist_rec"])
    both = np.concatenate([helpers.expand_hist16(s["hist16"]), s["hist_exc"]])
    assert np.array_equal(np.sort(both), np.sort(s["hist_rec"]))
    if len(s["hist_rec"]):  # the exceptions are the few records that are not `fast` (or do not fit 6-bit qualities / 2-bit sets)
        assert len(s["hist_exc"]) <= 0.05 * len(s["hist_rec"]) + 64
    assert np.all((s["hist_exc"] >> 31 == 0) | (((s["hist_exc"] >> 6) & 127) > 62) | ((((s["hist_exc"] >> 20) & 127) > 62) & (((s["hist_exc"] >> 20) & 127) != 127)) | (((s["hist_exc"] >> 28) & 7) > 3))


def test_score_transfer_form_expands_to_the_stream(staged):
    """score16 + score_exc (what crosses PCIe) rebuild score_rec bit for bit; the exceptions are the few redundant and
    mismatching records."""
    d, ctx, s = staged
    assert s["score16"] is not None and len(s["score16"]) == len(s["score_rec"])
    assert np.array_equal(helpers.expand_score16(s), s["score_rec"])
    assert int(s["score_exc_off"][-1]) == len(s["score_exc"])
    if d["name"] != "ltee" and len(s["score_rec"]):
        assert len(s["score_exc"]) <= 0.15 * int(s["n_score"]) + 64


def test_unique_only_coverage_matches_oracle(staged):
    d, ctx, s = staged
    names = helpers.contig_names(d)
    slot0 = helpers.visit_slot0(names, d["contig_lens"])
    for tid, length in enumerate(d["contig_lens"]):
        lo = int(slot0[tid])
        mine = helpers.emulate_coverage_hist(s["hist_off"][lo:lo + length + 1])
        lines = open("%s/%d.unique_only_coverage_distribution.tab" % (d["oracle_dir"], tid)).read().strip().split("\n")[1:]
        ora = np.zeros(len(lines) + 1, dtype=np.int64)
        for l in lines:
            j, c = l.split("\t")
            ora[int(j)] = int(c)
        n = max(len(mine), len(ora))
        mine = np.pad(mine, (0, n - len(mine)))
        ora = np.pad(ora, (0, n - len(ora)))
        assert np.array_equal(mine[1:], ora[1:])  # the reference never prints depth 0


def test_score_stream_tallies_match_oracle(staged):
    d, ctx, s = staged
    o = helpers.oracle_columns(d["oracle_columns"])
    assert len(o) == s["n_base"] + s["n_ins"], "insert sub-columns differ"
    slot = helpers.oracle_slots(o, s, helpers.visit_slot0(helpers.contig_names(d), d["contig_lens"]))
    assert len(set(slot.tolist())) == len(o)
    t = helpers.emulate_tally(s)
    assert np.array_equal(t["unique"][slot], o["unique"].astype(np.int64))
    assert np.array_equal(t["raw_redundant"][slot], o["raw_redundant"].astype(np.int64))
    assert np.array_equal(t["n"][slot], o["n"].astype(np.int64))
    assert np.array_equal(t["redundant"][slot], o["redundant"])  # bit-exact: same order, same divisions
    assert np.array_equal(s["slot_ref"][slot], o["ref"])


def test_redundant_records_lead_each_slot(staged):
    """Stream layout the scoring kernel relies on: round-major, lane-interleaved; within a slot, redundant records first."""
    d, ctx, s = staged
    rec = s["score_rec"]
    beg, cnt = bq.slot_ranges(s)
    n_slots = len(cnt)
    # rounds: every slot sits in exactly one lane of one round; a round holds one reference base
    order = s["round_slot"].astype(np.int64).reshape(-1, 32)
    used = order[order != 0xFFFFFFFF]
    assert np.array_equal(np.sort(used), np.arange(n_slots))
    ref = np.minimum(s["slot_ref"], 4).astype(np.int64)
    for row in order:
        live = row[row != 0xFFFFFFFF]
        assert len(set(ref[live])) <= 1
    # geometry of the stream: slot of lane l of round r starts at round_off[r] + 4 l; the round is as long as its deepest slot
    roff = s["round_off"].astype(np.int64)
    r_idx, l_idx = np.nonzero(order != 0xFFFFFFFF)
    assert np.array_equal(beg[order[r_idx, l_idx]], roff[r_idx] + 4 * l_idx)
    vecs = (cnt + 7) // 8
    deepest = np.zeros(len(order), np.int64)
    np.maximum.at(deepest, r_idx, vecs[order[r_idx, l_idx]])
    assert np.array_equal(np.diff(roff), deepest * 256) and roff[-1] == s["n_score_padded"] == len(rec)
    pos = bq.record_positions(s)
    assert len(np.unique(pos)) == len(pos) == s["n_score"], "no two records share a word"
    real = np.zeros(len(rec), bool)
    real[pos] = True
    g = s["geometry"]
    trash = g["n_st"] * g["n_q"] + 6   # pad words count into the trash counter after the class counters, no other bit
    pad = (trash >> 2) * 128 + (trash & 3) * 8
    assert g["n_q"] % 4 == 0 and g["words"] == g["n_st"] * g["n_q"] // 4 + 2 and g["words"] <= 64
    assert np.all(rec[~real] == pad) and np.all(rec[real] != pad), "pad words address the trash counter; records never equal them"
    # within a slot: redundant records (kind 3) first
    kind = rec[pos] >> 30
    slot = np.repeat(np.arange(n_slots), cnt)
    red = (kind == 3).astype(np.int64)
    first = np.cumsum(cnt) - cnt
    n_red = np.add.reduceat(red, first[cnt > 0]) if len(red) else np.zeros(0, np.int64)
    j = np.arange(len(pos)) - np.repeat(first, cnt)
    lead = np.zeros(n_slots, np.int64)
    lead[cnt > 0] = n_red
    assert np.array_equal(red == 1, j < lead[slot]), "redundant records lead their slot"
    assert red.sum() > 0
    # counters: a matching HOT record counts in its class, everything else in the special counters after the classes
    hot_match = (kind == 0) & (((rec[pos] >> 28) & 1) == 1)
    sq = (rec[pos] >> 16) & 0xFF
    counter = rec[pos] & 0x1FFF
    assert np.array_equal(counter[hot_match], (sq[hot_match] >> 2) * 128 + (sq[hot_match] & 3) * 8)
    special = (counter >> 7) * 4 + ((counter & 31) >> 3) - g["n_st"] * g["n_q"]
    assert np.all((special[~hot_match] >= 0) & (special[~hot_match] <= 6))


def test_shards_partition_the_stream(datasets):
    d = datasets["multi"]
    full = bq.Context(device=-1)
    full.stage_bam(d["bam"], d["fasta"], read_file_sets=helpers.read_file_sets(d))
    sf = full.stream()
    whole = helpers.emulate_hist(sf["hist_rec"], 3, 42)
    acc = np.zeros_like(whole)
    n_base = n_score = 0
    for rank in range(3):
        c = bq.Context(device=-1)
        c.stage_bam(d["bam"], d["fasta"], read_file_sets=helpers.read_file_sets(d), shard=(rank, 3))
        s = c.stream()
        acc += helpers.emulate_hist(s["hist_rec"], 3, 42)
        n_base += s["n_base"]
        n_score += s["n_score"]
        c.close()
    assert n_base == sf["n_base"] and n_score == sf["n_score"]
    assert np.array_equal(acc, whole)
    full.close()


def test_empty_and_subset_targets(datasets):
    d = datasets["multi"]
    names = helpers.contig_names(d)
    c = bq.Context(device=-1)
    c.stage_bam(d["bam"], d["fasta"], read_file_sets=helpers.read_file_sets(d), seq_ids=[names[1]])
    s = c.stream()
    assert s["n_base"] == d["contig_lens"][1]
    with pytest.raises(bq.BrqError):
        c.stage_bam(d["bam"], d["fasta"], seq_ids=["no_such_contig"])
    c.close()


def test_compact_histogram_stream_routes_wide_read_sets_to_exceptions(tmp_path):
    """Five read files: read sets 4 and above do not fit the two bits of a 16-bit record, so those records travel as
    exceptions; together the two streams are still exactly the positional records, and count like the oracle."""
    d = dict(seed=17, contig_lens=[1500], prefix="five",
             read_sets=[dict(name="a", paired=True, read_len=60, coverage=12.0, frag_mean=180, frag_sd=15),
                        dict(name="b", paired=True, read_len=60, coverage=12.0, frag_mean=180, frag_sd=15),
                        dict(name="c", paired=False, read_len=36, coverage=10.0)],
             n_polymorphic=4, n_fixed=2, n_gaps=1, mutation_cutoff=10.0, polymorphism_cutoff=2.0, precision=1e-6, places=8,
             del_prop=5.0, del_seed=0.0)
    d["bam"], d["fasta"] = str(tmp_path / "reference.bam"), str(tmp_path / "reference.fasta")
    ctx = bq.Context(device=-1)
    ctx.synth_write(helpers.synth_spec(d), d["bam"], d["fasta"])
    ctx.stage_bam(d["bam"], d["fasta"], read_file_sets=helpers.read_file_sets(d))
    s = ctx.stream()
    h, h16, exc = s["hist_rec"], s["hist16"], s["hist_exc"]
    assert h.dtype == np.uint32 and h16 is not None
    assert np.array_equal(np.sort(np.concatenate([helpers.expand_hist16(h16), exc])), np.sort(h))
    sets = (h >> 28) & 7
    assert sets.max() == 4 and np.all(helpers.expand_hist16(h16) >> 28 & 7 <= 3)
    fast_wide = ((h >> 31) == 1) & (sets > 3)
    assert fast_wide.sum() > 0 and np.count_nonzero(((exc >> 31) == 1) & (((exc >> 28) & 7) > 3)) == fast_wide.sum()
    # the covariate histogram of both streams together is the oracle's
    out = str(tmp_path / "oracle")
    os.makedirs(out)
    ec, _ = helpers.cli_args(d, out)
    dump = os.path.join(out, "counts.tab")
    helpers.run_oracle(*ec, "--counts-dump", dump)
    mine = helpers.emulate_hist(np.concatenate([helpers.expand_hist16(h16), exc]), 5, 42)
    assert np.array_equal(mine, helpers.oracle_counts(dump))
    ctx.close()
