"""Device staging (csrc/expand.cu: the reads cross PCIe, kernels expand the CIGARs, classify the records and lay out the
streams in HBM) against host staging (csrc/staging.cpp) on a B200: every array of the stream must be equal, bit for bit."""
import numpy as np
import pytest

import breseq_b200 as bq
import helpers

pytestmark = pytest.mark.gpu

ARRAYS = ["score_rec", "score_off", "score_cnt", "side_rec", "side_off", "slot_ref", "ins_parent", "ins_count", "round_slot",
          "round_off", "hist_rec", "hist_off", "hist16", "hist_exc"]
SCALARS = ["n_base", "n_ins", "n_score", "n_hist", "n_score_padded", "n_side", "side_stride", "geometry", "n_hist16", "n_hist_exc", "n_rounds"]


def staged(d, staging, **kw):
    ctx = bq.Context(device=0)
    args = dict(helpers.stage_kwargs(d))
    args.update(kw)
    ctx.stage_bam(d["bam"], d["fasta"], staging=staging, **args)
    s = ctx.stream()
    out = {k: (None if s[k] is None else np.array(s[k], copy=True)) for k in ARRAYS}
    out.update({k: s[k] for k in SCALARS})
    out["device_built"] = s["device_built"]
    return ctx, out


def compare(h, d):
    assert not h["device_built"] and d["device_built"]
    for k in SCALARS:
        assert h[k] == d[k], k
    for k in ARRAYS:
        if h[k] is None:
            assert d[k] is None, k
            continue
        assert d[k] is not None and h[k].shape == d[k].shape, k
        bad = np.flatnonzero(h[k] != d[k])
        assert len(bad) == 0, "%s differs at %d of %d entries, first at %d: host %x device %x" % (
            k, len(bad), len(h[k]), bad[0], int(h[k][bad[0]]), int(d[k][bad[0]]))


@pytest.mark.parametrize("name", list(helpers.DATASETS))
def test_device_built_stream_equals_host_staging(name, datasets):
    d = datasets[name]
    ch, h = staged(d, "host")
    cd, dv = staged(d, "device")
    compare(h, dv)
    ch.close()
    cd.close()


@pytest.mark.parametrize("rank", [0, 1, 2])
def test_device_built_shards(rank, datasets):
    d = datasets["multi"]
    ch, h = staged(d, "host", shard=(rank, 3))
    cd, dv = staged(d, "device", shard=(rank, 3))
    compare(h, dv)
    ch.close()
    cd.close()


def test_preprocess_read_starts_on_the_device(datasets):
    d = datasets["tiny"]
    out = []
    for staging in ("host", "device"):
        ctx = bq.Context(device=0)
        ctx.stage_bam(d["bam"], d["fasta"], staging=staging, preprocess_stage=True, **helpers.stage_kwargs(d))
        out.append(ctx.preprocess_read_starts())
        ctx.close()
    assert np.array_equal(out[0], out[1])


def test_device_staging_needs_a_device():
    ctx = bq.Context(device=-1)
    d = helpers.DATASETS["tiny"]
    spec = helpers.synth_spec(d)
    with pytest.raises(bq.BrqError, match="device staging needs"):
        ctx.stage_synthetic(spec, staging="device")
    ctx.close()


def test_hand_derived_cigar_cases_on_the_device(tmp_path):
    """The known-answer reads of tests/test_pileup_semantics.py (soft / hard clips, deletions, insertions with padding,
    reference skips, flagged reads) through the device expander: the same streams as host staging."""
    import minibam
    import test_pileup_semantics as kat
    for name, (read, _, _) in sorted(kat.CASES.items()):
        bam, fasta = str(tmp_path / (name + ".bam")), str(tmp_path / (name + ".fasta"))
        minibam.write(bam, fasta, [("chr", kat.REF)], [read])
        d = dict(bam=bam, fasta=fasta, read_sets=[])
        out = []
        for staging in ("host", "device"):
            ctx = bq.Context(device=0)
            ctx.stage_bam(bam, fasta, staging=staging)
            s = ctx.stream()
            o = {k: (None if s[k] is None else np.array(s[k], copy=True)) for k in ARRAYS}
            o.update({k: s[k] for k in SCALARS})
            o["device_built"] = s["device_built"]
            out.append(o)
            ctx.close()
        compare(out[0], out[1])
