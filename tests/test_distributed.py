"""The N > 1 path on CPU: two `gloo` ranks, each staging its own contiguous reference-coordinate shard
(brq_stage_options.shard_rank / shard_count), one sum-allreduce of the integer histograms -- the only
collective of the path -- then every rank holds the whole-genome table.  The kernels' part is played by
the numpy statements in helpers.py (no GPU here); the GPU version of the same flow is bench.py --gpus N."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import breseq_b200 as bq
import helpers

WORLD = 2


def _worker(rank, world, d, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = bq.Context(device=-1)
        ctx.stage_bam(d["bam"], d["fasta"], read_file_sets=helpers.read_file_sets(d), shard=(rank, world))
        s = ctx.stream()
        n_files = len(helpers.readfile_names(d))
        counts = torch.from_numpy(helpers.emulate_hist(s["hist_rec"], n_files, 42))
        cov = helpers.emulate_coverage_hist(s["hist_off"])
        cov_t = torch.zeros(4096, dtype=torch.int64)
        cov_t[:len(cov)] = torch.from_numpy(cov)
        geom = torch.tensor([int(s["n_base"]), int(s["n_score"]), int(s["n_hist"])], dtype=torch.int64)
        dist.all_reduce(counts)   # the one collective of the path
        dist.all_reduce(cov_t)
        dist.all_reduce(geom)
        t = helpers.emulate_tally(s)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), counts=counts.numpy(), cov=cov_t.numpy(), geom=geom.numpy(),
                 unique=t["unique"], n=t["n"], n_base=int(s["n_base"]), n_ins=int(s["n_ins"]))
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["lambda", "multi"])
def test_two_rank_shards_allreduce_to_the_whole_genome_table(name, datasets, tmp_path):
    d = datasets[name]
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(WORLD, d, port, str(tmp_path)), nprocs=WORLD, join=True)
    r = [np.load(str(tmp_path / ("rank%d.npz" % k))) for k in range(WORLD)]
    # every rank ends with the same, whole-genome histogram: bit-exact against the oracle
    ora = helpers.oracle_counts(d["oracle_counts"])
    for k in range(WORLD):
        assert np.array_equal(r[k]["counts"], ora)
        assert np.array_equal(r[k]["cov"], r[0]["cov"])
    # the shards partition the columns and the records
    full = bq.Context(device=-1)
    full.stage_bam(d["bam"], d["fasta"], read_file_sets=helpers.read_file_sets(d))
    sf = full.stream()
    assert list(r[0]["geom"]) == [int(sf["n_base"]), int(sf["n_score"]), int(sf["n_hist"])]
    assert sum(int(x["n_base"]) for x in r) == int(sf["n_base"])
    # per-column tallies of the shards, concatenated in rank order, are the unsharded ones (base columns)
    tf = helpers.emulate_tally(sf)
    nb = int(sf["n_base"])
    cat_unique = np.concatenate([x["unique"][:int(x["n_base"])] for x in r])
    cat_n = np.concatenate([x["n"][:int(x["n_base"])] for x in r])
    assert np.array_equal(cat_unique, tf["unique"][:nb]) and np.array_equal(cat_n, tf["n"][:nb])
    full.close()


def test_record_balanced_shards_from_windowed_reads():
    """bench.py's N > 1 input path: cut points with equal aligned bases (balance by record count), each rank
    generating only the reads that can overlap its range.  Every shard staged from its windowed reads must be the shard
    staged from the whole read set, array by array, and the shards must balance better than an even split by columns."""
    d = dict(helpers.DATASETS["multi"])
    d["n_gaps"] = 3
    world = 3
    ctx = bq.Context(device=-1)
    spec = helpers.synth_spec(d)
    bounds = ctx.synth_shard_bounds(spec, world)
    assert bounds[0] == 0 and bounds[-1] == sum(d["contig_lens"]) and bounds == sorted(bounds)
    records = []
    for rank in range(world):
        lo, hi = bounds[rank], bounds[rank + 1]
        ctx.stage_synthetic(spec, read_file_sets=helpers.read_file_sets(d), shard_bounds=(lo, hi))
        a = ctx.stream()
        a = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in a.items()}
        win = bq.SynthSpec(seed=d["seed"], read_sets=d["read_sets"], contig_lens=d["contig_lens"], contig_prefix=d["prefix"],
                           n_polymorphic=d["n_polymorphic"], n_fixed=d["n_fixed"], n_gaps=d["n_gaps"], window=(lo, hi))
        w = bq.Context(device=-1)
        w.stage_synthetic(win, read_file_sets=helpers.read_file_sets(d), shard_bounds=(lo, hi))
        b = w.stream()
        assert b["n_reads"] < a["n_reads"]
        for k in ("score_rec", "score_off", "score_cnt", "side_rec", "side_off", "hist_rec", "hist_off", "round_slot", "slot_ref",
                  "ins_parent", "ins_count"):
            assert np.array_equal(a[k], b[k]), k
        records.append(int(a["n_score"]))
        w.close()
    ctx.close()
    assert max(records) < 1.1 * (sum(records) / world), records


class _FakeExchangeContext:
    """Stands in for a device context in bench.setup_fused_exchange: records what it was asked, fails where told to."""

    def __init__(self, rank, fail_export=False, fail_attach=False):
        self.rank, self.fail_export, self.fail_attach, self.calls = rank, fail_export, fail_attach, []

    def hist_exchange_export(self):
        if self.fail_export:
            raise bq.BrqError("cudaIpcGetMemHandle: not permitted")
        return bytes([self.rank]) * 64

    def hist_exchange_attach(self, handles, rank):
        self.calls.append((len(handles), rank))
        if self.fail_attach and len(handles) > 1:
            raise bq.BrqError("cudaIpcOpenMemHandle: peer access is not supported")
        assert all(len(h) == 64 for h in handles)


def _exchange_worker(rank, world, port, out_dir, failing_rank, how):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, helpers.ROOT)
        import bench
        ctx = _FakeExchangeContext(rank, fail_export=(how == "export" and rank == failing_rank), fail_attach=(how == "attach" and rank == failing_rank))
        on = bench.setup_fused_exchange(ctx, dist, rank, world)
        with open(os.path.join(out_dir, "rank%d.txt" % rank), "w") as f:
            f.write("%d %r\n" % (on, ctx.calls))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("how", ["none", "export", "attach"])
def test_ranks_agree_on_the_fused_exchange_or_all_fall_back(how, tmp_path):
    """bench.py --gpus N: the ranks attach to each other's inboxes for the fused histogram exchange, or -- if any one of them cannot
    -- every rank detaches and reports that the NCCL allreduce is to be used.  No rank is left waiting for the others."""
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_exchange_worker, args=(WORLD, port, str(tmp_path), 1, how), nprocs=WORLD, join=True)
    got = [open(str(tmp_path / ("rank%d.txt" % k))).read().split(" ", 1) for k in range(WORLD)]
    assert [g[0] for g in got] == (["1", "1"] if how == "none" else ["0", "0"])
    calls = [eval(g[1]) for g in got]
    if how == "none":
        assert calls == [[(2, 0)], [(2, 1)]]
    elif how == "export":
        assert calls == [[], []]                         # nobody attached
    else:
        assert calls == [[(2, 0), (1, 0)], [(2, 1)]]     # rank 0 had attached and detaches again
