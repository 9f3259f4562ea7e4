"""The N > 1 path on GPUs: a run sharded by reference range over NCCL ranks (one process per GPU, launched the way the driver
launches bench.py) must write the evidence file of the unsharded run, byte for byte: record-balanced range cuts, device
staging of every rank's own reads, the sum of both integer histograms over the ranks (an NCCL allreduce, or fused into pass 1:
csrc/exchange.cu), every rank's evidence share gathered to
rank 0 and walked together.  Needs two GPUs (skipped on a one-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py`)."""
import filecmp
import os
import subprocess
import sys

import pytest

import breseq_b200 as bq
import helpers

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
import breseq_b200 as bq, helpers
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
d = dict(helpers.DATASETS[%(name)r]); out = %(out)r
d["bam"], d["fasta"] = os.path.join(out, "reference.bam"), os.path.join(out, "reference.fasta")
ctx = bq.Context(device=local)
bounds = ctx.bam_shard_bounds(d["bam"], world)
ctx.stage_bam(d["bam"], d["fasta"], shard_bounds=(bounds[rank], bounds[rank + 1]), staging="device", **helpers.stage_kwargs(d))
depth = torch.tensor([ctx.max_coverage_depth()], dtype=torch.int64, device=dev)
dist.all_reduce(depth, op=dist.ReduceOp.MAX)
ctx.set_min_coverage_depth(int(depth.item()))
if not %(fused)r:
    ctx.error_count(helpers.covariates(d))
class DevArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}
if %(fused)r:
    # the collective fused into pass 1 (csrc/exchange.cu): handles exchanged once, then error_count() sums over the ranks itself;
    # called twice so that both inbox copies and their clearing are exercised
    handles = [None] * world
    dist.all_gather_object(handles, ctx.hist_exchange_export())
    ctx.hist_exchange_attach(handles, rank)
    dist.barrier()
    ctx.error_count(helpers.covariates(d))
    ctx.error_count(helpers.covariates(d))
    ctx.error_count(helpers.covariates(d))
else:
    c, n, v, m = ctx.hist_device()
    with torch.cuda.stream(torch.cuda.ExternalStream(ctx.cuda_stream(), device=dev)):
        assert v == c + 8 * n
        dist.all_reduce(torch.as_tensor(DevArray(c, n + m), device=dev))   # both histograms: one allocation, one collective
ctx.derive_error_table()
nt = len(d["contig_lens"])
if rank == 0:
    ctx.write_error_count_files(out, os.path.join(out, "error_rates.tab"), helpers.readfile_names(d))
ctx.score_columns(bq.Context.score_params(d["mutation_cutoff"], d["polymorphism_cutoff"], d["precision"], d["places"]))
blob = ctx.evidence_export([d["del_prop"]] * nt)
sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
dist.all_gather(sizes, torch.tensor([len(blob)], dtype=torch.int64, device=dev))
cap = max(int(x.item()) for x in sizes)
buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
buf[:len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
parts = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
dist.gather(buf, parts, dst=0)
if rank == 0:
    shares = [bytes(p[:int(s.item())].cpu().numpy().tobytes()) for p, s in zip(parts, sizes)]
    ctx.write_evidence_merged(os.path.join(out, "ra_mc_evidence.gd"), shares, [d["del_prop"]] * nt, [d["del_seed"]] * nt)
ctx.close()
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("fused", [False, True], ids=["nccl", "fused"])
@pytest.mark.parametrize("name", ["multi", "lambda"])
def test_two_nccl_ranks_write_the_unsharded_evidence(name, fused, datasets, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    d = datasets[name]
    out = str(tmp_path)
    for f in ("reference.bam", "reference.fasta"):
        os.symlink(os.path.join(d["dir"], f), os.path.join(out, f))
    script = os.path.join(out, "worker.py")
    open(script, "w").write(WORKER % dict(root=ROOT, name=name, out=out, fused=fused))
    port = 29600 + (os.getpid() % 2000)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), script], capture_output=True, text=True, timeout=150)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert filecmp.cmp(os.path.join(out, "ra_mc_evidence.gd"), d["oracle_gd"], shallow=False)
    assert filecmp.cmp(os.path.join(out, "error_rates.tab"), d["oracle_rates"], shallow=False)
    for g in range(len(d["contig_lens"])):
        nm = "%d.unique_only_coverage_distribution.tab" % g
        assert filecmp.cmp(os.path.join(out, nm), os.path.join(d["oracle_dir"], nm), shallow=False), nm
