#!/usr/bin/env python
"""bench.py -- aligned bases/sec of error_count + identify_mutations on B200.

  python bench.py --gpus 1 --steps 10 --warmup 3          (one JSON line on stdout)
  torchrun ... bench.py --gpus N ...                        (one rank per GPU; strong scaling)
  python bench.py --impl reference ...                      (the reference's own CPU code, same metric)
  python bench.py --config c1                               (BASELINE configs[1] instead of configs[2])

Workload (BASELINE.json configs[2], SURVEY.md 8d C2, the configuration the north star quotes its target on): ONE E. coli
REL606-sized reference (4 629 812 bp, one contig), synthetic 1000x pe150 reads of one paired read set (read_set=2, Q=42),
population / polymorphism mode cutoffs (settings.cpp:858-913: mutation 10, polymorphism 2, precision 1e-6, 8 places).
At N GPUs the genome is cut into N contiguous reference-coordinate ranges with the same number of aligned bases (prefix
sum of the read starts), one per rank: STRONG scaling.  The only data-path collective is the sum-allreduce
of the integer covariate and coverage histograms; the MC / UN intervals and the evidence ids cross the range boundaries,
so every rank exports its share of the evidence and rank 0 walks the shares together and writes ONE ra_mc_evidence.gd.

A "step" is one pass of both kernels' path over the resident streams:
  covariate histogram + coverage histogram -> allreduce -> table derivation + text canonicalisation + likelihood tables
  -> per-slot scoring (tally + fit).
`value` times that with the streams already in HBM.  `e2e` is the same job through the C ABI from HOST buffers (the decoded
reads of the rank's range in page-locked host memory): H2D of the reads, CIGAR expansion and record classification on the
device (csrc/expand.cu), both passes, the allreduce, D2H of histograms / walk events / flagged slots, the host
finalisation, the gather of the evidence shares and the files (error_rates.tab, base_qual_error_prob.*.tab, coverage
distribution, ra_mc_evidence.gd) written by rank 0.  `e2e_from_bam` goes one step further back, to where the reference's
arm starts: a BAM on disk (BASELINE configs[1]'s, 4.6 Mb at 100x: the largest whose file this run writes in seconds)
through brq_run_error_count + brq_run_identify_mutations to the same files on disk, on rank 0's GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GENOME = 4629812
COVARIATES = "read_set=2,obs_base,ref_base,quality=42"
CONFIGS = {
    # name: label, depth, (mutation, polymorphism, precision, places), deletion propagation cutoff, CPU sample divisors
    "c2": dict(label="E. coli REL606 4.6 Mb population/polymorphism mode, synthetic 1000x pe150 (BASELINE configs[2])",
               coverage=1000.0, cutoffs=(10.0, 2.0, 1e-6, 8), del_prop=300.0, cpu_div=256, cpu_baseline_div=128),
    "c1": dict(label="E. coli REL606 4.6 Mb clone mode, synthetic 100x pe150 (BASELINE configs[1])",
               coverage=100.0, cutoffs=(10.0, 10.0, 1e-6, 3), del_prop=30.0, cpu_div=64, cpu_baseline_div=16),
}
# kept for tests / scripts that import the workload definition (configs[1])
READ_SETS = [dict(name="REL606_pe150", paired=True, read_len=150, coverage=100.0, frag_mean=400, frag_sd=40)]
MUTATION_CUTOFF, POLYMORPHISM_CUTOFF, PRECISION, PLACES = CONFIGS["c1"]["cutoffs"]


def read_sets(cfg):
    return [dict(name="REL606_pe150", paired=True, read_len=150, coverage=float(cfg["coverage"]), frag_mean=400, frag_sd=40)]


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=3)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def oracle_cli():
    path = os.path.join(ROOT, "oracle", "_build", "oracle_cli")
    if not os.path.exists(path):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    return path


def ref_cli():
    """The reference's own sources compiled against the htslib shim (oracle/ref_build.sh), if prebuilt."""
    path = os.path.join(ROOT, "oracle", "_ref", "ref_cli")
    return path if os.path.exists(path) else None


def cpu_kind():
    return "reference" if ref_cli() else "port"


def write_sample_bam(tmp, cfg, div, seed=2):
    """The config's read model on a 1/div-length reference, written as BAM + FASTA by a CHILD process (the CPU arms must not
    load the product's library into the process that times them)."""
    bam, fasta = os.path.join(tmp, "s.bam"), os.path.join(tmp, "s.fasta")
    code = ("import sys; sys.path.insert(0, %r); import breseq_b200 as bq; ctx = bq.Context(device=-1); "
            "spec = bq.SynthSpec(seed=%d, read_sets=%r, contig_lens=[%d], contig_prefix='REL606s', n_polymorphic=4, n_fixed=2, n_gaps=1); "
            "ctx.synth_write(spec, %r, %r); ctx.close()" % (ROOT, seed, read_sets(cfg), GENOME // div, bam, fasta))
    subprocess.run([sys.executable, "-c", code], check=True)
    return bam, fasta


def run_cpu_once(bam, fasta, out, cfg, cli=None, count_records=True):
    """Both passes of the CPU implementation on one BAM; returns (records, seconds).

    `cli` = oracle/_ref/ref_cli (the reference's own code, kind "reference") when it was built, else
    oracle/_build/oracle_cli (the restatement, kind "port").  Both take the same command line and time
    the two entry points with steady_clock, the bracket the reference's own ExecutionTime uses."""
    cli = cli or ref_cli() or oracle_cli()
    os.makedirs(out, exist_ok=True)
    sets = "REL606_pe150:2"
    mc, pc, prec, places = cfg["cutoffs"]
    ec = ["error_count", "--bam", bam, "--fasta", fasta, "--out", out, "--covariates", COVARIATES, "--readfiles", "r1,r2",
          "--read-sets", sets]
    im = ["identify_mutations", "--bam", bam, "--fasta", fasta, "--out", out, "--error-rates", os.path.join(out, "error_rates.tab"),
          "--gd", os.path.join(out, "o.gd"), "--read-sets", sets, "--del-prop", str(cfg["del_prop"]), "--del-seed", "0", "--mutation-cutoff",
          str(mc), "--polymorphism-cutoff", str(pc), "--precision", str(prec), "--places", str(places)]
    a = json.loads(subprocess.run([cli] + ec, check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1])
    b = json.loads(subprocess.run([cli] + im, check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1])
    records = b["records"]
    if not records and count_records:  # the reference binary does not count records: one untimed run of the port does
        records = json.loads(subprocess.run([oracle_cli()] + im, check=True, capture_output=True, text=True)
                             .stdout.strip().splitlines()[-1])["records"]
    return records, a["seconds"] + b["seconds"]


def config_block(name, cfg, n_gpus):
    return {"workload": cfg["label"] + "; one 4 629 812 bp genome cut into %d reference-coordinate range%s of equal aligned bases"
                        % (n_gpus, "" if n_gpus == 1 else "s"),
            "config": name, "covariates": COVARIATES, "coverage": cfg["coverage"],
            "cutoffs": dict(zip(("mutation", "polymorphism", "precision", "places"), cfg["cutoffs"])),
            "l2_policy": "inputs (tens of GB of streams per run in HBM) are far larger than the 126 MB L2; no explicit flush",
            "parallelism": "reference-range sharding x%d (strong scaling), one sum of the integer histograms over the ranks (fused into "
                           "pass 1: in-kernel NVLink reductions, csrc/exchange.cu; BRQ_BENCH_NCCL=1: an NCCL allreduce), "
                           "evidence shares gathered to rank 0" % n_gpus}


def setup_fused_exchange(ctx, dist, rank, world):
    """Attach the ranks' contexts to each other for the fused histogram exchange (csrc/exchange.cu).  Every rank has to end up on
    the same side: if one of them cannot export or map a peer's inbox (no peer access between two of the GPUs, IPC not permitted
    in the container), all of them detach, the caller sums with an NCCL allreduce instead and the bench line says so.
    Returns whether the exchange is on."""
    why, own = None, None
    try:
        own = ctx.hist_exchange_export()
    except Exception as e:   # breseq_b200.BrqError
        why = str(e)
    handles = [None] * world
    dist.all_gather_object(handles, own)
    if why is None and all(h is not None for h in handles):
        try:
            ctx.hist_exchange_attach(handles, rank)
        except Exception as e:
            why = str(e)
    elif why is None:
        why = "a peer could not export its inbox"
    reasons = [None] * world
    dist.all_gather_object(reasons, why)
    failed = [r for r in reasons if r is not None]
    if failed:
        if why is None:
            ctx.hist_exchange_attach([own], 0)   # detach: one rank, no exchange
        if rank == 0:
            print("bench: fused histogram exchange unavailable (%s): NCCL allreduce instead" % failed[0], file=sys.stderr)
    dist.barrier()
    return not failed


def reference_arm(args, name, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(os.cpu_count() or 1, 32)
    div = cfg["cpu_div"]
    with tempfile.TemporaryDirectory() as tmp:
        bam, fasta = write_sample_bam(tmp, cfg, div)
        times, records = [], 0
        n_sample, _ = run_cpu_once(bam, fasta, os.path.join(tmp, "count"), cfg, cli=oracle_cli())  # untimed: records in the sample

        def one_step(tag):
            # `cores` independent processes, one coordinate range each (the reference itself is
            # single-threaded on this path; sharding by reference range is how it would be spread)
            t0 = time.perf_counter()
            results = [None] * cores

            def worker(i):
                results[i] = (n_sample, run_cpu_once(bam, fasta, os.path.join(tmp, "%s_%d" % (tag, i)), cfg, count_records=False)[1])
            th = [threading.Thread(target=worker, args=(i,)) for i in range(cores)]
            [t.start() for t in th]
            [t.join() for t in th]
            return sum(r[0] for r in results), time.perf_counter() - t0
        n_warm = min(args.warmup, 1)   # a warm-up step costs as much as a timed one here: one is enough to fault the binary in
        for w in range(n_warm):
            one_step("w%d" % w)
        for k in range(args.steps):
            n, dt = one_step("s%d" % k)
            records, _ = n, times.append(dt)
        dt = sum(times) / len(times)
        value = records / dt
    kind = cpu_kind()
    sample = ("%s read model on a 1/%d-length reference (%d bp; a per-record rate: the CPU path is linear in records), %d concurrent "
              "single-threaded processes, one reference range each" % (name, div, GENOME // div, cores))
    line = {"metric": "aligned bases/sec, error_count+identify_mutations", "value": value, "unit": "aligned bases/s",
            "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": n_warm, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(name, cfg, args.gpus),
            "cpu_baseline": {"value": value, "unit": "aligned bases/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "aligned bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def e2e_from_bam(bq, local, tmp):
    """BAM on disk -> error_rates.tab + ra_mc_evidence.gd on disk through the one-call adapters, BASELINE configs[1] on one GPU."""
    cfg = CONFIGS["c1"]
    ctx = bq.Context(device=-1)
    spec = bq.SynthSpec(seed=2, read_sets=read_sets(cfg), contig_lens=[GENOME], contig_prefix="REL606", n_polymorphic=40, n_fixed=10, n_gaps=3)
    bam, fasta = os.path.join(tmp, "c1.bam"), os.path.join(tmp, "c1.fasta")
    t0 = time.perf_counter()
    ctx.synth_write(spec, bam, fasta)
    t_write = time.perf_counter() - t0
    ctx.close()
    mc, pc, prec, places = cfg["cutoffs"]
    runs = []
    n_records = 0
    for k in range(5):
        out = os.path.join(tmp, "from_bam_%d" % k)
        os.makedirs(out)
        ctx = bq.Context(device=local)
        t0 = time.perf_counter()
        bq.error_count(bam, fasta, out, ["r1", "r2"], covariates=COVARIATES, read_file_sets=spec.read_file_sets(),
                       error_rates_file_name=os.path.join(out, "error_rates.tab"), ctx=ctx)
        t1 = time.perf_counter()
        bq.identify_mutations(bam, fasta, os.path.join(out, "ra_mc_evidence.gd"), [cfg["del_prop"]], [0.0], mc, pc, prec, places,
                              error_rates_file_name=os.path.join(out, "error_rates.tab"), read_file_sets=spec.read_file_sets(), ctx=ctx)
        t2 = time.perf_counter()
        n_records = int(ctx.stream_summary()["n_score"])
        ctx.close()
        runs.append({"error_count_s": t1 - t0, "identify_mutations_s": t2 - t1, "total_s": t2 - t0})
    best = min(runs, key=lambda r: r["total_s"])
    return {"value": n_records / best["total_s"], "unit": "aligned bases/s", "records": n_records, "seconds": best["total_s"],
            "error_count_seconds": best["error_count_s"], "identify_mutations_seconds": best["identify_mutations_s"],
            "runs_total_seconds": [r["total_s"] for r in runs], "bam_bytes": os.path.getsize(bam), "bam_write_seconds": t_write,
            "workload": CONFIGS["c1"]["label"] + ": BAM + FASTA on disk -> error_rates.tab, base_qual_error_prob.*.tab, coverage distribution and "
                        "ra_mc_evidence.gd on disk through brq_run_error_count + brq_run_identify_mutations (BGZF inflate and BAM decode on "
                        "the host, everything after it on one GPU; a fresh context per run, the best of five: the host side is at the "
                        "mercy of whatever else the box's cores are doing)",
            "n_gpus_used": 1}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the genome (debugging only; reported in config)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiler runs)")
    ap.add_argument("--no-bam", action="store_true", help="skip the e2e_from_bam leg (profiler runs)")
    ap.add_argument("--coverage", type=float, default=None, help="override the read depth (debugging only; reported in config)")
    ap.add_argument("--staging", default="auto", choices=["auto", "host", "device"])
    args = ap.parse_args()
    name = args.config
    cfg = dict(CONFIGS[name])
    if args.coverage:
        cfg["coverage"] = float(args.coverage)
        cfg["del_prop"] = 0.3 * float(args.coverage)
    if args.impl == "reference":
        reference_arm(args, name, cfg)
        return

    import numpy as np
    import torch
    import breseq_b200 as bq

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL's version banner, for one) write to stdout: keep fd 1 clean for the one JSON line
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    mc, pc, prec, places = cfg["cutoffs"]
    genome = int(GENOME * args.scale)
    rs = read_sets(cfg)
    ctx = bq.Context(device=local)
    full = bq.SynthSpec(seed=2, read_sets=rs, contig_lens=[genome], contig_prefix="REL606", n_polymorphic=40, n_fixed=10, n_gaps=3)
    bounds = ctx.synth_shard_bounds(full, world)   # equal aligned bases per range; every rank computes the same cuts
    lo, hi = bounds[rank], bounds[rank + 1]
    spec = bq.SynthSpec(seed=2, read_sets=rs, contig_lens=[genome], contig_prefix="REL606", n_polymorphic=40, n_fixed=10, n_gaps=3,
                        window=(lo, hi) if world > 1 else (0, 0))
    t0 = time.perf_counter()
    ctx.stage_synthetic(spec, read_file_sets=spec.read_file_sets(), shard_bounds=(lo, hi) if world > 1 else None, staging=args.staging)
    ctx.sync()
    t_stage = time.perf_counter() - t0
    s = ctx.stream_summary()
    n_records, n_slots, n_hist = int(s["n_score"]), int(s["n_base"] + s["n_ins"]), int(s["n_hist"])
    ctx.upload()
    ctx.sync()
    params = bq.Context.score_params(mc, pc, prec, places)

    class DevArray:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}

    ext_stream = torch.cuda.ExternalStream(ctx.cuda_stream(), device=dev) if world > 1 else None
    if world > 1:  # the ranks sum their coverage histograms too: one depth axis for all
        depth = torch.tensor([ctx.max_coverage_depth()], dtype=torch.int64, device=dev)
        dist.all_reduce(depth, op=dist.ReduceOp.MAX)
        ctx.set_min_coverage_depth(int(depth.item()))
    hist_view = {}
    # the collective of pass 1: fused into brq_error_count (csrc/exchange.cu: the ranks add their histograms into each other's
    # memory over NVLink, no library call, no host in the loop), or BRQ_BENCH_NCCL=1: an NCCL allreduce between the calls
    fused = 1 < world <= 16 and not os.environ.get("BRQ_BENCH_NCCL")
    if fused:
        fused = setup_fused_exchange(ctx, dist, rank, world)

    def allreduce_hist():
        if world == 1 or fused:
            return
        c, n, v, m = ctx.hist_device()
        # the collective is ordered on the context's own stream: no host synchronisation
        if hist_view.get("key") != (c, n, v, m):  # the buffer is allocated once: wrap it once
            assert v == c + 8 * n   # the coverage histogram follows the counts: ONE collective sums both
            hist_view["key"] = (c, n, v, m)
            hist_view["t"] = torch.as_tensor(DevArray(c, n + m), device=dev)
        with torch.cuda.stream(ext_stream):
            dist.all_reduce(hist_view["t"])

    call_s = {"error_count": 0.0, "allreduce": 0.0, "derive": 0.0, "score": 0.0}

    def step():
        # (host time inside each call; only the last one waits for the device)
        t0 = time.perf_counter(); ctx.error_count(COVARIATES)
        t1 = time.perf_counter(); allreduce_hist()
        t2 = time.perf_counter(); ctx.derive_error_table()
        t3 = time.perf_counter(); ctx.score_columns(params)
        t4 = time.perf_counter()
        for k, v in zip(call_s, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            call_s[k] += v

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    # the clock sampler starts before the warm-up: its first nvidia-smi query initialises the tool and holds a driver lock for
    # tens of milliseconds, which a kernel launch of the first timed step would otherwise wait for
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    launches0 = ctx.launch_count()
    for k in call_s:
        call_s[k] = 0.0
    barrier()
    ctx.event_record(0)
    t0 = time.perf_counter()
    k_ms = {"hist": 0.0, "coverage": 0.0, "derive": 0.0, "score": 0.0, "tally": 0.0, "fit": 0.0}
    step_walls = []
    for _ in range(args.steps):
        ts = time.perf_counter()
        step()
        step_walls.append((time.perf_counter() - ts) * 1e3)
        for k, v in ctx.kernel_ms().items():
            k_ms[k] += v
    if os.environ.get("BRQ_BENCH_STEP_TIMES"):  # debugging aid: every step ends in a synchronisation, so its wall time is its duration
        print("step wall ms: " + " ".join("%.3f" % w for w in step_walls), file=sys.stderr)
    ctx.event_record(1)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ctx.event_elapsed_ms(0, 1)
    clocks = sampler.summary()
    launches = ctx.launch_count() - launches0
    # every step ends in a synchronisation (the scoring call reads its error word), so a step's wall time is its duration, host
    # gaps included; the CUDA events bracket the same region on the launching stream
    step_ms = max(dev_ms, sum(step_walls)) / args.steps
    if os.environ.get("BRQ_BENCH_STEP_TIMES"):
        print("timed region: events %.3f ms, sum of step walls %.3f ms, wall incl. the closing barrier %.3f ms" % (dev_ms, sum(step_walls), wall * 1e3), file=sys.stderr)
    for k in k_ms:
        k_ms[k] /= args.steps

    # ---- end to end through the C ABI from HOST buffers: H2D of the reads + device staging + kernels + D2H + finalisation + files
    # (the decoded reads stay in pageable host memory, where a caller's BAM decoder leaves them: the library stages them through
    # its own page-locked ring, csrc/expand.cu UploadRing, inside the timed region.  BRQ_BENCH_PIN=1 registers them page-locked
    # once, outside the timed region: measured, it changes nothing, the copy already overlaps the expander's host-side planning)
    t_pin = 0.0
    if bool(s["device_built"]) and os.environ.get("BRQ_BENCH_PIN"):
        tp = time.perf_counter()
        ctx.pin_reads()
        t_pin = time.perf_counter() - tp
    device_built = bool(s["device_built"])
    with tempfile.TemporaryDirectory() as tmp:
        phase = {}

        def timed(name_, fn, *a):
            t = time.perf_counter()
            r = fn(*a)
            phase[name_] = phase.get(name_, 0.0) + time.perf_counter() - t
            return r

        def gather_evidence():
            blob = ctx.evidence_export([cfg["del_prop"]])
            if world == 1:
                return [blob]
            n = torch.tensor([len(blob)], dtype=torch.int64, device=dev)
            sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
            dist.all_gather(sizes, n)
            cap = max(int(x.item()) for x in sizes)
            buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
            buf[:len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
            parts = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
            dist.gather(buf, parts, dst=0)
            if rank != 0:
                return None
            return [bytes(p[:int(sz.item())].cpu().numpy().tobytes()) for p, sz in zip(parts, sizes)]

        def e2e_step():
            if device_built:
                timed("h2d_reads_and_expand", lambda: (ctx.restage(), ctx.sync()))
            else:
                timed("h2d_stream", lambda: (ctx.upload(), ctx.sync()))
            # (neither call waits for its kernels: the phase ends in a synchronisation so that it reads as kernel time)
            timed("pass1_kernels", lambda: (ctx.error_count(COVARIATES), allreduce_hist(), ctx.derive_error_table(), ctx.sync()))
            if rank == 0:
                timed("pass1_files", ctx.write_error_count_files, tmp, os.path.join(tmp, "error_rates.tab"), ["r1", "r2"])
            timed("pass2_kernels", ctx.score_columns, params)
            shares = timed("d2h_finalise_gather", gather_evidence)
            if rank == 0:
                timed("merged_gd", ctx.write_evidence_merged, os.path.join(tmp, "ra_mc_evidence.gd"), shares, [cfg["del_prop"]], [0.0])
        e2e_step()
        barrier()
        phase.clear()  # the first pass through the C ABI warms caches and allocations: not part of the averages
        ctx.d2h_bytes(reset=True)
        t0 = time.perf_counter()
        n_e2e = max(2, min(args.steps, 5))
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / n_e2e
        d2h = ctx.d2h_bytes() // n_e2e  # counted by the library: histograms, error table, walk events, flagged slots and their records
        e2e_phase_ms = {k: 1e3 * v / n_e2e for k, v in phase.items()}
        gd_rows = sum(1 for l in open(os.path.join(tmp, "ra_mc_evidence.gd")) if not l.startswith("#")) if rank == 0 else 0
    h2d = int(ctx.stream_summary()["bytes_host"])

    # ---- max over ranks, whole-job aggregate
    stats = torch.tensor([step_ms, e2e_s, float(n_records), k_ms["score"], k_ms["hist"], k_ms["tally"], float(h2d), float(d2h), t_stage],
                         dtype=torch.float64, device="cuda")
    per_rank = None
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        allr = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(allr, stats)
        per_rank = {"records": [int(x[2].item()) for x in allr], "tally_ms": [round(x[5].item(), 4) for x in allr],
                    "score_ms": [round(x[3].item(), 4) for x in allr], "hist_ms": [round(x[4].item(), 4) for x in allr],
                    "bounds": bounds}
        step_ms, e2e_s, total_records = mx[0].item(), mx[1].item(), sm[2].item()
        h2d, d2h = int(sm[6].item()), int(sm[7].item())
    else:
        total_records = float(n_records)

    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        score_bytes = 4 * n_records + 104 * n_slots           # 4 B/record + 8 B offsets + 96 B result per slot
        # pass 1 is charged the bytes its records really have (2 per fast record, 4 per exception; SURVEY.md 8d budgets 8)
        if s["hist_compact"]:
            hist_bytes = 2 * int(s["n_hist16"]) + 4 * int(s["n_hist_exc"]) + 8 * int(s["n_base"])
        else:
            hist_bytes = int(s["hist_record_bytes"]) * n_hist + 8 * int(s["n_base"])
        # the dominant kernel is the tally kernel: it moves all of the scoring pass's algorithmic bytes (the fit kernel
        # re-reads the work-list slots); its duration is measured with CUDA events on the launching stream (rank 0's here)
        achieved = score_bytes / (k_ms["tally"] * 1e-3) / 1e9 if k_ms["tally"] > 0 else 0.0
        pass_gbs = score_bytes / (k_ms["score"] * 1e-3) / 1e9 if k_ms["score"] > 0 else 0.0
        traffic = None
        try:  # DRAM bytes of one launch from the committed ncu capture of the same workload and rank count
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                t = json.load(f)
            if t.get("config") == name and int(t.get("n_gpus", 1)) == world and args.scale == 1.0 and args.coverage is None:
                traffic = int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
        except Exception:
            pass
        cb = config_block(name, cfg, world)
        cb["records"] = int(total_records)
        cb["records_rank0"] = n_records
        cb["slots_rank0"] = n_slots
        cb["genome_scale"] = args.scale
        if args.coverage is not None or args.scale != 1.0:  # another shape than the named config: say so in the label
            cb["workload"] = ("NOT a BASELINE config (--scale %.3f / --coverage %g override of %s): " % (args.scale, cfg["coverage"], name)) + cb["workload"]
        cb["kernel_ms_rank0"] = k_ms
        cb["host_call_ms_rank0"] = {k: 1e3 * v / args.steps for k, v in call_s.items()}
        cb["histogram_collective"] = "none (one rank)" if world == 1 else ("fused into pass 1 (csrc/exchange.cu; its wait for the peers is inside kernel_ms.coverage)" if fused else "NCCL allreduce")
        cb["step_wall_ms_rank0"] = [round(w, 3) for w in step_walls]
        cb["staging"] = "device (csrc/expand.cu)" if device_built else "host (csrc/staging.cpp)"
        cb["staging_seconds"] = t_stage
        cb["staging_note"] = "read synthesis + H2D + device staging of the rank's range, once, before the timed regions (max over ranks)"
        cb["pin_reads_seconds"] = t_pin
        cb["e2e_host_memory"] = ("page-locked (the decoded reads registered once, %.2f s, outside the timed region)" % t_pin) if t_pin > 0 else \
                                "pageable source, staged through the library's page-locked ring inside the timed region"
        cb["e2e_phase_ms_rank0"] = e2e_phase_ms
        cb["e2e_evidence_rows"] = gd_rows
        cb["hist_kernel_gbs"] = hist_bytes / (k_ms["hist"] * 1e-3) / 1e9 if k_ms["hist"] > 0 else 0.0
        if per_rank:
            cb["per_rank"] = per_rank
        cpu = None
        if not args.no_cpu:
            with tempfile.TemporaryDirectory() as tmp:
                div = cfg["cpu_baseline_div"]
                bam, fasta = write_sample_bam(tmp, cfg, div)
                n_cpu, t_cpu = run_cpu_once(bam, fasta, os.path.join(tmp, "o"), cfg)
            cpu = {"value": n_cpu / t_cpu, "unit": "aligned bases/s", "cores": 1, "kind": cpu_kind(),
                   "sample": "%s read model on a 1/%d-length reference (%d bp): %d records in %.1f s, one thread; a per-record rate "
                             "(the CPU path is linear in records), the full workload is %d x the sample"
                             % (name, div, GENOME // div, n_cpu, t_cpu, int(total_records) // max(1, n_cpu))}
        from_bam = None
        if not args.no_bam:
            with tempfile.TemporaryDirectory() as tmp:
                ctx.close()   # (its streams make room for the second workload)
                ctx = None
                from_bam = e2e_from_bam(bq, local, tmp)
        line = {"metric": "aligned bases/sec, error_count+identify_mutations", "value": total_records / (step_ms * 1e-3),
                "unit": "aligned bases/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": cb, "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": total_records / e2e_s, "unit": "aligned bases/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "seconds_per_step": e2e_s},
                "e2e_from_bam": from_bam,
                "roofline": {"bound": "hbm", "kernel": "tally_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_kind": peak_kind,
                             "algorithmic_bytes": score_bytes, "scoring_pass_frac": pass_gbs / peak},
                "cpu_baseline": cpu}
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if ctx is not None:
        ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
