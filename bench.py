#!/usr/bin/env python
"""bench.py -- aligned bases/sec of error_count + identify_mutations on B200.

  python bench.py --gpus 1 --steps 10 --warmup 3          (one JSON line on stdout)
  torchrun ... bench.py --gpus N ...                        (one rank per GPU; weak scaling)
  python bench.py --impl reference ...                      (the CPU implementation, same metric)

Workload (BASELINE.json configs[1]): E. coli REL606-sized reference
(4 629 812 bp, one contig), synthetic 100x pe150 reads, one paired read set (read_set=2, Q=42).
At N GPUs every rank holds one such coordinate range (its own contig of an N x 4.6 Mb genome):
weak scaling, no data-path collective except the sum-allreduce of the integer histograms.

A "step" is one pass of both kernels' path over the resident stream:
  covariate histogram + coverage histogram -> [allreduce] -> table derivation + text
  canonicalisation + class table -> per-slot scoring.
`value` times that with the stream already in HBM; `e2e` adds the host->device copy of the
pinned stream, the device->host copy of the per-slot results and the host finalisation that
writes error_rates.tab and ra_mc_evidence.gd, i.e. what the reference-facing call does after
BAM staging.
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
READ_SETS = [dict(name="REL606_pe150", paired=True, read_len=150, coverage=100.0, frag_mean=400, frag_sd=40)]
COVARIATES = "read_set=2,obs_base,ref_base,quality=42"
MUTATION_CUTOFF, POLYMORPHISM_CUTOFF, PRECISION, PLACES = 10.0, 10.0, 1e-6, 3  # clone / consensus mode (settings.cpp:914-960)
CPU_SAMPLE_DIV = 64  # the reference arm runs the same model on a 1/64-length reference, once per host core and step
CPU_BASELINE_DIV = 16  # the cpu_baseline leg of the main arm: one thread, about 10 s of CPU work


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


def cpu_sample(tmp, seed=2, div=CPU_SAMPLE_DIV):
    """The C1 model on a 1/div-length reference, written as BAM + FASTA for the CPU arms."""
    import breseq_b200 as bq
    ctx = bq.Context(device=-1)
    spec = bq.SynthSpec(seed=seed, read_sets=READ_SETS, contig_lens=[GENOME // div], contig_prefix="REL606s",
                        n_polymorphic=4, n_fixed=2, n_gaps=1)
    bam, fasta = os.path.join(tmp, "s.bam"), os.path.join(tmp, "s.fasta")
    ctx.synth_write(spec, bam, fasta)
    ctx.close()
    return bam, fasta


def ref_cli():
    """The reference's own sources compiled against the htslib shim (oracle/ref_build.sh), if prebuilt."""
    path = os.path.join(ROOT, "oracle", "_ref", "ref_cli")
    return path if os.path.exists(path) else None


def cpu_kind():
    return "reference" if ref_cli() else "port"


def run_cpu_once(bam, fasta, out, cli=None, count_records=True):
    """Both passes of the CPU implementation on one BAM; returns (records, seconds).

    `cli` = oracle/_ref/ref_cli (the reference's own code, kind "reference") when it was built, else
    oracle/_build/oracle_cli (the restatement, kind "port").  Both take the same command line and time
    the two entry points with steady_clock, the bracket the reference's own ExecutionTime uses."""
    cli = cli or ref_cli() or oracle_cli()
    os.makedirs(out, exist_ok=True)
    sets = READ_SETS[0]["name"] + ":2"
    ec = ["error_count", "--bam", bam, "--fasta", fasta, "--out", out, "--covariates", COVARIATES, "--readfiles", "r1,r2",
          "--read-sets", sets]
    im = ["identify_mutations", "--bam", bam, "--fasta", fasta, "--out", out, "--error-rates", os.path.join(out, "error_rates.tab"),
          "--gd", os.path.join(out, "o.gd"), "--read-sets", sets, "--del-prop", "30", "--del-seed", "0", "--mutation-cutoff",
          str(MUTATION_CUTOFF), "--polymorphism-cutoff", str(POLYMORPHISM_CUTOFF), "--places", str(PLACES)]
    a = json.loads(subprocess.run([cli] + ec, check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1])
    b = json.loads(subprocess.run([cli] + im, check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1])
    records = b["records"]
    if not records and count_records:  # the reference binary does not count records: one untimed run of the port does
        records = json.loads(subprocess.run([oracle_cli()] + im, check=True, capture_output=True, text=True)
                             .stdout.strip().splitlines()[-1])["records"]
    return records, a["seconds"] + b["seconds"]


def config_block(n_gpus):
    return {"workload": "E. coli REL606 4.6 Mb clone mode, synthetic 100x pe150 reads (BASELINE configs[1]); one 4 629 812 bp "
                        "coordinate range per GPU", "covariates": COVARIATES, "records_per_gpu": None,
            "l2_policy": "inputs (about 3 GB per GPU in HBM) are far larger than the 126 MB L2; no explicit flush",
            "parallelism": "reference-range sharding x%d, one sum-allreduce of the integer histograms" % n_gpus}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(os.cpu_count() or 1, 32)
    with tempfile.TemporaryDirectory() as tmp:
        bam, fasta = cpu_sample(tmp)
        times, records = [], 0
        n_sample, _ = run_cpu_once(bam, fasta, os.path.join(tmp, "count"), cli=oracle_cli())  # untimed: records in the sample

        def one_step(tag):
            # `cores` independent processes, one coordinate range each (the reference itself is
            # single-threaded on this path; sharding by reference range is how it would be spread)
            t0 = time.perf_counter()
            procs, results = [], [None] * cores

            def worker(i):
                results[i] = (n_sample, run_cpu_once(bam, fasta, os.path.join(tmp, "%s_%d" % (tag, i)), count_records=False)[1])
            th = [threading.Thread(target=worker, args=(i,)) for i in range(cores)]
            [t.start() for t in th]
            [t.join() for t in th]
            return sum(r[0] for r in results), time.perf_counter() - t0
        for w in range(args.warmup if args.warmup < 2 else 1):
            one_step("w%d" % w)
        for k in range(args.steps):
            n, dt = one_step("s%d" % k)
            records, _ = n, times.append(dt)
        dt = sum(times) / len(times)
        value = records / dt
    kind = cpu_kind()
    line = {"metric": "aligned bases/sec, error_count+identify_mutations", "value": value, "unit": "aligned bases/s",
            "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(args.gpus),
            "cpu_baseline": {"value": value, "unit": "aligned bases/s", "cores": cores, "kind": kind,
                             "sample": "C1 read model on a 1/%d-length reference (%d bp), %d concurrent single-threaded "
                                       "processes, one reference range each" % (CPU_SAMPLE_DIV, GENOME // CPU_SAMPLE_DIV, cores)},
            "e2e": {"value": value, "unit": "aligned bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the genome (debugging only; reported in config)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiler runs)")
    ap.add_argument("--coverage", type=float, default=None, help="override the read depth (other BASELINE shapes; reported in config)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
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
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ctx = bq.Context(device=local)
    if args.coverage:
        READ_SETS[0]["coverage"] = float(args.coverage)
    genome = int(GENOME * args.scale)
    spec = bq.SynthSpec(seed=2 + rank, read_sets=READ_SETS, contig_lens=[genome], contig_prefix="REL606_range%d" % rank,
                        n_polymorphic=40, n_fixed=10, n_gaps=3)
    t0 = time.perf_counter()
    ctx.stage_synthetic(spec, read_file_sets=spec.read_file_sets())
    t_stage = time.perf_counter() - t0
    s = ctx.stream()
    n_records, n_slots, n_hist = int(s["n_score"]), int(s["n_base"] + s["n_ins"]), int(s["n_hist"])
    ctx.upload()
    ctx.sync()
    params = bq.Context.score_params(MUTATION_CUTOFF, POLYMORPHISM_CUTOFF, PRECISION, PLACES)

    class DevArray:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}

    ext_stream = torch.cuda.ExternalStream(ctx.cuda_stream(), device=torch.device("cuda", local)) if world > 1 else None

    hist_view = {}

    def allreduce_hist():
        if world == 1:
            return
        c, n, v, m = ctx.hist_device()
        # every rank sizes its coverage histogram by its own deepest column: reduce the counts only, which is all the
        # table derivation needs.  The collective is ordered on the context's own stream: no host synchronisation.
        if hist_view.get("key") != (c, n):  # the buffer is allocated once: wrap it once
            hist_view["key"], hist_view["t"] = (c, n), torch.as_tensor(DevArray(c, n), device="cuda:%d" % local)
        with torch.cuda.stream(ext_stream):
            dist.all_reduce(hist_view["t"])

    def step():
        ctx.error_count(COVARIATES)
        allreduce_hist()
        ctx.derive_error_table()
        ctx.score_columns(params)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    for _ in range(max(args.warmup, 3)):
        step()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
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
    step_ms = max(dev_ms, wall * 1e3) / args.steps  # host gaps count: the step is not done before its table is built
    for k in k_ms:
        k_ms[k] /= args.steps

    # ---- end to end through the C ABI with host buffers: H2D + kernels + D2H + host finalisation
    with tempfile.TemporaryDirectory() as tmp:
        phase = {}

        def timed(name, fn, *a):
            t = time.perf_counter()
            fn(*a)
            phase[name] = phase.get(name, 0.0) + time.perf_counter() - t

        def e2e_step():
            timed("h2d", lambda: (ctx.upload(), ctx.sync()))
            # (neither call waits for its kernels any more: the phase ends in a synchronisation so that it reads as kernel time)
            timed("pass1_kernels", lambda: (ctx.error_count(COVARIATES), allreduce_hist(), ctx.derive_error_table(), ctx.sync()))
            timed("pass1_files", ctx.write_error_count_files, tmp, os.path.join(tmp, "error_rates.tab"), ["r1", "r2"])
            timed("pass2_kernels", ctx.score_columns, params)
            timed("d2h_finalise_gd", ctx.write_evidence, os.path.join(tmp, "ra_mc_evidence.gd"), [30.0], [0.0])
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
        d2h = ctx.d2h_bytes() // n_e2e  # counted by the library: histograms, error table, walk events, flagged slots
        e2e_phase_ms = {k: 1e3 * v / n_e2e for k, v in phase.items()}
    h2d = int(s["bytes_host"])

    # ---- max over ranks, whole-job aggregate
    stats = torch.tensor([step_ms, e2e_s, float(n_records), k_ms["score"], k_ms["hist"]], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        step_ms, e2e_s, total_records = mx[0].item(), mx[1].item(), sm[2].item()
    else:
        total_records = float(n_records)

    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        score_bytes = 4 * n_records + 104 * n_slots           # 4 B/record + 8 B offsets + 96 B result per slot
        # pass 1 is charged the bytes its records really have (4 per record at the default covariates, not SURVEY.md 8d's 8)
        if s.get("hist16") is not None:  # the compact form: 2 bytes per fast record, 4 per exception
            hist_bytes = 2 * len(s["hist16"]) + 4 * len(s["hist_exc"]) + 8 * int(s["n_base"])
        else:
            hist_bytes = int(s["hist_rec"].itemsize) * n_hist + 8 * int(s["n_base"])
        # the dominant kernel is the tally kernel: it moves all of the scoring pass's algorithmic bytes (the fit kernel
        # re-reads a few hundred slots); its duration is measured with CUDA events on the launching stream
        achieved = score_bytes / (k_ms["tally"] * 1e-3) / 1e9 if k_ms["tally"] > 0 else 0.0
        pass_gbs = score_bytes / (k_ms["score"] * 1e-3) / 1e9 if k_ms["score"] > 0 else 0.0
        traffic = None
        try:  # DRAM bytes of one launch from the committed ncu capture of the same workload
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                t = json.load(f)
            traffic = int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
        except Exception:
            pass
        cfg = config_block(world)
        cfg["records_per_gpu"] = n_records
        cfg["slots_per_gpu"] = n_slots
        cfg["genome_scale"] = args.scale
        cfg["coverage"] = READ_SETS[0]["coverage"]
        if args.coverage is not None or args.scale != 1.0:  # another shape than configs[1]: say so in the label
            cfg["workload"] = ("E. coli REL606-like %.2f x 4.6 Mb, synthetic %gx pe150 reads (not BASELINE configs[1]: --scale / --coverage "
                               "override; 1000x at full scale is configs[2]'s shape); one coordinate range per GPU" % (args.scale, READ_SETS[0]["coverage"]))
            cfg["l2_policy"] = "inputs are far larger than the 126 MB L2; no explicit flush"
            traffic = None  # the committed ncu capture is of configs[1]
        cfg["kernel_ms"] = k_ms
        cfg["staging_seconds"] = t_stage
        cfg["e2e_phase_ms"] = e2e_phase_ms
        cfg["hist_kernel_gbs"] = hist_bytes / (k_ms["hist"] * 1e-3) / 1e9 if k_ms["hist"] > 0 else 0.0
        cpu = None
        if not args.no_cpu:
            with tempfile.TemporaryDirectory() as tmp:
                bam, fasta = cpu_sample(tmp, div=CPU_BASELINE_DIV)
                n_cpu, t_cpu = run_cpu_once(bam, fasta, os.path.join(tmp, "o"))
            cpu = {"value": n_cpu / t_cpu, "unit": "aligned bases/s", "cores": 1, "kind": cpu_kind(),
                   "sample": "C1 read model on a 1/%d-length reference (%d bp): %d records in %.1f s, one thread"
                             % (CPU_BASELINE_DIV, GENOME // CPU_BASELINE_DIV, n_cpu, t_cpu)}
        line = {"metric": "aligned bases/sec, error_count+identify_mutations", "value": total_records / (step_ms * 1e-3),
                "unit": "aligned bases/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": cfg, "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": total_records / e2e_s, "unit": "aligned bases/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h},
                "roofline": {"bound": "hbm", "kernel": "tally_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_kind": peak_kind,
                             "algorithmic_bytes": score_bytes, "scoring_pass_frac": pass_gbs / peak},
                "cpu_baseline": cpu}
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
