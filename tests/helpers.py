"""Shared test plumbing: dataset definitions, the oracle runner, and numpy re-statements of what
the kernels count (used to check the HOST staging on machines without a GPU)."""
import json
import os
import subprocess

import numpy as np

import breseq_b200 as bq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_CLI = os.path.join(ROOT, "oracle", "_build", "oracle_cli")

# written by oracle --columns-out (oracle.h: ColumnDump)
ORACLE_COLUMN = np.dtype([("tid", "<u4"), ("pos1", "<u4"), ("insert_count", "<u4"), ("n", "<u4"), ("ll", "<f8", 5),
                          ("consensus_score", "<f8"), ("variant_score", "<f8"), ("f", "<f8", 5), ("log10_likelihood", "<f8"),
                          ("unique", "<f8", 2), ("redundant", "<f8", 2), ("raw_redundant", "<i4", 2), ("total", "<i4"),
                          ("best", "u1"), ("major", "u1"), ("minor", "u1"), ("variant", "u1"), ("ref", "u1"),
                          ("base_predicted", "u1"), ("unique_only", "u1"), ("emitted", "u1"), ("iterations", "<u4")])
assert ORACLE_COLUMN.itemsize == 176

# name -> generator + run settings.  Sizes are chosen so the oracle finishes in seconds.
DATASETS = {
    # stand-in for tests/lambda_polymorphism: 48.5 kb, ~107x se35, polymorphism mode
    "lambda": dict(seed=1, contig_lens=[48502], prefix="NC_001416",
                   read_sets=[dict(name="lambda_reads", paired=False, read_len=35, coverage=107.0)],
                   n_polymorphic=60, n_fixed=10, n_gaps=2, mutation_cutoff=10.0, polymorphism_cutoff=2.0,
                   precision=1e-6, places=8, del_prop=20.0, del_seed=0.0),
    # three contigs, one paired + one single-end read group => read_set=3, consensus-mode cutoffs
    "multi": dict(seed=7, contig_lens=[6000, 3500, 5000], prefix="ctg",
                  read_sets=[dict(name="pe", paired=True, read_len=150, coverage=60.0, frag_mean=400, frag_sd=40),
                             dict(name="se", paired=False, read_len=36, coverage=30.0)],
                  n_polymorphic=25, n_fixed=8, n_gaps=2, mutation_cutoff=10.0, polymorphism_cutoff=10.0,
                  precision=1e-6, places=3, del_prop=15.0, del_seed=0.0),
    # small enough that its BAM and the reference's outputs for it are committed under tests/golden/tiny
    "tiny": dict(seed=11, contig_lens=[1200, 800], prefix="tiny",
                 read_sets=[dict(name="tp", paired=True, read_len=100, coverage=30.0, frag_mean=250, frag_sd=25),
                            dict(name="ts", paired=False, read_len=36, coverage=15.0)],
                 n_polymorphic=8, n_fixed=4, n_gaps=1, mutation_cutoff=10.0, polymorphism_cutoff=2.0,
                 precision=1e-6, places=8, del_prop=6.0, del_seed=0.0),
    # deep columns (mean depth above 512 switches the tally kernel to a whole warp per slot), polymorphism mode
    "deep": dict(seed=5, contig_lens=[700], prefix="amplicon",
                 read_sets=[dict(name="amp", paired=False, read_len=100, coverage=1400.0)],
                 n_polymorphic=10, n_fixed=3, n_gaps=1, mutation_cutoff=10.0, polymorphism_cutoff=2.0,
                 precision=1e-6, places=8, del_prop=100.0, del_seed=0.0),
    # stand-in for BASELINE configs[3]: se36 + pe50 read groups (3 read files), polymorphism mode, and the
    # read_pos covariate in BOTH passes: every scoring record is a class of its own table row (no shared table)
    "ltee": dict(seed=13, contig_lens=[3000], prefix="ltee",
                 read_sets=[dict(name="s36", paired=False, read_len=36, coverage=35.0),
                            dict(name="p50", paired=True, read_len=50, coverage=50.0, frag_mean=160, frag_sd=15)],
                 n_polymorphic=12, n_fixed=4, n_gaps=1, mutation_cutoff=10.0, polymorphism_cutoff=2.0,
                 precision=1e-6, places=8, del_prop=8.0, del_seed=0.0,
                 covariates="read_set=3,obs_base,ref_base,quality=42,read_pos=50,base_repeat=4", big_table=True),
    # the north-star shape in small: polymorphism mode at 1000x, where the work list is a sixth of the columns and the screen
    # kernel (one pass of likelihood bounds) settles nearly all of it
    "pop1000": dict(seed=23, contig_lens=[12000], prefix="pop",
                    read_sets=[dict(name="pp", paired=True, read_len=150, coverage=1000.0, frag_mean=400, frag_sd=40)],
                    n_polymorphic=12, n_fixed=3, n_gaps=1, mutation_cutoff=10.0, polymorphism_cutoff=2.0,
                    precision=1e-6, places=8, del_prop=300.0, del_seed=0.0, no_golden=True),
}

REF_CLI = os.path.join(ROOT, "oracle", "_ref", "ref_cli")  # the reference's own sources (oracle/ref_build.sh)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def synth_spec(d):
    return bq.SynthSpec(seed=d["seed"], read_sets=d["read_sets"], contig_lens=d["contig_lens"], contig_prefix=d["prefix"],
                        n_polymorphic=d["n_polymorphic"], n_fixed=d["n_fixed"], n_gaps=d["n_gaps"])


def read_file_sets(d):
    return [(rs["name"], 2 if rs.get("paired") else 1) for rs in d["read_sets"]]


def readfile_names(d):
    out = []
    for rs in d["read_sets"]:
        out += [rs["name"] + "_R1", rs["name"] + "_R2"] if rs.get("paired") else [rs["name"]]
    return out


def covariates(d, n_qual=42):
    if d.get("covariates"):
        return d["covariates"]
    return "read_set=%d,obs_base,ref_base,quality=%d" % (len(readfile_names(d)), n_qual)


def stage_kwargs(d):
    """Staging options a dataset's covariate string asks for (brq_stage_options.use_read_pos / use_base_repeat)."""
    c = covariates(d)
    return dict(read_file_sets=read_file_sets(d), use_read_pos="read_pos" in c, use_base_repeat="base_repeat" in c)


def run_oracle(*args):
    p = subprocess.run([ORACLE_CLI] + [str(a) for a in args], check=True, capture_output=True, text=True)
    return json.loads(p.stdout.strip().splitlines()[-1])


def cli_args(d, outdir, rates=None, gd=None):
    """Command lines shared by oracle_cli and ref_cli (same options by construction)."""
    sets = ",".join("%s:%d" % s for s in read_file_sets(d))
    n = len(d["contig_lens"])
    rates = rates or os.path.join(outdir, "error_rates.tab")
    gd = gd or os.path.join(outdir, "ra_mc_evidence.gd")
    ec = ["error_count", "--bam", d["bam"], "--fasta", d["fasta"], "--out", outdir, "--covariates", covariates(d),
          "--readfiles", ",".join(readfile_names(d)), "--read-sets", sets]
    im = ["identify_mutations", "--bam", d["bam"], "--fasta", d["fasta"], "--out", outdir, "--error-rates", rates, "--gd", gd,
          "--read-sets", sets, "--del-prop", ",".join([str(d["del_prop"])] * n), "--del-seed", ",".join([str(d["del_seed"])] * n),
          "--mutation-cutoff", d["mutation_cutoff"], "--polymorphism-cutoff", d["polymorphism_cutoff"],
          "--precision", d["precision"], "--places", d["places"]]
    return ec, im


def generate_inputs(name, outdir):
    """BAM + FASTA of a named dataset, written by the product's (seeded, integer-only) generator."""
    d = dict(DATASETS[name])
    os.makedirs(outdir, exist_ok=True)
    d["name"], d["dir"] = name, outdir
    d["bam"], d["fasta"] = os.path.join(outdir, "reference.bam"), os.path.join(outdir, "reference.fasta")
    ctx = bq.Context(device=-1)
    ctx.synth_write(synth_spec(d), d["bam"], d["fasta"])
    ctx.close()
    return d


def make_dataset(name, outdir):
    """Generate BAM + FASTA with the product's generator, then run both oracle passes on it."""
    d = generate_inputs(name, outdir)
    odir = os.path.join(outdir, "oracle")
    os.makedirs(odir, exist_ok=True)
    d["oracle_dir"] = odir
    d["oracle_counts"] = os.path.join(odir, "error_counts.tab")
    d["oracle_rates"] = os.path.join(odir, "error_rates.tab")
    d["oracle_gd"] = os.path.join(odir, "ra_mc_evidence.gd")
    d["oracle_columns"] = os.path.join(odir, "columns.bin")
    ec, im = cli_args(d, odir)
    run_oracle(*ec, "--counts-dump", d["oracle_counts"])
    run_oracle(*im, "--columns-out", d["oracle_columns"])
    return d


def run_reference(d, outdir, per_position=True, coverage_tsv=False):
    """Both passes of the reference's own sources (oracle/_ref/ref_cli) on a dataset; returns seconds."""
    os.makedirs(outdir, exist_ok=True)
    ec, im = cli_args(d, outdir)
    if per_position:
        im += ["--per-position", os.path.join(outdir, "per_position_file.tab")]
    if coverage_tsv:
        im += ["--coverage-tsv", os.path.join(outdir, "@.coverage.tsv")]
    sec = 0.0
    for args in (ec, im):
        p = subprocess.run([REF_CLI] + [str(a) for a in args], check=True, capture_output=True, text=True)
        sec += json.loads(p.stdout.strip().splitlines()[-1])["seconds"]
    return sec


PREPROCESS_TAB = "preprocess_error_count.tab"  # seq id <tab> no_pos_hash_per_position_pr (%.17g), what the stage 03 call leaves in the Summary


def run_preprocess(cli, d, outdir):
    """error_count(..., preprocess_stage = true) (breseq_cmdline.cpp:1969: coverage only) through oracle_cli or ref_cli;
    returns the text of <outdir>/preprocess_error_count.tab."""
    os.makedirs(outdir, exist_ok=True)
    ec, _ = cli_args(d, outdir)
    subprocess.run([cli] + [str(a) for a in ec] + ["--preprocess", "--no-errors"], check=True, capture_output=True, text=True)
    return open(os.path.join(outdir, PREPROCESS_TAB)).read()


def product_preprocess_tab(d, shards=1):
    """The same table from the product's staging layer (host only: no device is needed for it); shards add up."""
    names = contig_names(d)
    total = np.zeros((len(names), 2), np.float64)
    for rank in range(shards):
        ctx = bq.Context(device=-1)
        ctx.stage_bam(d["bam"], d["fasta"], preprocess_stage=True, shard=(rank, shards), **stage_kwargs(d))
        total += ctx.preprocess_read_starts()
        ctx.close()
    rows = sorted((names[t], total[t, 0] / total[t].sum() if total[t].sum() else 1.0) for t in range(len(names)))
    return "".join("%s\t%.17g\n" % r for r in rows)


def pass_output_names(d):
    """Files the two entry points write for a dataset (the drop-in boundary's file contract)."""
    # a table with read_pos / base_repeat has 10^5 .. 10^6 rows: its golden is a checksum (test_golden.py), not a copy
    names = ["ra_mc_evidence.gd"] if d.get("big_table") else ["error_rates.tab", "ra_mc_evidence.gd"]
    names += ["base_qual_error_prob.%s.tab" % rf for rf in readfile_names(d)]
    names += ["%d.unique_only_coverage_distribution.tab" % g for g in range(len(d["contig_lens"]))]
    return names


def oracle_counts(path):
    lines = open(path).read().strip().split("\n")[2:]
    return np.array([int(float(l.split("\t")[-1])) for l in lines], dtype=np.int64)


def oracle_columns(path):
    return np.fromfile(path, dtype=ORACLE_COLUMN)


def oracle_slots(o, stream, target_slot0):
    """Slot index (product numbering) of every oracle column row."""
    nb = int(stream["n_base"])
    slot = np.where(o["insert_count"] == 0, target_slot0[o["tid"]] + o["pos1"].astype(np.int64) - 1, -1)
    key = {(int(p), int(k)): nb + i for i, (p, k) in enumerate(zip(stream["ins_parent"], stream["ins_count"]))}
    for i in np.nonzero(o["insert_count"] > 0)[0]:
        slot[i] = key[(int(target_slot0[o["tid"][i]]) + int(o["pos1"][i]) - 1, int(o["insert_count"][i]))]
    return slot


def visit_slot0(contig_names, contig_lens):
    """First base slot of each BAM tid when targets are visited in alphabetical order."""
    order = sorted(range(len(contig_names)), key=lambda i: contig_names[i])
    slot0 = np.zeros(len(contig_names), dtype=np.int64)
    acc = 0
    for i in order:
        slot0[i] = acc
        acc += contig_lens[i]
    return slot0


def contig_names(d):
    n = len(d["contig_lens"])
    if n == 1:
        return [d["prefix"]]
    width = len(str(n))
    return ["%s%0*d" % (d["prefix"], width, i + 1) for i in range(n)]


# ---- numpy statements of what the kernels count (test-side only) --------------------------------
def emulate_hist(hist_rec, n_sets, n_qual):
    """Covariate histogram for the default layout read_set, ref_base, obs_base, quality (record layout: brq_types.h)."""
    r = hist_rec.astype(np.uint64)   # 4-byte records, or 8-byte ones whose low word has the same layout

    def f(sh, m):
        return ((r >> np.uint64(sh)) & np.uint64(m)).astype(np.int64)
    refA, obsA, qa, validA = f(0, 7), f(3, 7), f(6, 127), f(13, 1)
    refB, obsB, qb, validB, rset = f(14, 7), f(17, 7), f(20, 127), f(27, 1), f(28, 7) + 8 * f(61, 7)
    fast = f(31, 1)   # the dominant kind: A valid with ref == obs, B ('.', '.') or absent (its quality field then reads 127)
    assert np.array_equal(fast == 1, (validA == 1) & (refA == obsA) & (((validB == 1) & (refB == 4) & (obsB == 4)) | (validB == 0)))
    assert np.all(qb[(fast == 1) & (validB == 0)] == 127)
    N = n_sets
    counts = np.zeros(N * 25 * n_qual, np.int64)
    np.add.at(counts, (rset + refA * N + obsA * 5 * N + qa * 25 * N)[validA == 1], 1)
    np.add.at(counts, (rset + refB * N + obsB * 5 * N + qb * 25 * N)[validB == 1], 1)
    return counts


def expand_hist16(h16):
    """4-byte records of the 16-bit fast records (csrc/brq_types.h: hist16_expand)."""
    r = h16.astype(np.uint32)
    base, qa, qb, rset = r & 3, (r >> 2) & 63, (r >> 8) & 63, r >> 14
    lo = base | base << 3 | qa << 6 | np.uint32(1 << 13) | rset << 28 | np.uint32(1 << 31)
    b = np.where(qb == 63, np.uint32(127 << 20), np.uint32(4 << 14 | 4 << 17 | 1 << 27) | qb << 20)
    return (lo | b).astype(np.uint32)


def expand_score16(s):
    """score_rec rebuilt from its transfer form (csrc/brq_types.h: score_word_from16, expand_score_kernel), in numpy."""
    g = s["geometry"]
    n_sq = g["n_st"] * g["n_q"]

    def counter(which):
        i = n_sq + which
        return (i >> 2) * 128 + (i & 3) * 8
    lo16 = s["score16"].astype(np.uint32)
    p = np.arange(len(lo16), dtype=np.int64)
    roff = s["round_off"].astype(np.int64)
    r = np.searchsorted(roff, p, side="right") - 1
    lane = ((p - roff[r]) >> 2) & 31
    key = r * 32 + lane
    sl = s["round_slot"][key]
    ref = np.where(sl == 0xFFFFFFFF, 5, s["slot_ref"][np.minimum(sl, len(s["slot_ref"]) - 1)]).astype(np.uint32)
    lo = lo16 & 0x7FFF
    c, top = lo & 0x1FFF, (lo >> 13) & 1
    sq = ((c >> 7) << 2) | ((c >> 3) & 3)
    out = (sq << 16) | (ref << 24) | np.uint32(1 << 28) | lo                                   # HOT, matching the slot's base
    out = np.where(c == np.where(top == 1, counter(2), counter(3)), np.uint32(2 << 30) | lo, out)  # COLD
    out = np.where(c == np.where(top == 1, counter(0), counter(1)), np.uint32(1 << 30) | lo, out)  # IDLE
    out = np.where(c == counter(6), lo, out)                                                       # pad
    flagged = np.flatnonzero(lo16 & 0x8000)
    if len(flagged):
        order = np.argsort(key[flagged], kind="stable")          # memory order within a lane is record order
        k = key[flagged][order]
        first = np.flatnonzero(np.r_[True, k[1:] != k[:-1]])
        rank = np.arange(len(k)) - np.repeat(first, np.diff(np.r_[first, len(k)]))
        out[flagged[order]] = s["score_exc"][s["score_exc_off"].astype(np.int64)[k] + rank]
    return out.astype(np.uint32)


def emulate_coverage_hist(hist_off):
    red = (hist_off[:-1] >> np.uint64(63)).astype(bool)
    off = (hist_off & np.uint64((1 << 63) - 1)).astype(np.int64)
    depth = np.diff(off)
    return np.bincount(depth[~red])


def emulate_tally(stream, base_quality_cutoff=3):
    """What the tally kernel counts per slot, from the decoded stream (numpy, test-side only)."""
    assert stream["geometry"]["base_quality_cutoff"] == base_quality_cutoff
    r = bq.decode_score_records(stream)
    n_slots = len(stream["score_off"]) - 1
    assert len(r["slot"]) == int(stream["n_score"])
    sid, uniq, top = r["slot"], r["unique"], r["top"]
    out = {}
    for name, u in (("unique", True), ("raw_redundant", False)):
        out[name] = np.stack([np.bincount(sid[(uniq == u) & (top == 0)], minlength=n_slots),
                              np.bincount(sid[(uniq == u) & (top == 1)], minlength=n_slots)], axis=1)
    out["n"] = np.bincount(sid[r["scores"]], minlength=n_slots)
    red = np.zeros((n_slots, 2))
    for i in np.nonzero(~uniq)[0]:  # order-dependent double sum, arrival order
        red[sid[i], top[i]] += 1.0 / float(r["x1"][i])
    out["redundant"] = red
    return out


def parse_gd(path):
    rows = []
    for line in open(path):
        if line.startswith("#") or not line.strip():
            continue
        f = line.rstrip("\n").split("\t")
        n_spec = {"RA": 5, "MC": 5, "UN": 3}[f[0]]
        rows.append(dict(type=f[0], id=f[1], spec=f[3:3 + n_spec], kv=dict(x.split("=", 1) for x in f[3 + n_spec:])))
    return rows
