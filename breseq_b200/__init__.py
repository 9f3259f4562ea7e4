"""breseq_b200 -- B200-native read-alignment evidence pileup behind breseq's entry points.

Python here is a thin ctypes binding over the C ABI in ``include/brq.h`` (``libbrq.so``: C++ host
staging + hand-written sm_100a kernels).  The two module-level functions mirror the reference's
free functions for this path:

* :func:`error_count`        <- ``breseq::error_count()``        (/root/reference/src/breseq/error_count.h:41-52)
* :func:`identify_mutations` <- ``breseq::identify_mutations()`` (/root/reference/src/breseq/identify_mutations.h:46-60)

There is no CPU fallback: if ``libbrq.so`` is missing, or no CUDA device is usable, the compute
calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BRQ_LIB_PATH") or os.path.join(_HERE, "libbrq.so")


class BrqError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("threads", C.c_int32)]


class _ReadFileSet(C.Structure):
    _fields_ = [("base_name", C.c_char_p), ("n_files", C.c_uint32)]


class _StageOptions(C.Structure):
    _fields_ = [("seq_ids", C.POINTER(C.c_char_p)), ("n_seq_ids", C.c_uint32),
                ("read_file_sets", C.POINTER(_ReadFileSet)), ("n_read_file_sets", C.c_uint32),
                ("coverage_group_of_tid", C.POINTER(C.c_uint32)), ("n_targets", C.c_uint32),
                ("use_base_repeat", C.c_uint32), ("use_read_pos", C.c_uint32), ("shard_rank", C.c_uint32),
                ("shard_count", C.c_uint32), ("base_quality_cutoff", C.c_uint32),
                ("preprocess_stage", C.c_uint32), ("unmatched_end_minimum_read_length", C.c_uint32),
                ("require_match_fraction", C.c_double), ("shard_lo", C.c_uint64), ("shard_hi", C.c_uint64),
                ("staging", C.c_uint32), ("reserved", C.c_uint32), ("user_evidence_gd", C.c_char_p)]


class _SynthReadSet(C.Structure):
    _fields_ = [("name", C.c_char_p), ("paired", C.c_uint32), ("read_len", C.c_uint32),
                ("coverage", C.c_double), ("frag_mean", C.c_double), ("frag_sd", C.c_double)]


class _SynthSpec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("contig_lens", C.POINTER(C.c_uint32)), ("n_contigs", C.c_uint32),
                ("contig_prefix", C.c_char_p), ("fasta", C.c_char_p), ("sets", C.POINTER(_SynthReadSet)),
                ("n_sets", C.c_uint32), ("n_polymorphic", C.c_uint32), ("n_fixed", C.c_uint32), ("n_gaps", C.c_uint32),
                ("min_freq_ppm", C.c_uint32), ("max_freq_ppm", C.c_uint32), ("window_lo", C.c_uint64), ("window_hi", C.c_uint64)]


class _StreamInfo(C.Structure):
    _fields_ = [("n_base", C.c_uint64), ("n_ins", C.c_uint64), ("n_score_records", C.c_uint64),
                ("n_hist_records", C.c_uint64), ("n_reads", C.c_uint64), ("n_score_padded", C.c_uint64),
                ("bytes_host", C.c_uint64),
                ("n_targets", C.c_uint32), ("pinned", C.c_uint32), ("device_built", C.c_uint32), ("hist_compact", C.c_uint32),
                ("hist_record_bytes", C.c_uint32), ("side_stride", C.c_uint32),
                ("n_side", C.c_uint64), ("base_quality_cutoff", C.c_uint32), ("hot_mapq", C.c_uint32),
                ("table_q_lo", C.c_uint32), ("table_n_q", C.c_uint32), ("table_n_st", C.c_uint32), ("table_words", C.c_uint32),
                ("score_rec", C.POINTER(C.c_uint32)), ("side_rec", C.POINTER(C.c_uint32)), ("side_off", C.POINTER(C.c_uint32)),
                ("score_off", C.POINTER(C.c_uint64)),
                ("hist_rec", C.c_void_p), ("hist_off", C.POINTER(C.c_uint64)),
                ("slot_ref", C.POINTER(C.c_uint8)), ("ins_parent", C.POINTER(C.c_uint64)),
                ("ins_count", C.POINTER(C.c_uint32)), ("round_slot", C.POINTER(C.c_uint32)), ("n_rounds", C.c_uint64),
                ("score_cnt", C.POINTER(C.c_uint32)), ("round_off", C.POINTER(C.c_uint64)),
                ("hist16", C.POINTER(C.c_uint16)), ("hist_exc", C.POINTER(C.c_uint32)), ("n_hist16", C.c_uint64), ("n_hist_exc", C.c_uint64),
                ("score16", C.POINTER(C.c_uint16)), ("score_exc", C.POINTER(C.c_uint32)), ("score_exc_off", C.POINTER(C.c_uint32)),
                ("n_score_exc", C.c_uint64)]


class _ScoreParams(C.Structure):
    _fields_ = [("mutation_cutoff", C.c_double), ("polymorphism_cutoff", C.c_double),
                ("polymorphism_precision_decimal", C.c_double), ("polymorphism_precision_places", C.c_uint32),
                ("base_quality_cutoff", C.c_uint32), ("total_reference_length", C.c_uint64),
                ("flags", C.c_uint32), ("reserved", C.c_uint32)]


#: numpy view of ``brq_column`` (96 bytes per slot)
COLUMN_DTYPE = np.dtype([("ll", "<f8", 5), ("consensus_score", "<f8"), ("variant_score", "<f8"),
                         ("redundant", "<f8", 2), ("unique", "<u4", 2), ("raw_redundant", "<u4", 2),
                         ("n", "<u4"), ("bits", "<u4")])
assert COLUMN_DTYPE.itemsize == 96

CO_BASE_PREDICTED, CO_UNIQUE_ONLY, CO_EMIT, CO_RECHECK, CO_FIT = 1 << 12, 1 << 13, 1 << 14, 1 << 15, 1 << 24
SCORE_FIT_ALL_COLUMNS = 1
SCORE_POLYMORPHISM_PREDICTION = 2
SCORE_KEEP_BOUNDS = 4

_lib = None


class CoverageFit(C.Structure):  # brq_coverage_fit
    _fields_ = [("average", C.c_double), ("variance", C.c_double), ("relative_variance", C.c_double),
                ("nbinom_size_parameter", C.c_double), ("nbinom_mean_parameter", C.c_double),
                ("deletion_coverage_propagation_cutoff", C.c_double), ("censor_start", C.c_uint32), ("censor_end", C.c_uint32)]


class RaFilterOptions(C.Structure):  # brq_ra_filter_options: the members of breseq::Settings test_RA_evidence reads
    _fields_ = [("polymorphism_prediction", C.c_int32), ("mutation_log10_e_value_cutoff", C.c_double),
                ("consensus_frequency_cutoff", C.c_double),
                ("consensus_minimum_variant_coverage", C.c_uint32), ("consensus_minimum_total_coverage", C.c_uint32),
                ("consensus_minimum_variant_coverage_each_strand", C.c_uint32), ("consensus_minimum_total_coverage_each_strand", C.c_uint32),
                ("consensus_reject_indel_homopolymer_length", C.c_uint32), ("consensus_reject_surrounding_homopolymer_length", C.c_uint32),
                ("polymorphism_log10_e_value_cutoff", C.c_double), ("polymorphism_frequency_cutoff", C.c_double),
                ("polymorphism_minimum_variant_coverage", C.c_uint32), ("polymorphism_minimum_total_coverage", C.c_uint32),
                ("polymorphism_minimum_variant_coverage_each_strand", C.c_uint32), ("polymorphism_minimum_total_coverage_each_strand", C.c_uint32),
                ("polymorphism_reject_indel_homopolymer_length", C.c_uint32), ("polymorphism_reject_surrounding_homopolymer_length", C.c_uint32),
                ("polymorphism_fisher_strand_p_value_cutoff", C.c_double), ("polymorphism_ks_quality_p_value_cutoff", C.c_double),
                ("polymorphism_no_indels", C.c_int32)]


def load_library():
    """dlopen ``libbrq.so``; fails loudly when the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BrqError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback for the CUDA path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    lib.brq_create.restype = C.c_void_p
    lib.brq_create.argtypes = [P(_Config)]
    lib.brq_destroy.argtypes = [C.c_void_p]
    lib.brq_last_error.restype = C.c_char_p
    lib.brq_last_error.argtypes = [C.c_void_p]
    lib.brq_version.restype = C.c_char_p
    sig = {
        "brq_stage_bam": [C.c_void_p, C.c_char_p, C.c_char_p, P(_StageOptions)],
        "brq_synth_write": [C.c_void_p, P(_SynthSpec), C.c_char_p, C.c_char_p],
        "brq_stage_synthetic": [C.c_void_p, P(_SynthSpec), P(_StageOptions)],
        "brq_stream": [C.c_void_p, P(_StreamInfo)],
        "brq_stream_summary": [C.c_void_p, P(_StreamInfo)],
        "brq_pin_reads": [C.c_void_p],
        "brq_max_coverage_depth": [C.c_void_p, P(C.c_uint64)],
        "brq_set_min_coverage_depth": [C.c_void_p, C.c_uint64],
        "brq_restage": [C.c_void_p],
        "brq_synth_shard_bounds": [C.c_void_p, P(_SynthSpec), C.c_uint32, P(C.c_uint64)],
        "brq_bam_shard_bounds": [C.c_void_p, C.c_char_p, C.c_uint32, P(C.c_uint64)],
        "brq_upload": [C.c_void_p],
        "brq_sync": [C.c_void_p],
        "brq_error_count": [C.c_void_p, C.c_char_p, C.c_int, C.c_int],
        "brq_hist_device": [C.c_void_p, P(C.c_void_p), P(C.c_uint64), P(C.c_void_p), P(C.c_uint64)],
        "brq_preprocess_read_starts": [C.c_void_p, P(P(C.c_uint64)), P(C.c_uint32)],
        "brq_hist_download": [C.c_void_p, P(P(C.c_uint64)), P(C.c_uint64), P(P(C.c_uint64)), P(C.c_uint64), P(C.c_uint64)],
        "brq_derive_error_table": [C.c_void_p],
        "brq_error_table": [C.c_void_p, P(P(C.c_double)), P(C.c_uint64)],
        "brq_write_error_count_files": [C.c_void_p, C.c_char_p, C.c_char_p, P(C.c_char_p), C.c_uint32, C.c_int, C.c_int, C.c_char_p],
        "brq_load_error_table": [C.c_void_p, C.c_char_p],
        "brq_score_columns": [C.c_void_p, P(_ScoreParams)],
        "brq_columns_download": [C.c_void_p, P(C.c_void_p), P(C.c_uint64), P(P(C.c_uint32)), P(C.c_uint32)],
        "brq_columns_device": [C.c_void_p, P(C.c_void_p), P(C.c_uint64)],
        "brq_write_evidence": [C.c_void_p, C.c_char_p, P(C.c_double), P(C.c_double), C.c_uint32, C.c_int,
                               P(C.c_uint64), P(C.c_uint64), P(C.c_uint64)],
        "brq_d2h_bytes": [C.c_void_p, P(C.c_uint64), C.c_int],
        "brq_cuda_stream": [C.c_void_p, P(C.c_void_p)],
        "brq_evidence_export": [C.c_void_p, P(C.c_double), C.c_uint32, P(C.c_void_p), P(C.c_uint64)],
        "brq_write_evidence_merged": [C.c_void_p, P(C.c_void_p), P(C.c_uint64), C.c_uint32, C.c_char_p, P(C.c_double), P(C.c_double),
                                      C.c_uint32, C.c_int, P(C.c_uint64), P(C.c_uint64), P(C.c_uint64)],
        "brq_write_per_position_file": [C.c_void_p, C.c_char_p, P(C.c_double), C.c_uint32],
        "brq_write_coverage_tsv": [C.c_void_p, C.c_char_p],
        "brq_write_per_position_counts": [C.c_void_p, C.c_char_p, C.c_char_p],
        "brq_write_coverage_table": [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint32, C.c_int, C.c_int, C.c_int],
        "brq_write_coverage_table_with_average": [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_double],
        "brq_fit_coverage_distribution": [C.c_void_p, C.c_uint32, C.c_double, P(CoverageFit)],
        "brq_fit_coverage_file": [C.c_void_p, C.c_char_p, C.c_double, P(CoverageFit)],
        "brq_test_ra_evidence": [C.c_void_p, C.c_char_p, C.c_char_p, P(RaFilterOptions), C.c_char_p, P(C.c_uint32)],
        "brq_predict_ra_mutations": [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, P(C.c_uint32)],
        "brq_hist_exchange_export": [C.c_void_p, C.c_void_p, P(C.c_uint64)],
        "brq_hist_exchange_attach": [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32],
        "brq_run_error_count": [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, P(C.c_char_p), C.c_uint32,
                                C.c_int, C.c_int, C.c_char_p, P(_StageOptions)],
        "brq_run_identify_mutations": [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, P(C.c_double),
                                       P(C.c_double), C.c_uint32, P(_ScoreParams), C.c_int, P(_StageOptions)],
        "brq_launch_count": [],
        "brq_event_record": [C.c_void_p, C.c_int],
        "brq_event_elapsed_ms": [C.c_void_p, C.c_int, C.c_int, P(C.c_float)],
        "brq_kernel_ms": [C.c_void_p, P(C.c_float), P(C.c_float), P(C.c_float), P(C.c_float)],
        "brq_score_phase_ms": [C.c_void_p, P(C.c_float), P(C.c_float)],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.brq_binomial_frequency_bounds.argtypes = [C.c_double, C.c_double, C.c_double, P(C.c_double), P(C.c_double)]
    lib.brq_binomial_frequency_bounds.restype = None
    lib.brq_fisher_strand_p_value.argtypes = [C.c_uint32] * 4
    lib.brq_fisher_strand_p_value.restype = C.c_double
    lib.brq_ra_filter_defaults.argtypes = [C.c_int, P(RaFilterOptions)]
    lib.brq_ra_filter_defaults.restype = None
    _lib = lib
    return lib


#: every symbol include/brq.h declares (checked by the CPU test-suite)
EXPORTS = ["brq_create", "brq_destroy", "brq_last_error", "brq_version", "brq_stage_bam", "brq_synth_write",
           "brq_stage_synthetic", "brq_stream", "brq_upload", "brq_sync", "brq_error_count", "brq_hist_device",
           "brq_hist_download", "brq_derive_error_table", "brq_error_table", "brq_write_error_count_files",
           "brq_load_error_table", "brq_score_columns", "brq_columns_download", "brq_columns_device",
           "brq_write_evidence", "brq_cuda_stream", "brq_evidence_export", "brq_write_evidence_merged", "brq_d2h_bytes", "brq_write_per_position_file", "brq_write_coverage_tsv", "brq_run_error_count", "brq_run_identify_mutations", "brq_launch_count", "brq_kernel_ms",
           "brq_event_record", "brq_event_elapsed_ms", "brq_score_phase_ms", "brq_preprocess_read_starts",
           "brq_stream_summary", "brq_max_coverage_depth", "brq_set_min_coverage_depth", "brq_pin_reads", "brq_restage", "brq_synth_shard_bounds", "brq_bam_shard_bounds",
           "brq_fit_coverage_distribution", "brq_fit_coverage_file", "brq_hist_exchange_export", "brq_hist_exchange_attach", "brq_write_coverage_table", "brq_write_per_position_counts",
           "brq_ra_filter_defaults", "brq_test_ra_evidence", "brq_predict_ra_mutations", "brq_binomial_frequency_bounds", "brq_fisher_strand_p_value", "brq_write_coverage_table_with_average"]


def _b(s):
    return None if s is None else (s if isinstance(s, bytes) else str(s).encode())


def _str_array(items):
    arr = (C.c_char_p * max(1, len(items)))()
    for i, s in enumerate(items):
        arr[i] = _b(s)
    return arr


def binomial_frequency_bounds(k, n, alpha=0.05):
    """(lower, upper): exact one-sided confidence bounds on k / n (stats.cpp:2394-2414)."""
    lo, hi = C.c_double(), C.c_double()
    load_library().brq_binomial_frequency_bounds(k, n, alpha, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def fisher_strand_p_value(minor_top, minor_bottom, major_top, major_bottom):
    """fisher_strand_p_value of an RA row from the strand counts of its two alleles (stats.cpp:2144-2171)."""
    return load_library().brq_fisher_strand_p_value(minor_top, minor_bottom, major_top, major_bottom)


class SynthSpec:
    """Synthetic aligned reads (SURVEY.md section 8d value distributions)."""

    def __init__(self, seed, read_sets, contig_lens=None, fasta=None, contig_prefix="contig",
                 n_polymorphic=60, n_fixed=10, n_gaps=2, min_freq_ppm=50000, max_freq_ppm=500000, window=(0, 0)):
        self.keep = []
        sets = (_SynthReadSet * len(read_sets))()
        for i, rs in enumerate(read_sets):
            name = _b(rs["name"])
            self.keep.append(name)
            sets[i] = _SynthReadSet(name, int(rs.get("paired", False)), int(rs["read_len"]), float(rs["coverage"]),
                                    float(rs.get("frag_mean", 400)), float(rs.get("frag_sd", 40)))
        lens = None
        if contig_lens is not None:
            lens = (C.c_uint32 * len(contig_lens))(*contig_lens)
        self.keep += [sets, lens]
        self.c = _SynthSpec(seed, lens, 0 if contig_lens is None else len(contig_lens), _b(contig_prefix), _b(fasta),
                            sets, len(read_sets), n_polymorphic, n_fixed, n_gaps, min_freq_ppm, max_freq_ppm, window[0], window[1])
        self.read_sets = read_sets

    def read_file_sets(self):
        """The run's cReadFileSets: one per read group, two files when paired."""
        return [(rs["name"], 2 if rs.get("paired") else 1) for rs in self.read_sets]


def _stage_options(seq_ids=None, read_file_sets=None, coverage_groups=None, use_base_repeat=False, use_read_pos=False,
                   shard=(0, 1), base_quality_cutoff=3, preprocess_stage=False, unmatched_end_minimum_read_length=50,
                   require_match_fraction=0.9, shard_bounds=None, staging="auto", user_evidence_gd=None):
    keep = []
    o = _StageOptions()
    if seq_ids:
        arr = _str_array(list(seq_ids))
        keep.append(arr)
        o.seq_ids, o.n_seq_ids = arr, len(seq_ids)
    if read_file_sets:
        names = [_b(n) for n, _ in read_file_sets]
        arr = (_ReadFileSet * len(read_file_sets))()
        for i, (n, k) in enumerate(read_file_sets):
            arr[i] = _ReadFileSet(names[i], k)
        keep += [names, arr]
        o.read_file_sets, o.n_read_file_sets = arr, len(read_file_sets)
    if coverage_groups is not None:
        arr = (C.c_uint32 * len(coverage_groups))(*coverage_groups)
        keep.append(arr)
        o.coverage_group_of_tid, o.n_targets = arr, len(coverage_groups)
    o.use_base_repeat = int(use_base_repeat)
    o.use_read_pos = int(use_read_pos)
    o.shard_rank, o.shard_count = shard
    o.base_quality_cutoff = base_quality_cutoff
    o.preprocess_stage = int(preprocess_stage)
    o.unmatched_end_minimum_read_length = unmatched_end_minimum_read_length
    o.require_match_fraction = require_match_fraction
    if shard_bounds is not None:
        o.shard_lo, o.shard_hi = shard_bounds
    o.staging = {"auto": 0, "host": 1, "device": 2}[staging]
    if user_evidence_gd:
        path = _b(user_evidence_gd)
        keep.append(path)
        o.user_evidence_gd = path
    return o, keep


def slot_ranges(stream):
    """(first word, record count) of every slot in ``score_rec`` (round-major, lane-interleaved: see record_positions)."""
    return stream["score_off"][:-1].astype(np.int64), stream["score_cnt"].astype(np.int64)


def record_positions(stream):
    """Word index in ``score_rec`` of every record, slot by slot in stream order: record j of a slot whose first word
    is b sits at b + (j >> 3) * 256 + ((j >> 2) & 1) * 128 + (j & 3) (csrc/brq_types.h: score_index)."""
    beg, cnt = slot_ranges(stream)
    j = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    return np.repeat(beg, cnt) + (j >> 3) * 256 + ((j >> 2) & 1) * 128 + (j & 3)


def decode_score_records(stream):
    """The staged scoring records as plain arrays, one entry per record in stream order (padding dropped).

    ``score_rec`` holds table-coordinate words (csrc/brq_types.h): this undoes the encoding.  Returns a dict with
    ``slot``, ``unique``, ``top``, ``scores`` (bool), ``x1``, and for scoring records ``obs``, ``qual``, ``read_set``,
    ``mapq``, ``match`` (-1 / False where the stream does not keep the value)."""
    g = stream["geometry"]
    beg, cnt = slot_ranges(stream)
    n_slots = len(beg)
    d = stream["score_rec"][record_positions(stream)].astype(np.int64)
    slot = np.repeat(np.arange(n_slots), cnt)
    kind = d >> 30
    n = len(d)
    out = {"slot": slot, "kind": kind, "unique": kind != 3, "top": (d >> 13) & 1, "scores": (kind == 0) | (kind == 2),
           "x1": np.ones(n, np.int64), "obs": np.full(n, -1), "qual": np.full(n, -1), "read_set": np.full(n, -1),
           "mapq": np.full(n, -1), "match": np.zeros(n, bool)}
    hot = kind == 0
    t = (d[hot] >> 16) & 0xFF
    out["obs"][hot] = (d[hot] >> 24) & 7
    out["qual"][hot] = g["q_lo"] + t % max(1, g["n_q"])
    out["read_set"][hot] = (t // max(1, g["n_q"])) >> 1
    out["mapq"][hot] = g["hot_mapq"]
    out["match"][hot] = ((d[hot] >> 28) & 1) == 1
    # side list: per slot, the entries of its very redundant records first, then those of its cold records, in order
    side, soff = stream["side_rec"].astype(np.int64)[::stream["side_stride"]], stream["side_off"].astype(np.int64)
    x1 = (d >> 16) & 0x1FF
    big = (kind == 3) & (x1 == 0x1FF)
    cold = kind == 2
    uses = big | cold
    used = np.bincount(slot[uses], minlength=n_slots)   # entries every slot uses; its range is padded to an even count
    assert np.array_equal(np.diff(soff), (used + 1) & ~1), "side_off does not match the records that use the side list"
    before = np.cumsum(used) - used                      # side-using records before each slot
    rank_in_slot = (np.cumsum(uses) - 1)[uses] - before[slot[uses]]
    entry = side[soff[slot[uses]] + rank_in_slot] if uses.any() else np.zeros(0, np.int64)
    pad_at = soff[:-1] + used
    assert np.all(side[pad_at[used % 2 == 1]] == 0xFFFFFFFF), "odd ranges end in a pad entry"
    e_big, e_cold = entry[big[uses]], entry[cold[uses]]
    assert np.all(e_big >> 31 == 1) and np.all(e_cold >> 31 == 0)
    out["x1"][kind == 3] = x1[kind == 3]
    out["x1"][big] = e_big & 0x7FFFFFFF
    out["obs"][cold] = e_cold & 7
    out["qual"][cold] = (e_cold >> 3) & 127
    out["read_set"][cold] = (e_cold >> 11) & 31
    out["mapq"][cold] = (e_cold >> 16) & 255
    out["match"][cold] = ((e_cold >> 27) & 1) == 1
    assert np.array_equal(out["top"][cold], (e_cold >> 10) & 1)
    return out


class Context:
    """One pileup context = one GPU (``device`` >= 0) or host-only staging (``device`` = -1)."""

    def __init__(self, device=0, threads=0):
        self.lib = load_library()
        cfg = _Config(device, threads)
        self.h = self.lib.brq_create(C.byref(cfg))
        if not self.h:
            raise BrqError("brq_create failed")
        self.device = device
        err = self.lib.brq_last_error(self.h)
        if device >= 0 and err:
            msg = err.decode()
            self.close()
            raise BrqError(msg)

    def close(self):
        if getattr(self, "h", None):
            self.lib.brq_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise BrqError(self.lib.brq_last_error(self.h).decode())

    # ---- staging
    def stage_bam(self, bam, fasta, **kw):
        o, keep = _stage_options(**kw)
        self._check(self.lib.brq_stage_bam(self.h, _b(bam), _b(fasta), C.byref(o)))

    def stage_synthetic(self, spec, **kw):
        o, keep = _stage_options(**kw)
        self._check(self.lib.brq_stage_synthetic(self.h, C.byref(spec.c), C.byref(o)))

    def synth_write(self, spec, bam_out, fasta_out):
        self._check(self.lib.brq_synth_write(self.h, C.byref(spec.c), _b(bam_out), _b(fasta_out)))

    def stream_summary(self):
        """Counts and geometry of the staged stream (no array views: nothing is copied from HBM)."""
        info = _StreamInfo()
        self._check(self.lib.brq_stream_summary(self.h, C.byref(info)))
        return {"n_base": info.n_base, "n_ins": info.n_ins, "n_score": info.n_score_records, "n_hist": info.n_hist_records,
                "n_reads": info.n_reads, "bytes_host": info.bytes_host, "n_score_padded": info.n_score_padded, "n_side": info.n_side,
                "n_rounds": info.n_rounds, "device_built": bool(info.device_built), "hist_compact": bool(info.hist_compact),
                "n_hist16": info.n_hist16, "n_hist_exc": info.n_hist_exc, "hist_record_bytes": info.hist_record_bytes,
                "side_stride": info.side_stride}

    def max_coverage_depth(self):
        d = C.c_uint64()
        self._check(self.lib.brq_max_coverage_depth(self.h, C.byref(d)))
        return d.value

    def set_min_coverage_depth(self, depth):
        """Floor of the coverage histogram's depth axis (the maximum over the ranks of a sharded run, so the ranks can sum)."""
        self.lib.brq_set_min_coverage_depth(self.h, int(depth))

    def pin_reads(self):
        """Page-lock the host copy of the reads (device staging then copies them at PCIe speed)."""
        self._check(self.lib.brq_pin_reads(self.h))

    def restage(self):
        """Stage again from the host copy of the reads: H2D + expansion on the device (or host staging)."""
        self._check(self.lib.brq_restage(self.h))

    def synth_shard_bounds(self, spec, n_shards):
        b = (C.c_uint64 * (n_shards + 1))()
        self._check(self.lib.brq_synth_shard_bounds(self.h, C.byref(spec.c), n_shards, b))
        return list(b)

    def bam_shard_bounds(self, bam, n_shards):
        b = (C.c_uint64 * (n_shards + 1))()
        self._check(self.lib.brq_bam_shard_bounds(self.h, _b(bam), n_shards, b))
        return list(b)

    def stream(self):
        info = _StreamInfo()
        self._check(self.lib.brq_stream(self.h, C.byref(info)))
        n_slots = info.n_base + info.n_ins

        def view(ptr, n, dtype):
            if n == 0:
                return np.zeros(0, dtype)
            return np.ctypeslib.as_array(ptr, shape=(int(n),)).view(dtype)
        return {
            "n_base": info.n_base, "n_ins": info.n_ins, "n_score": info.n_score_records, "n_hist": info.n_hist_records,
            "n_reads": info.n_reads, "bytes_host": info.bytes_host, "pinned": bool(info.pinned), "n_targets": info.n_targets,
            "device_built": bool(info.device_built), "n_hist16": info.n_hist16, "n_hist_exc": info.n_hist_exc, "n_rounds": info.n_rounds,
            "n_score_padded": info.n_score_padded,
            "score_rec": view(info.score_rec, info.n_score_padded, np.uint32),
            "score_off": view(info.score_off, n_slots + 1, np.uint64),
            "n_side": info.n_side, "side_stride": info.side_stride,
            "side_rec": view(info.side_rec, info.n_side * info.side_stride, np.uint32),
            "side_off": view(info.side_off, n_slots + 1, np.uint32),
            "geometry": {"base_quality_cutoff": info.base_quality_cutoff, "hot_mapq": info.hot_mapq, "q_lo": info.table_q_lo,
                         "n_q": info.table_n_q, "n_st": info.table_n_st, "words": info.table_words},
            "hist_rec": view(C.cast(info.hist_rec, C.POINTER(C.c_uint64 if info.hist_record_bytes == 8 else C.c_uint32)),
                             info.n_hist_records, np.uint64 if info.hist_record_bytes == 8 else np.uint32),
            "hist_off": view(info.hist_off, info.n_base + 1, np.uint64),
            # compact form the device reads (None when the stream has 8-byte histogram records): csrc/brq_types.h
            "hist16": view(info.hist16, info.n_hist16, np.uint16) if info.hist16 else None,
            "hist_exc": view(info.hist_exc, info.n_hist_exc, np.uint32) if info.hist16 else None,
            # transfer form of score_rec (None when not built): low halves, exception words, CSR per (round, lane)
            "score16": view(info.score16, info.n_score_padded, np.uint16) if info.score16 else None,
            "score_exc": view(info.score_exc, info.n_score_exc, np.uint32) if info.score16 else None,
            "score_exc_off": view(info.score_exc_off, info.n_rounds * 32 + 1, np.uint32) if info.score16 else None,
            "slot_ref": view(info.slot_ref, n_slots, np.uint8),
            "ins_parent": view(info.ins_parent, info.n_ins, np.uint64),
            "ins_count": view(info.ins_count, info.n_ins, np.uint32),
            "round_slot": view(info.round_slot, info.n_rounds * 32, np.uint32),
            "score_cnt": view(info.score_cnt, n_slots, np.uint32),
            "round_off": view(info.round_off, info.n_rounds + 1, np.uint64),
        }

    def upload(self):
        self._check(self.lib.brq_upload(self.h))

    def sync(self):
        self._check(self.lib.brq_sync(self.h))

    # ---- pass 1
    def error_count(self, covariates, do_coverage=True, do_errors=True):
        self._check(self.lib.brq_error_count(self.h, _b(covariates), int(do_coverage), int(do_errors)))

    def preprocess_read_starts(self):
        """Per BAM tid, the position-strand combinations (without, with) a read start inside the junction read-end bound
        (stream staged with ``preprocess_stage=True``; error_count.cpp:157-166, 191-194)."""
        p, n = C.POINTER(C.c_uint64)(), C.c_uint32()
        self._check(self.lib.brq_preprocess_read_starts(self.h, C.byref(p), C.byref(n)))
        return np.ctypeslib.as_array(p, shape=(n.value * 2,)).copy().reshape(n.value, 2)

    def hist_device(self):
        c, n, v, m = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64()
        self._check(self.lib.brq_hist_device(self.h, C.byref(c), C.byref(n), C.byref(v), C.byref(m)))
        return c.value, n.value, v.value, m.value

    def hist_download(self):
        c, v = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint64)()
        n, stride, groups = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.brq_hist_download(self.h, C.byref(c), C.byref(n), C.byref(v), C.byref(stride), C.byref(groups)))
        counts = np.ctypeslib.as_array(c, shape=(n.value,)).copy()
        cov = np.ctypeslib.as_array(v, shape=(stride.value * groups.value,)).copy().reshape(groups.value, stride.value)
        return counts, cov

    def derive_error_table(self):
        self._check(self.lib.brq_derive_error_table(self.h))

    def error_table(self):
        t, n = C.POINTER(C.c_double)(), C.c_uint64()
        self._check(self.lib.brq_error_table(self.h, C.byref(t), C.byref(n)))
        return np.ctypeslib.as_array(t, shape=(n.value,)).copy()

    def write_error_count_files(self, output_dir, error_rates_file=None, readfiles=(), do_coverage=True, do_errors=True,
                                counts_dump=None):
        arr = _str_array(list(readfiles))
        self._check(self.lib.brq_write_error_count_files(self.h, _b(output_dir), _b(error_rates_file), arr, len(readfiles),
                                                         int(do_coverage), int(do_errors), _b(counts_dump)))

    def load_error_table(self, path):
        self._check(self.lib.brq_load_error_table(self.h, _b(path)))

    # ---- pass 2
    @staticmethod
    def score_params(mutation_cutoff=10.0, polymorphism_cutoff=2.0, precision_decimal=1e-6, precision_places=8,
                     base_quality_cutoff=3, total_reference_length=0, fit_all_columns=False, polymorphism_prediction=False, keep_bounds=False):
        return _ScoreParams(mutation_cutoff, polymorphism_cutoff, precision_decimal, precision_places, base_quality_cutoff,
                            total_reference_length, (SCORE_FIT_ALL_COLUMNS if fit_all_columns else 0) |
                            (SCORE_POLYMORPHISM_PREDICTION if polymorphism_prediction else 0) | (SCORE_KEEP_BOUNDS if keep_bounds else 0), 0)

    def score_columns(self, params=None):
        p = params or self.score_params()
        self._check(self.lib.brq_score_columns(self.h, C.byref(p)))

    def columns_download(self):
        cols, n, fl, nf = C.c_void_p(), C.c_uint64(), C.POINTER(C.c_uint32)(), C.c_uint32()
        self._check(self.lib.brq_columns_download(self.h, C.byref(cols), C.byref(n), C.byref(fl), C.byref(nf)))
        buf = (C.c_char * (n.value * 96)).from_address(cols.value)
        columns = np.frombuffer(buf, dtype=COLUMN_DTYPE).copy()
        flagged = np.ctypeslib.as_array(fl, shape=(nf.value,)).copy() if nf.value else np.zeros(0, np.uint32)
        return columns, flagged

    def columns_device(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self.lib.brq_columns_device(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def write_evidence(self, gd_file, deletion_propagation_cutoff, deletion_seed_cutoff, skip_missing_coverage_prediction=False):
        n = len(deletion_propagation_cutoff)
        prop = (C.c_double * n)(*deletion_propagation_cutoff)
        seed = (C.c_double * n)(*deletion_seed_cutoff)
        ra, mc, un = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.brq_write_evidence(self.h, _b(gd_file), prop, seed, n, int(skip_missing_coverage_prediction),
                                                C.byref(ra), C.byref(mc), C.byref(un)))
        return {"RA": ra.value, "MC": mc.value, "UN": un.value}

    def evidence_export(self, deletion_propagation_cutoff):
        """This context's share of the evidence of a run sharded by reference range, as bytes (see write_evidence_merged)."""
        n = len(deletion_propagation_cutoff)
        prop = (C.c_double * n)(*deletion_propagation_cutoff)
        p, size = C.c_void_p(), C.c_uint64()
        self._check(self.lib.brq_evidence_export(self.h, prop, n, C.byref(p), C.byref(size)))
        return C.string_at(p.value, size.value)

    def write_evidence_merged(self, gd_file, shards, deletion_propagation_cutoff, deletion_seed_cutoff,
                              skip_missing_coverage_prediction=False):
        """``ra_mc_evidence.gd`` of a sharded run from the shares of all its contexts (any order)."""
        n = len(deletion_propagation_cutoff)
        prop = (C.c_double * n)(*deletion_propagation_cutoff)
        seed = (C.c_double * n)(*deletion_seed_cutoff)
        bufs = [C.create_string_buffer(b, len(b)) for b in shards]
        ptrs = (C.c_void_p * len(bufs))(*[C.cast(b, C.c_void_p) for b in bufs])
        sizes = (C.c_uint64 * len(bufs))(*[len(b) for b in shards])
        ra, mc, un = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.brq_write_evidence_merged(self.h, ptrs, sizes, len(bufs), _b(gd_file), prop, seed, n,
                                                       int(skip_missing_coverage_prediction), C.byref(ra), C.byref(mc), C.byref(un)))
        return {"RA": ra.value, "MC": mc.value, "UN": un.value}

    def cuda_stream(self):
        """cudaStream_t of the context as an integer (torch.cuda.ExternalStream takes it)."""
        p = C.c_void_p()
        self._check(self.lib.brq_cuda_stream(self.h, C.byref(p)))
        return p.value or 0

    def d2h_bytes(self, reset=False):
        n = C.c_uint64()
        self.lib.brq_d2h_bytes(self.h, C.byref(n), int(reset))
        return n.value

    def write_per_position_file(self, path, deletion_propagation_cutoff):
        """The reference's per-position debug file (identify_mutations.cpp:1693-1733)."""
        n = len(deletion_propagation_cutoff)
        prop = (C.c_double * n)(*deletion_propagation_cutoff)
        self._check(self.lib.brq_write_per_position_file(self.h, _b(path), prop, n))

    def write_coverage_tsv(self, pattern):
        """``<seq>.coverage.tsv`` of --predict-copy-number; '@' in ``pattern`` becomes the target name."""
        self._check(self.lib.brq_write_coverage_tsv(self.h, _b(pattern)))

    def write_per_position_counts(self, covariates, path):
        """``error_counts.tab`` of a covariate string with ref_pos: every position's non-empty bins (error_count.cpp:193-198)."""
        self._check(self.lib.brq_write_per_position_counts(self.h, _b(covariates), _b(path)))

    def write_coverage_table(self, region, path, resolution=0, total_only=False, csv=False, per_read_group=False, reference_average=None):
        """BAM2COV's table for ``seq_id:start-end`` of the staged BAM (coverage_output.cpp:190-283, 307-470); ``reference_average``
        (BAM2COV -a): the sequence's fit average, printed as one more '#' line."""
        if reference_average is None:
            self._check(self.lib.brq_write_coverage_table(self.h, _b(region), _b(path), resolution, int(total_only), int(csv), int(per_read_group)))
        else:
            self._check(self.lib.brq_write_coverage_table_with_average(self.h, _b(region), _b(path), resolution, int(total_only), int(csv),
                                                                       int(per_read_group), float(reference_average)))

    @staticmethod
    def _fit_dict(f):
        return {"average": f.average, "variance": f.variance, "relative_variance": f.relative_variance,
                "nbinom_size_parameter": f.nbinom_size_parameter, "nbinom_mean_parameter": f.nbinom_mean_parameter,
                "deletion_coverage_propagation_cutoff": f.deletion_coverage_propagation_cutoff,
                "censor_start": f.censor_start, "censor_end": f.censor_end}

    def fit_coverage_distribution(self, coverage_group, deletion_propagation_pr_cutoff):
        """Censored negative-binomial fit of a coverage group's unique-only coverage histogram (the last error_count's) and
        the deletion-propagation cutoff (CoverageDistribution::fit, coverage_distribution.cpp:115-400)."""
        f = CoverageFit()
        self._check(self.lib.brq_fit_coverage_distribution(self.h, coverage_group, C.c_double(deletion_propagation_pr_cutoff), C.byref(f)))
        return self._fit_dict(f)

    def fit_coverage_file(self, path, deletion_propagation_pr_cutoff):
        """The same fit from a ``<group>.unique_only_coverage_distribution.tab`` (host only)."""
        f = CoverageFit()
        self._check(self.lib.brq_fit_coverage_file(self.h, _b(path), C.c_double(deletion_propagation_pr_cutoff), C.byref(f)))
        return self._fit_dict(f)

    def ra_filter_defaults(self, polymorphism_prediction=False):
        """The thresholds breseq's Settings hold for the mode without further options, as a dict."""
        o = RaFilterOptions()
        self.lib.brq_ra_filter_defaults(int(bool(polymorphism_prediction)), C.byref(o))
        return {name: getattr(o, name) for name, _ in RaFilterOptions._fields_}

    def test_RA_evidence(self, gd_in, fasta, gd_out, polymorphism_prediction=False, **settings):
        """The Output stage's filter over the RA rows of ``gd_in`` (identify_mutations.cpp:687-749): ``gd_out`` holds the rows
        that stay, annotated with prediction= / *_reject=.  ``settings``: members of breseq::Settings by name, over the mode's
        defaults.  Returns the counts {rows, consensus, polymorphism, rejected_kept, deleted}.  Host only."""
        o = RaFilterOptions()
        self.lib.brq_ra_filter_defaults(int(bool(polymorphism_prediction)), C.byref(o))
        names = {name for name, _ in RaFilterOptions._fields_}
        for k, v in settings.items():
            if k not in names:
                raise BrqError("test_RA_evidence: breseq::Settings has no member %r that the filter reads" % k)
            setattr(o, k, v)
        n = (C.c_uint32 * 5)()
        self._check(self.lib.brq_test_ra_evidence(self.h, _b(gd_in), _b(fasta), C.byref(o), _b(gd_out), n))
        return dict(zip(("rows", "consensus", "polymorphism", "rejected_kept", "deleted"), list(n)))

    def predict_ra_mutations(self, gd_in, fasta, gd_out, polymorphism_prediction=False, targeted_sequencing=False,
                             call_mutations_overlapping_missing_coverage=False):
        """SNP / DEL / INS / SUB rows from the RA rows test_RA_evidence() accepted (mutation_predictor.cpp:1955-2211): ``gd_out`` =
        the mutation rows in front of the evidence rows.  Returns the counts {SNP, DEL, INS, SUB, ra_marked_deleted}.  Host only."""
        n = (C.c_uint32 * 5)()
        self._check(self.lib.brq_predict_ra_mutations(self.h, _b(gd_in), _b(fasta), int(bool(polymorphism_prediction)), int(bool(targeted_sequencing)),
                                                      int(bool(call_mutations_overlapping_missing_coverage)), _b(gd_out), n))
        return dict(zip(("SNP", "DEL", "INS", "SUB", "ra_marked_deleted"), list(n)))

    def hist_exchange_export(self):
        """64-byte handle of this context's inbox for the fused collective of pass 1 (exchange it with the peer ranks)."""
        h = C.create_string_buffer(64)
        cap = C.c_uint64()
        self._check(self.lib.brq_hist_exchange_export(self.h, h, C.byref(cap)))
        return h.raw

    def hist_exchange_attach(self, handles, rank):
        """``handles``: every rank's 64-byte handle in rank order.  From now on error_count() sums the ranks' histograms itself."""
        blob = b"".join(handles)
        assert len(blob) == 64 * len(handles)
        self._check(self.lib.brq_hist_exchange_attach(self.h, blob, len(handles), rank))

    def kernel_ms(self):
        a, b, c, d = C.c_float(), C.c_float(), C.c_float(), C.c_float()
        self.lib.brq_kernel_ms(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        e, f = C.c_float(), C.c_float()
        self.lib.brq_score_phase_ms(self.h, C.byref(e), C.byref(f))
        return {"hist": a.value, "coverage": b.value, "derive": c.value, "score": d.value, "tally": e.value, "fit": f.value}

    def event_record(self, slot):
        self._check(self.lib.brq_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_float()
        self._check(self.lib.brq_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return self.lib.brq_launch_count()


# ----------------------------------------------------------------------------------------------
# The reference's entry points for this path (same argument meaning; Settings fields passed flat)
# ----------------------------------------------------------------------------------------------
def error_count(bam, fasta, output_dir, readfiles, do_coverage=True, do_errors=True, preprocess_stage=False,
                min_qual_score=0, covariates="", *, call_mutations_seq_ids=None, read_file_sets=None,
                coverage_group_of_tid=None, error_rates_file_name=None, device=0, ctx=None,
                unmatched_end_minimum_read_length=50, require_match_fraction=0.9):
    """``breseq::error_count()`` (error_count.cpp:50-68): writes ``error_rates.tab``,
    ``base_qual_error_prob.<readfile>.tab`` and ``<group>.unique_only_coverage_distribution.tab``.

    ``min_qual_score`` is accepted and unused, as in the reference (error_count.h:295).
    With ``preprocess_stage`` (the stage 03 call, breseq_cmdline.cpp:1969) the return value is what the reference leaves in
    ``Summary::preprocess_error_count``: ``no_pos_hash_per_position_pr`` per BAM tid (error_count.cpp:217-229; the two
    ``Settings`` fields its read-end bound reads are keywords); otherwise None.
    """
    own = ctx is None
    ctx = ctx or Context(device)
    try:
        o, keep = _stage_options(seq_ids=call_mutations_seq_ids, read_file_sets=read_file_sets,
                                 coverage_groups=coverage_group_of_tid, use_base_repeat="base_repeat" in covariates,
                                 use_read_pos="read_pos" in covariates, preprocess_stage=preprocess_stage,
                                 unmatched_end_minimum_read_length=unmatched_end_minimum_read_length,
                                 require_match_fraction=require_match_fraction)
        arr = _str_array(list(readfiles))
        ctx._check(ctx.lib.brq_run_error_count(ctx.h, _b(bam), _b(fasta), _b(output_dir), _b(error_rates_file_name), arr,
                                               len(readfiles), int(do_coverage), int(do_errors), _b(covariates), C.byref(o)))
        if preprocess_stage:
            c = ctx.preprocess_read_starts().astype(np.float64)
            total = c.sum(axis=1)
            return [float(c[t, 0] / total[t]) if total[t] else 1.0 for t in range(len(c))]
        return None
    finally:
        if own:
            ctx.close()


def identify_mutations(bam, fasta, gd_file, deletion_propagation_cutoff, deletion_seed_cutoff, mutation_cutoff,
                       polymorphism_cutoff, polymorphism_precision_decimal, polymorphism_precision_places,
                       print_per_position_file=False, *, error_rates_file_name, base_quality_cutoff=3,
                       skip_missing_coverage_prediction=False, call_mutations_seq_ids=None, read_file_sets=None,
                       total_reference_length=0, device=0, ctx=None, per_position_file_name=None,
                       coverage_tsv_pattern=None, user_evidence_genome_diff_file_name=None, polymorphism_prediction=False):
    """``breseq::identify_mutations()`` (identify_mutations.cpp:48-88): writes ``ra_mc_evidence.gd`` and, like the
    reference, the per-position debug file when ``print_per_position_file`` (Settings::
    mutation_identification_per_position_file_name = ``per_position_file_name``) and ``<seq>.coverage.tsv`` when
    ``coverage_tsv_pattern`` is given (Settings::predict_copy_number / complete_coverage_text_file_name)."""
    if print_per_position_file and not per_position_file_name:
        raise BrqError("print_per_position_file needs per_position_file_name")
    own = ctx is None
    ctx = ctx or Context(device)
    try:
        o, keep = _stage_options(seq_ids=call_mutations_seq_ids, read_file_sets=read_file_sets, base_quality_cutoff=base_quality_cutoff,
                                 user_evidence_gd=user_evidence_genome_diff_file_name)
        n = len(deletion_propagation_cutoff)
        prop = (C.c_double * n)(*deletion_propagation_cutoff)
        seed = (C.c_double * n)(*deletion_seed_cutoff)
        p = Context.score_params(mutation_cutoff, polymorphism_cutoff, polymorphism_precision_decimal,
                                 polymorphism_precision_places, base_quality_cutoff, total_reference_length,
                                 polymorphism_prediction=polymorphism_prediction)
        ctx._check(ctx.lib.brq_run_identify_mutations(ctx.h, _b(bam), _b(fasta), _b(error_rates_file_name), _b(gd_file),
                                                      prop, seed, n, C.byref(p), int(skip_missing_coverage_prediction),
                                                      C.byref(o)))
        if print_per_position_file:
            ctx.write_per_position_file(per_position_file_name, list(deletion_propagation_cutoff))
        if coverage_tsv_pattern:
            ctx.write_coverage_tsv(coverage_tsv_pattern)
    finally:
        if own:
            ctx.close()
