// C ABI (include/brq.h) over the staging layer, the sm_100a kernels and the host finalisation.
#include "../../include/brq.h"

#include "bam_io.h"
#include "expand.h"
#include "coverage_fit.h"
#include "ra_filter.h"
#include "coverage_table.h"
#include "finalize.h"
#include "kernels.h"
#include "staging.h"
#include "synth.h"

#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <sys/stat.h>

using namespace brq;

namespace {

void* pinned_alloc(size_t bytes, bool* pinned) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 256, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    throw std::runtime_error("cudaHostAlloc failed for " + std::to_string(bytes) + " bytes");
  }
  *pinned = true;
  return p;
}
void pinned_release(void* p, bool pinned) { if (pinned) cudaFreeHost(p); else free(p); }

}  // namespace

struct brq_ctx {
  int device = -1;
  int threads = 8;
  std::string error;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {nullptr};
  cudaEvent_t user_ev[4] = {nullptr};
  float ms_hist = 0, ms_cov = 0, ms_derive = 0, ms_score = 0;

  BamHeader hdr;
  RefSet ref;
  ReadBatch reads;
  StageConfig stage_cfg;
  PileupStream st;
  bool staged = false, uploaded = false;

  StreamDev ds;                   // the staged stream in HBM (uploaded from the host's staging, or built there by the expander)
  ReadsDev d_reads;               // device staging: the aligned reads in HBM
  ExpandScratch xs;
  FlaggedRecordsHost flagged_records;  // device-built streams: the records of the flagged slots, for the host re-evaluation
  DevBuf<uint32_t> d_flagged, d_worklist, d_survivors, d_scalars;  // d_scalars: [0] err, [1] n_flagged, [2] n_work, [3] fit hand-out, [4] n_survivors, [5] screen hand-out
  DevBuf<uint8_t> d_score16;   // transfer form of the scoring stream (low halves)
  DevBuf<uint32_t> d_score_exc, d_score_exc_off;
  // both integer histograms in ONE allocation (the covariate counts, then the coverage histogram): a sharded run sums them
  // with one collective
  DevBuf<unsigned long long> d_hist;
  DevBuf<CoverageColumn> d_coverage_columns;   // brq_write_coverage_table
  // the fused collective of pass 1 (exchange.cu): this rank's inbox (two copies + counters, one allocation its peers map through
  // CUDA IPC), the peers' inboxes as mapped here, the step counter that picks the copy, and the kernel's completion counter
  void* exchange_inbox = nullptr;
  uint64_t exchange_capacity = 0;
  HistPeers peers{};
  std::vector<void*> peer_mappings;
  uint64_t exchange_step = 0;
  DevBuf<uint32_t> d_exchange_done;
  struct HistView { unsigned long long* p = nullptr; } d_counts, d_cov;
  DevBuf<double> d_log10;
  DevBuf<ClassTerms> d_lut;
  DevBuf<HotTerms> d_coldT;
  DevBuf<double> d_tallyT;
  DevBuf<HotRatios> d_hotR;
  DevBuf<double> d_prob;
  DevBuf<uint8_t> d_slot_mapq;
  uint8_t h_slot_mapq[256];
  TableGeometry geo;
  bool have_device_tables = false;
  float ms_tally = 0, ms_fit = 0;
  DevBuf<ColumnOut> d_cols, d_fcols;
  DevBuf<WalkOut> d_walk;
  DevBuf<WalkEvent> d_events;
  DevBuf<uint8_t> d_mark;
  DevBuf<uint32_t> d_seg_first, d_seg_last;
  DevBuf<double> d_seg_prop;
  DevBuf<uint64_t> d_ins_parent;
  std::vector<WalkEvent> h_events;  // the columns the host's interval walk has to look at
  std::vector<double> walk_prop;    // the propagation cutoffs h_events was compacted for
  std::vector<ColumnOut> h_fcols;   // full results of the flagged slots, in the order of h_flagged
  bool have_walk = false;
  std::unique_ptr<WorkerPool> pool; // parked host threads for short data-parallel jobs
  std::string shard_blob;           // brq_evidence_export
  // host copies of a device-built stream's arrays, made on demand (brq_stream, the per-position file, the coverage TSV)
  struct Mirror {
    bool valid = false;
    std::vector<uint32_t> score_rec, side_rec, side_off, round_slot, score_cnt;
    std::vector<uint64_t> score_off, hist_off, round_off;
    std::vector<uint8_t> slot_ref, hist_pos;
    std::vector<uint16_t> hist16;
    std::vector<uint32_t> hist_exc;
  } mirror;
  std::string reads_source;         // the BAM + FASTA the reads / reference came from, with size and mtime (load_inputs)
  uint64_t min_cov_depth = 0;       // brq_set_min_coverage_depth
  std::vector<void*> pinned_reads;  // page-locked arrays of `reads` (brq_pin_reads)
  bool host_hist_valid = false;     // h_counts / h_cov hold the counts of the current stream's last brq_error_count
  uint64_t d2h_bytes = 0;           // device -> host bytes since the last brq_d2h_bytes(reset) (bench bookkeeping)

  CovSpec spec;
  bool have_spec = false, have_table = false, have_counts = false, have_cols = false;
  uint64_t cov_stride = 0, n_groups = 0;
  std::vector<uint64_t> h_counts, h_cov;
  std::vector<double> h_log10, h_log10_text, h_prob;
  double* h_log10_pinned = nullptr;  // landing buffer of the device-derived log10 table
  size_t n_log10_pinned = 0;
  bool host_table_pending = false;   // the host copies above are stale until host_table_ready()
  bool derive_timed = true;
  bool hist_check_pending = false;   // error_count's kernels have not been checked / timed yet (finish_error_count)
  DevBuf<uint32_t> d_table_err;
  ClassLut h_lut;
  ScoreParams sp;
  brq_score_params last_params;
  std::vector<ColumnOut> h_cols;
  std::vector<uint32_t> h_flagged;
  uint32_t flagged_cap = 0;

  void need_device() const {
    if (device < 0) throw std::runtime_error("this context has no CUDA device (brq_config.device < 0): compute calls are unavailable");
  }
  void check_device_errors(const char* what) {
    uint32_t scal[2], table_err = 0;
    CUDA_OK(cudaMemcpyAsync(scal, d_scalars.p, 8, cudaMemcpyDeviceToHost, stream));
    if (d_table_err.p) CUDA_OK(cudaMemcpyAsync(&table_err, d_table_err.p, 4, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    if (table_err) {
      CUDA_OK(cudaMemsetAsync(d_table_err.p, 0, 4, stream));
      throw std::runtime_error(std::string(what) + ": an error-table value is outside the range of the device's six-digit text round trip (csrc/canonical.h)");
    }
    if (scal[0]) {
      std::string m = std::string(what) + ": ";
      if (scal[0] & BRQ_ERR_QUALITY_RANGE) m += "covariate 'quality' exceeded its maximum; ";
      if (scal[0] & BRQ_ERR_READSET_RANGE) m += "covariate 'read_set' exceeded its maximum; ";
      if (scal[0] & BRQ_ERR_READPOS_RANGE) m += "covariate 'read_pos' exceeded its maximum; ";
      if (scal[0] & BRQ_ERR_CLASS_OVERFLOW) m += "too many distinct record classes in one column; ";
      if (scal[0] & BRQ_ERR_DEPTH_RANGE) m += "column depth beyond the coverage histogram; ";
      if (scal[0] & BRQ_ERR_PEER_TIMEOUT) m += "a peer rank's histograms did not arrive (brq_hist_exchange_attach: every rank has to call brq_error_count); ";
      CUDA_OK(cudaMemsetAsync(d_scalars.p, 0, 4, stream));
      throw std::runtime_error(m);
    }
  }
};

namespace {

template <class F>
int guarded(brq_ctx* ctx, F&& f) {
  if (!ctx) return 1;
  // every call runs on the context's device whatever the calling thread's current device is, and leaves that as it was
  int prev_device = -1;
  const bool guard = ctx->device >= 0 && cudaGetDevice(&prev_device) == cudaSuccess && prev_device != ctx->device;
  if (guard) cudaSetDevice(ctx->device);
  struct Restore { bool on; int dev; ~Restore() { if (on) cudaSetDevice(dev); } } restore{guard, prev_device};
  try { f(); ctx->error.clear(); return 0; }
  catch (const std::exception& e) { ctx->error = e.what(); return 1; }
  catch (...) { ctx->error = "unknown error"; return 1; }
}

void apply_stage_options(brq_ctx* c, const brq_stage_options* o) {
  StageConfig& s = c->stage_cfg;
  s = StageConfig();
  s.threads = c->threads;
  if (c->device >= 0) { s.alloc = pinned_alloc; s.release = pinned_release; }
  if (!o) return;
  for (uint32_t i = 0; i < o->n_seq_ids; ++i) s.call_seq_ids.push_back(o->seq_ids[i]);
  for (uint32_t i = 0; i < o->n_read_file_sets; ++i) s.read_file_sets.push_back({o->read_file_sets[i].base_name, o->read_file_sets[i].n_files});
  if (o->coverage_group_of_tid) s.coverage_group_of_tid.assign(o->coverage_group_of_tid, o->coverage_group_of_tid + o->n_targets);
  s.use_base_repeat = o->use_base_repeat != 0;
  s.use_read_pos = o->use_read_pos != 0;
  // Settings::base_quality_cutoff: 0 is a value the reference accepts (every quality scores); BRQ_DEFAULT_BASE_QUALITY_CUTOFF = unset
  s.base_quality_cutoff = o->base_quality_cutoff == BRQ_DEFAULT_BASE_QUALITY_CUTOFF ? 3 : o->base_quality_cutoff;
  s.preprocess_stage = o->preprocess_stage != 0;
  s.unmatched_end_minimum_read_length = o->unmatched_end_minimum_read_length ? o->unmatched_end_minimum_read_length : 50;
  s.unmatched_end_length_factor = 1.0 - (o->require_match_fraction != 0.0 ? o->require_match_fraction : 0.9);
  s.shard_rank = o->shard_rank;
  s.shard_count = o->shard_count ? o->shard_count : 1;
  s.shard_lo = o->shard_lo; s.shard_hi = o->shard_hi; s.shard_explicit = o->shard_hi > o->shard_lo;
  s.staging_mode = (int)o->staging;
  if (o->user_evidence_gd && *o->user_evidence_gd) s.user_evidence = read_user_evidence_gd(o->user_evidence_gd);
}

void drop_stream(brq_ctx* c) {
  free_stream(c->st, c->stage_cfg);   // (a stream whose staging failed half-way still owns its buffers)
  c->staged = c->uploaded = false;
  c->have_counts = c->have_cols = c->have_walk = false;
  c->hist_check_pending = false;
  c->host_hist_valid = false;
  c->mirror = brq_ctx::Mirror();
}

// BAM records (c->reads) -> the streams.  On the device (the default when the context has one): the reads cross PCIe and
// expand.cu's kernels build the streams in HBM; BRQ_HOST_STAGING=1 or brq_stage_options.staging = 1 keeps the host's
// staging.cpp (the checker of the device path: tests compare the two streams array by array).
bool device_staging(const brq_ctx* c) {
  static const bool env_host = getenv("BRQ_HOST_STAGING") != nullptr;
  if (c->device < 0) return false;
  if (c->stage_cfg.staging_mode == 1) return false;
  if (c->stage_cfg.staging_mode == 2) return true;
  return !env_host;
}

void do_stage(brq_ctx* c) {
  if (c->stage_cfg.staging_mode == 2 && c->device < 0) throw std::runtime_error("device staging needs a context with a CUDA device");
  if (device_staging(c)) {
    static const bool times = getenv("BRQ_STAGE_TIMES") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    if (!c->d_reads.ring.parallel_for) {
      if (!c->pool) c->pool.reset(new WorkerPool((size_t)std::max(1, std::min(c->threads, 16) - 1)));
      WorkerPool* pool = c->pool.get();
      c->d_reads.ring.parallel_for = [pool](size_t n_parts, const std::function<void(size_t)>& body) {
        pool->run([&](size_t part, size_t n_workers) { for (size_t p = part; p < n_parts; p += n_workers) body(p); });
      };
    }
    c->d_reads.upload(c->reads, c->stream);
    if (times) { CUDA_OK(cudaStreamSynchronize(c->stream)); fprintf(stderr, "stage (device): upload %.1f ms (%.1f MB)\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), c->d_reads.bytes / 1e6); }
    const auto t1 = std::chrono::steady_clock::now();
    expand_on_device(c->hdr, c->ref, c->reads, c->d_reads, c->stage_cfg, c->xs, c->ds, c->st, c->stream);
    if (times) fprintf(stderr, "stage (device): expand %.1f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
    c->staged = true;
    c->uploaded = true;   // the stream was born in HBM
    return;
  }
  try {
    stage(c->hdr, c->ref, c->reads, c->stage_cfg, c->st);
  } catch (...) {
    free_stream(c->st, c->stage_cfg);
    throw;
  }
  c->staged = true;
  c->uploaded = false;
}

// host copies of a device-built stream (tests, optional outputs): a PileupStream whose pointers lead into c->mirror
PileupStream host_view(brq_ctx* c) {
  PileupStream v = c->st;
  if (!c->st.device_built) return v;
  brq_ctx::Mirror& m = c->mirror;
  const PileupStream& st = c->st;
  const uint64_t n_slots = st.n_slots();
  if (!m.valid) {
    auto down = [&](auto& vec, const auto* dev, size_t n) {
      vec.resize(n + 1);
      if (n) CUDA_OK(cudaMemcpyAsync(vec.data(), dev, n * sizeof(*dev), cudaMemcpyDeviceToHost, c->stream));
    };
    down(m.score_rec, c->ds.score_rec.p, st.n_score_padded); down(m.side_rec, c->ds.side_rec.p, st.n_side * st.geo.side_stride);
    down(m.side_off, c->ds.side_off.p, n_slots + 1); down(m.round_slot, c->ds.round_slot.p, st.n_rounds * 32);
    down(m.score_cnt, c->ds.score_cnt.p, n_slots); down(m.score_off, c->ds.score_off.p, n_slots + 1);
    down(m.hist_off, c->ds.hist_off.p, st.n_base + 1); down(m.round_off, c->ds.round_off.p, st.n_rounds + 1);
    down(m.slot_ref, c->ds.slot_ref.p, n_slots);
    if (st.hist_compact) {
      down(m.hist_pos, c->xs.hist_pos.p, st.n_hist * st.hist_bytes);
      down(m.hist16, reinterpret_cast<const uint16_t*>(c->ds.hist_rec.p), st.n_hist16);
      down(m.hist_exc, reinterpret_cast<const uint32_t*>(c->ds.hist_rec.p + c->ds.hist_exc_at), st.n_hist_exc);
    } else {
      down(m.hist_pos, c->ds.hist_rec.p, st.n_hist * st.hist_bytes);
    }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    m.valid = true;
  }
  v.score_rec = m.score_rec.data(); v.side_rec = m.side_rec.data(); v.side_off = m.side_off.data(); v.round_slot = m.round_slot.data();
  v.score_cnt = m.score_cnt.data(); v.score_off = m.score_off.data(); v.hist_off = m.hist_off.data(); v.round_off = m.round_off.data();
  v.slot_ref = m.slot_ref.data(); v.hist_rec = m.hist_pos.data();
  if (st.hist_compact) { v.hist16 = m.hist16.data(); v.hist_exc = m.hist_exc.data(); }
  return v;
}

SynthConfig synth_config(const brq_synth_spec* sp, int threads) {
  SynthConfig cfg;
  cfg.seed = sp->seed;
  cfg.threads = threads;
  for (uint32_t i = 0; i < sp->n_sets; ++i) {
    SynthReadSet s;
    s.name = sp->sets[i].name; s.paired = sp->sets[i].paired != 0; s.read_len = sp->sets[i].read_len;
    s.coverage = sp->sets[i].coverage; s.frag_mean = sp->sets[i].frag_mean; s.frag_sd = sp->sets[i].frag_sd;
    cfg.sets.push_back(s);
  }
  cfg.n_polymorphic = sp->n_polymorphic; cfg.n_fixed = sp->n_fixed; cfg.n_gaps = sp->n_gaps;
  if (sp->max_freq_ppm) { cfg.min_freq_ppm = sp->min_freq_ppm; cfg.max_freq_ppm = sp->max_freq_ppm; }
  cfg.window_lo = sp->window_lo; cfg.window_hi = sp->window_hi;
  return cfg;
}

void unpin_reads(brq_ctx* c) {
  for (void* p : c->pinned_reads) cudaHostUnregister(p);
  c->pinned_reads.clear();
}
void clear_inputs(brq_ctx* c) {
  unpin_reads(c);
  c->reads_source.clear();
  c->hdr = BamHeader(); c->ref = RefSet(); c->reads = ReadBatch();
}

// BAM + FASTA -> c->reads / c->hdr / c->ref, unless they are what the context already holds (the reference calls error_count()
// and identify_mutations() on the same reference.bam one after the other: the second call finds the reads decoded)
std::string source_key(const char* bam, const char* fasta) {
  std::string key;
  for (const char* path : {bam, fasta}) {
    struct stat sb;
    if (stat(path, &sb) != 0) return std::string();
    key += std::string(path) + "|" + std::to_string((long long)sb.st_size) + "|" + std::to_string((long long)sb.st_mtim.tv_sec) + "." + std::to_string((long long)sb.st_mtim.tv_nsec) + "|";
  }
  return key;
}
// shard: a configuration whose shard settings say which reference range this context stages (a rank of a sharded run): only
// the reads that overlap it are decoded, and a coordinate-sorted BAM is not read past it (bam_io.h: ReadSpans)
bool load_inputs(brq_ctx* c, const char* bam, const char* fasta, const StageConfig* shard = nullptr) {  // true when the inputs were (re)read
  std::string key = source_key(bam, fasta);
  const bool ranged = shard && (shard->shard_count > 1 || shard->shard_explicit || shard->shard_hi > shard->shard_lo);
  if (ranged && !key.empty()) {
    key += "shard " + std::to_string(shard->shard_rank) + "/" + std::to_string(shard->shard_count) + " " + std::to_string(shard->shard_lo) + "-" +
           std::to_string(shard->shard_hi) + (shard->shard_explicit ? "e" : "") + " of";
    for (const std::string& id : shard->call_seq_ids) key += " " + id;
  }
  if (!key.empty() && key == c->reads_source && !c->hdr.target_names.empty()) return false;
  drop_stream(c);
  clear_inputs(c);
  read_fasta(fasta, c->ref);
  if (ranged) {
    const ReadSpans spans = [&](const BamHeader& hdr) {
      PileupStream plan;
      std::vector<const std::string*> refseq;
      plan_segments(hdr, c->ref, *shard, plan, refseq);
      std::vector<std::pair<int32_t, int32_t>> by_tid(hdr.target_names.size(), std::make_pair(0, 0));
      for (const Segment& sg : plan.segments) by_tid[(size_t)sg.tid] = std::make_pair(sg.lo, sg.hi);
      return by_tid;
    };
    read_bam(bam, c->hdr, c->reads, c->threads, &spans);
  } else {
    read_bam(bam, c->hdr, c->reads, c->threads);
  }
  c->reads_source = key;
  return true;
}
bool same_stage_config(const StageConfig& a, const StageConfig& b) {
  auto sets = [](const StageConfig& x) { std::string t; for (const auto& s : x.read_file_sets) t += s.base_name + ":" + std::to_string(s.n_files) + ","; return t; };
  return a.call_seq_ids == b.call_seq_ids && a.coverage_group_of_tid == b.coverage_group_of_tid && sets(a) == sets(b) &&
         a.use_base_repeat == b.use_base_repeat && a.use_read_pos == b.use_read_pos && a.base_quality_cutoff == b.base_quality_cutoff &&
         a.want_hist == b.want_hist && a.want_score == b.want_score && a.preprocess_stage == b.preprocess_stage &&
         a.unmatched_end_minimum_read_length == b.unmatched_end_minimum_read_length && a.unmatched_end_length_factor == b.unmatched_end_length_factor &&
         a.shard_rank == b.shard_rank && a.shard_count == b.shard_count && a.shard_lo == b.shard_lo && a.shard_hi == b.shard_hi &&
         a.shard_explicit == b.shard_explicit && a.staging_mode == b.staging_mode && a.user_skip_cutoff == b.user_skip_cutoff &&
         a.user_evidence.size() == b.user_evidence.size() &&
         std::equal(a.user_evidence.begin(), a.user_evidence.end(), b.user_evidence.begin(), [](const UserRa& x, const UserRa& y) {
           return x.seq_id == y.seq_id && x.position == y.position && x.insert_position == y.insert_position && x.ref_base == y.ref_base && x.new_base == y.new_base; });
}

void synth_into(brq_ctx* c, const brq_synth_spec* sp) {
  clear_inputs(c);
  if (sp->contig_lens) {
    std::vector<uint32_t> lens(sp->contig_lens, sp->contig_lens + sp->n_contigs);
    synth_reference(sp->seed, lens, sp->contig_prefix ? sp->contig_prefix : "contig", c->ref);
  } else {
    read_fasta(sp->fasta, c->ref);
  }
  std::vector<SynthVariant> variants;
  synth_reads(synth_config(sp, c->threads), c->ref, c->hdr, c->reads, variants);
}

void upload(brq_ctx* c) {
  c->need_device();
  if (!c->staged) throw std::runtime_error("nothing staged");
  const PileupStream& st = c->st;
  if (st.device_built) { c->uploaded = true; return; }   // built in HBM by the expander
  const uint64_t n_slots = st.n_slots();
  c->ds.score_rec.ensure(st.n_score_padded + 64); c->ds.round_slot.ensure(st.n_rounds * 32 + 4); c->ds.score_off.ensure(n_slots + 1); c->ds.score_cnt.ensure(n_slots + 1); c->ds.round_off.ensure(st.n_rounds + 1); c->ds.round_side.ensure(st.n_rounds * 64 + 4); c->ds.slot_ref.ensure(n_slots);
  // the histogram records: the compact form when staging built it (16-bit fast records, then the 4-byte exceptions)
  c->ds.hist_exc_at = (st.n_hist16 * 2 + 31) & ~(size_t)15;
  c->ds.hist_rec.ensure(st.hist16 ? c->ds.hist_exc_at + st.n_hist_exc * 4 + 16 : st.n_hist * st.hist_bytes + 16);
  c->ds.side_rec.ensure(st.n_side * st.geo.side_stride + 4); c->ds.side_off.ensure(n_slots + 1); c->ds.hist_off.ensure(st.n_base + 1); c->ds.slot_group.ensure(st.n_base);
  if (st.score16) {  // the transfer form: low halves + exception words, expanded to score_rec on the device (end of this function)
    c->d_score16.ensure(st.n_score_padded * 2 + 64); c->d_score_exc.ensure(st.n_score_exc + 8); c->d_score_exc_off.ensure(st.n_rounds * 32 + 1);
    CUDA_OK(cudaMemcpyAsync(c->d_score16.p, st.score16, st.n_score_padded * 2, cudaMemcpyHostToDevice, c->stream));
    if (st.n_score_exc) CUDA_OK(cudaMemcpyAsync(c->d_score_exc.p, st.score_exc, st.n_score_exc * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->d_score_exc_off.p, st.score_exc_off, (st.n_rounds * 32 + 1) * 4, cudaMemcpyHostToDevice, c->stream));
  } else {
    CUDA_OK(cudaMemcpyAsync(c->ds.score_rec.p, st.score_rec, st.n_score_padded * 4, cudaMemcpyHostToDevice, c->stream));
  }
  CUDA_OK(cudaMemcpyAsync(c->ds.score_off.p, st.score_off, (n_slots + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->ds.score_cnt.p, st.score_cnt, n_slots * 4, cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->ds.round_off.p, st.round_off, (st.n_rounds + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->ds.round_side.p, st.round_side, st.n_rounds * 256, cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->ds.side_rec.p, st.side_rec, st.n_side * st.geo.side_stride * 4, cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->ds.side_off.p, st.side_off, (n_slots + 1) * 4, cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->ds.round_slot.p, st.round_slot, st.n_rounds * 128, cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->ds.slot_ref.p, st.slot_ref, n_slots, cudaMemcpyHostToDevice, c->stream));
  if (st.hist16) {
    CUDA_OK(cudaMemcpyAsync(c->ds.hist_rec.p, st.hist16, st.n_hist16 * 2, cudaMemcpyHostToDevice, c->stream));
    if (st.n_hist_exc) CUDA_OK(cudaMemcpyAsync(c->ds.hist_rec.p + c->ds.hist_exc_at, st.hist_exc, st.n_hist_exc * 4, cudaMemcpyHostToDevice, c->stream));
  } else {
    CUDA_OK(cudaMemcpyAsync(c->ds.hist_rec.p, st.hist_rec, st.n_hist * st.hist_bytes, cudaMemcpyHostToDevice, c->stream));
  }
  CUDA_OK(cudaMemcpyAsync(c->ds.hist_off.p, st.hist_off, (st.n_base + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->ds.slot_group.p, st.slot_group, st.n_base, cudaMemcpyHostToDevice, c->stream));
  if (st.score16)
    launch_expand_score(reinterpret_cast<const uint16_t*>(c->d_score16.p), c->ds.round_off.p, c->ds.round_slot.p, c->ds.slot_ref.p, c->d_score_exc.p,
                        c->d_score_exc_off.p, st.n_rounds, st.geo, c->ds.score_rec.p, c->stream);
  c->uploaded = true;
}

void error_count_device(brq_ctx* c, const std::string& covariates, bool do_coverage, bool do_errors) {
  c->need_device();
  static const bool call_times = getenv("BRQ_TIMING") != nullptr;
  const auto ct0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) { if (call_times) fprintf(stderr, "[brq] error_count: %s at %.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ct0).count()); };
  if (!c->uploaded) upload(c);
  c->spec = parse_covariates(covariates.empty() && !do_errors ? std::string("obs_base,ref_base,quality=1") : covariates);
  c->have_spec = true;
  const CovLayout lay = to_layout(c->spec);
  const PileupStream& st = c->st;
  // coverage histogram geometry: groups x (max depth + 1)
  const uint32_t n_groups = st.n_groups;
  c->cov_stride = std::max<uint64_t>(st.max_hist_depth, c->min_cov_depth) + 1;
  c->n_groups = n_groups;
  c->d_hist.ensure((size_t)lay.n_bins + c->cov_stride * n_groups);
  c->d_counts.p = c->d_hist.p;
  c->d_cov.p = c->d_hist.p + lay.n_bins;
  CUDA_OK(cudaMemsetAsync(c->d_hist.p, 0, ((size_t)lay.n_bins + c->cov_stride * n_groups) * 8, c->stream));
  CUDA_OK(cudaMemsetAsync(c->d_scalars.p, 0, 32, c->stream));
  CUDA_OK(cudaEventRecord(c->ev[0], c->stream));
  lap("memsets queued");
  if (do_errors) {
    // the reference ASSERTs when a counted covariate value exceeds the table (error_count.cpp:485-488); the stream's
    // maxima are known from staging, so the check costs the kernel nothing
    auto too_big = [&](int cov, const char* name, uint32_t seen) {
      if (st.n_hist && c->spec.used[cov] && !c->spec.clamp[cov] && seen >= c->spec.maxv[cov])
        throw std::runtime_error(std::string("Covariate '") + name + "' with value '" + std::to_string(seen) +
                                 "' exceeded enforced maximum value of '" + std::to_string(c->spec.maxv[cov] - 1) + "'.");
    };
    too_big(COV_QUALITY, "quality", st.max_hist_qual);
    too_big(COV_READ_SET, "read_set", st.max_read_set_seen);
    too_big(COV_READ_POS, "read_pos", st.max_hist_rpos);
    if (c->spec.used[COV_BASE_REPEAT] && c->spec.maxv[COV_BASE_REPEAT] > 32)
      throw std::runtime_error("base_repeat above 32 is not supported (the staged records keep five bits of it)");
    if (st.hist_bytes == 4 && (lay.off_rpos || lay.off_rep))
      throw std::runtime_error("the stream was staged without read_pos / base_repeat (brq_stage_options.use_read_pos, use_base_repeat)");
    if (st.hist_compact) launch_hist16(c->ds.hist_rec.p, st.n_hist16, reinterpret_cast<const uint32_t*>(c->ds.hist_rec.p + c->ds.hist_exc_at), st.n_hist_exc, lay, c->d_counts.p, c->stream);
    else launch_hist(c->ds.hist_rec.p, st.n_hist, st.hist_bytes == 8, lay, c->d_counts.p, c->stream);
  }
  lap("histogram kernel queued");
  CUDA_OK(cudaEventRecord(c->ev[1], c->stream));
  if (do_coverage) launch_coverage_hist(c->ds.hist_off.p, c->ds.slot_group.p, st.n_base, (uint32_t)c->cov_stride, n_groups, c->d_cov.p, c->d_scalars.p, c->stream);
  if (c->peers.world > 1) {
    // the ranks' histograms meet here, inside pass 1's own stream work (exchange.cu): no collective call by the caller
    const uint64_t words = (uint64_t)lay.n_bins + c->cov_stride * n_groups;
    if (words > c->exchange_capacity) throw std::runtime_error("the histograms are larger than the exchange inbox (brq_hist_exchange_export)");
    launch_hist_exchange(c->d_hist.p, words, c->peers, (uint32_t)(c->exchange_step++ & 1u), c->d_exchange_done.p, c->d_scalars.p, 5.0, c->stream);
  }
  CUDA_OK(cudaEventRecord(c->ev[2], c->stream));
  CUDA_OK(cudaGetLastError());
  lap("coverage kernel queued");
  // no synchronisation here: the kernels' error word and their event times are read by finish_error_count() when the
  // counts are downloaded or the timings asked for, and by the check that ends score_columns
  c->hist_check_pending = true;
  c->host_hist_valid = false;
  c->have_counts = true;
  c->have_table = false;
  c->host_table_pending = false;
}

void finish_error_count(brq_ctx* c) {
  if (!c->hist_check_pending) return;
  c->hist_check_pending = false;
  c->check_device_errors("error_count");
  CUDA_OK(cudaEventElapsedTime(&c->ms_hist, c->ev[0], c->ev[1]));
  CUDA_OK(cudaEventElapsedTime(&c->ms_cov, c->ev[1], c->ev[2]));
}

void download_hist(brq_ctx* c) {
  if (!c->have_counts) throw std::runtime_error("brq_error_count has not run");
  finish_error_count(c);
  c->h_counts.resize(c->spec.n_bins);
  c->h_cov.resize(c->cov_stride * c->n_groups);
  CUDA_OK(cudaMemcpyAsync(c->h_counts.data(), c->d_counts.p, c->h_counts.size() * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->h_cov.data(), c->d_cov.p, c->h_cov.size() * 8, cudaMemcpyDeviceToHost, c->stream));
  c->d2h_bytes += (c->h_counts.size() + c->h_cov.size()) * 8;
  CUDA_OK(cudaStreamSynchronize(c->stream));
  c->host_hist_valid = true;
}

// The error table reaches pass 2 through the reference's text round trip (six significant digits, error_count.cpp:629-690).
// Two ways in: derive_table() has the log10 table on the device and canonicalises it there (canonical.h: no host work
// inside a step; the host copies for the files and the re-evaluation of flagged slots are made when first asked for,
// host_table_ready()); a table loaded from a file or imported is canonicalised on the host and uploaded.
bool host_table_ready(brq_ctx* c) {
  if (!c->host_table_pending) return false;
  CUDA_OK(cudaStreamSynchronize(c->stream));
  c->h_log10.assign(c->h_log10_pinned, c->h_log10_pinned + c->n_log10_pinned);
  if (!c->pool) c->pool.reset(new WorkerPool((size_t)std::max(1, std::min(c->threads, 16) - 1)));
  canonicalise_table(c->h_log10, c->h_log10_text, c->h_prob, c->pool.get());
  c->host_table_pending = false;
  return true;
}

// every likelihood table of pass 2, built on the device from d_prob.  No synchronisation: the scoring kernels are
// stream-ordered behind the table build.
void build_device_tables(brq_ctx* c) {
  c->h_lut.clear();  // the host copy (re-evaluation of flagged slots) is rebuilt on demand
  c->have_device_tables = false;
  if (!c->staged) return;
  score_geometry(c->spec, c->st.mapq_seen, c->st.geo, c->sp, c->geo);
  if (c->device < 0) return;
  const TableGeometry& g = c->geo;
  c->d_lut.ensure(g.n_lut);
  c->d_coldT.ensure(g.n_cold);
  c->d_hotR.ensure(g.n_hotR);
  c->d_slot_mapq.ensure(g.mapqs.size());
  if (c->d_tallyT.n != g.n_tally_cells * 2 || !c->d_tallyT.p) {  // absent classes and the zero cells stay zero
    c->d_tallyT.ensure(g.n_tally_cells * 2);
    CUDA_OK(cudaMemsetAsync(c->d_tallyT.p, 0, g.n_tally_cells * 16, c->stream));
  }
  CUDA_OK(cudaMemsetAsync(c->d_coldT.p, 0, g.n_cold * sizeof(HotTerms), c->stream));
  uint8_t* slot_mapq = c->h_slot_mapq;
  for (size_t i = 0; i < g.mapqs.size(); ++i) slot_mapq[i] = (uint8_t)g.mapqs[i];
  CUDA_OK(cudaMemcpyAsync(c->d_slot_mapq.p, slot_mapq, g.mapqs.size(), cudaMemcpyHostToDevice, c->stream));
  TableBuildArgs a;
  a.prob = c->d_prob.p; a.slot_mapq = c->d_slot_mapq.p;
  a.n_st = g.n_st; a.n_mapq_slots = (uint32_t)g.mapqs.size(); a.Q = c->sp.max_qual;
  a.off_set = g.off_set; a.off_ref = g.off_ref; a.off_obs = g.off_obs; a.off_qual = g.off_qual; a.off_rpos = g.off_rpos; a.off_rep = g.off_rep;
  a.hot_slot = c->sp.mapq_slot[c->sp.hot_mapq];
  a.lut = c->d_lut.p; a.coldT = c->d_coldT.p; a.hotR = c->d_hotR.p; a.tallyT = c->d_tallyT.p;
  launch_build_tables(a, c->sp, c->stream);
  c->have_device_tables = true;
}

// h_log10 (loaded or imported) -> text-canonical probabilities on the host -> the device tables
void install_table(brq_ctx* c) {
  if (!host_table_ready(c)) {  // (a pending device-derived table becomes the host's first: re-install after a new staging call)
    if (!c->pool) c->pool.reset(new WorkerPool((size_t)std::max(1, std::min(c->threads, 16) - 1)));
    canonicalise_table(c->h_log10, c->h_log10_text, c->h_prob, c->pool.get());
  }
  c->have_table = true;
  if (c->device >= 0 && c->staged) {
    c->d_prob.ensure(c->h_prob.size());
    CUDA_OK(cudaMemcpyAsync(c->d_prob.p, c->h_prob.data(), c->h_prob.size() * 8, cudaMemcpyHostToDevice, c->stream));
  }
  build_device_tables(c);
}

void ensure_host_lut(brq_ctx* c) {
  if (c->h_lut.ready()) return;
  if (!c->have_table || !c->staged) throw std::runtime_error("no error table");
  host_table_ready(c);
  c->h_lut.reset(c->spec, c->h_prob, c->sp, c->geo);
}

void derive_table(brq_ctx* c) {
  c->need_device();
  if (!c->have_counts) throw std::runtime_error("brq_error_count has not run");
  if (!c->spec.used[COV_OBS_BASE]) throw std::runtime_error("the error table needs the obs_base covariate");
  const CovLayout lay = to_layout(c->spec);
  c->d_log10.ensure(lay.n_bins);
  CUDA_OK(cudaEventRecord(c->ev[3], c->stream));
  launch_derive_table(c->d_counts.p, lay, c->d_log10.p, c->stream);
  CUDA_OK(cudaEventRecord(c->ev[4], c->stream));
  // text round trip on the device, then the likelihood tables: nothing here waits for the host
  c->d_prob.ensure(lay.n_bins);
  c->d_table_err.ensure(1);
  CUDA_OK(cudaMemsetAsync(c->d_table_err.p, 0, 4, c->stream));
  launch_canonical_table(c->d_log10.p, lay.n_bins, c->d_prob.p, c->d_table_err.p, c->stream);
  if (c->n_log10_pinned < lay.n_bins) {
    if (c->h_log10_pinned) cudaFreeHost(c->h_log10_pinned);
    CUDA_OK(cudaHostAlloc((void**)&c->h_log10_pinned, (size_t)lay.n_bins * 8, cudaHostAllocDefault));
  }
  c->n_log10_pinned = lay.n_bins;
  CUDA_OK(cudaMemcpyAsync(c->h_log10_pinned, c->d_log10.p, (size_t)lay.n_bins * 8, cudaMemcpyDeviceToHost, c->stream));
  c->d2h_bytes += (uint64_t)lay.n_bins * 8;
  c->host_table_pending = true;   // h_log10 / h_prob: made by host_table_ready() when the files or flagged slots need them
  c->derive_timed = false;        // ev[3], ev[4] are read by kernel_ms()
  c->have_table = true;
  build_device_tables(c);
}

void score_device(brq_ctx* c, const brq_score_params* p) {
  c->need_device();
  if (!c->uploaded) upload(c);
  if (!c->have_table) throw std::runtime_error("no error table: call brq_derive_error_table or brq_load_error_table first");
  if (!c->have_device_tables) install_table(c);
  c->last_params = *p;
  uint64_t total = p->total_reference_length;
  if (!total) for (uint32_t l : c->hdr.target_lens) total += l;
  c->sp.log10_ref_length = log10((double)total);
  c->sp.mutation_cutoff = p->mutation_cutoff;
  c->sp.polymorphism_cutoff = p->polymorphism_cutoff;
  c->sp.precision_decimal = p->polymorphism_precision_decimal;
  c->sp.base_quality_cutoff = p->base_quality_cutoff;
  if (c->st.n_score && p->base_quality_cutoff != c->st.geo.cutoff)
    throw std::runtime_error("the stream was staged for base_quality_cutoff " + std::to_string(c->st.geo.cutoff) + ", not " +
                             std::to_string(p->base_quality_cutoff) + " (brq_stage_options.base_quality_cutoff)");
  c->sp.fit_all = (p->flags & BRQ_SCORE_FIT_ALL_COLUMNS) ? 1u : 0u;
  c->sp.keep_bounds = (p->flags & BRQ_SCORE_KEEP_BOUNDS) ? 1u : 0u;
  // the reference ASSERTs when a covariate value exceeds the table (error_count.cpp:485-488); the
  // stream's maxima are known from staging, so the check costs the kernels nothing
  if (c->st.n_score && c->st.max_qual_seen >= c->sp.max_qual)
    throw std::runtime_error("Covariate 'quality' with value '" + std::to_string(c->st.max_qual_seen) +
                             "' exceeded enforced maximum value of '" + std::to_string(c->sp.max_qual - 1) + "'.");
  if (c->st.n_score && c->st.max_read_set_seen >= c->sp.max_set)
    throw std::runtime_error("Covariate 'read_set' with value '" + std::to_string(c->st.max_read_set_seen) +
                             "' exceeded enforced maximum value of '" + std::to_string(c->sp.max_set - 1) + "'.");
  if (c->st.n_score && c->sp.n_rpos > 1 && c->st.max_score_rpos >= c->sp.n_rpos)
    throw std::runtime_error("Covariate 'read_pos' with value '" + std::to_string(c->st.max_score_rpos) +
                             "' exceeded enforced maximum value of '" + std::to_string(c->sp.n_rpos - 1) + "'.");
  const uint64_t n_slots = c->st.n_slots();
  c->d_cols.ensure(n_slots);
  c->flagged_cap = (uint32_t)std::min<uint64_t>(n_slots, 1u << 26);
  c->d_flagged.ensure(c->flagged_cap);
  CUDA_OK(cudaMemsetAsync(c->d_scalars.p + 1, 0, 28, c->stream));  // the error word [0] may still hold pass 1's verdict
  CUDA_OK(cudaEventRecord(c->ev[5], c->stream));
  c->d_worklist.ensure(n_slots);
  c->d_survivors.ensure(n_slots);
  launch_score_slots(c->ds.score_rec.p, c->ds.score_off.p, c->ds.score_cnt.p, c->ds.round_off.p, c->ds.side_rec.p, c->ds.side_off.p, reinterpret_cast<const uint2*>(c->ds.round_side.p), c->ds.slot_ref.p, c->ds.round_slot.p, c->st.n_rounds, n_slots, c->st.n_score, c->d_lut.p, c->d_tallyT.p, c->d_coldT.p, c->d_hotR.p, c->sp,
                     c->d_cols.p, c->d_worklist.p, c->d_survivors.p, c->d_flagged.p, c->d_scalars.p, c->flagged_cap, c->st.geo.side_stride, c->stream, c->ev[7]);
  CUDA_OK(cudaEventRecord(c->ev[6], c->stream));
  CUDA_OK(cudaGetLastError());
  c->check_device_errors("score_columns");
  if (getenv("BRQ_TIMING")) {
    uint32_t sc[8];
    CUDA_OK(cudaMemcpy(sc, c->d_scalars.p, sizeof sc, cudaMemcpyDeviceToHost));
    fprintf(stderr, "[brq] score: %u slots on the tally's work list, %u after the screen, %u flagged\n", sc[2], sc[4], sc[1]);
  }
  if (c->hist_check_pending) finish_error_count(c);  // (its error word was just checked; this reads the event times)
  CUDA_OK(cudaEventElapsedTime(&c->ms_score, c->ev[5], c->ev[6]));
  CUDA_OK(cudaEventElapsedTime(&c->ms_tally, c->ev[5], c->ev[7]));
  CUDA_OK(cudaEventElapsedTime(&c->ms_fit, c->ev[7], c->ev[6]));
  c->have_cols = true;
  c->have_walk = false;
  c->h_cols.clear();
}

void download_columns(brq_ctx* c) {
  if (!c->have_cols) throw std::runtime_error("brq_score_columns has not run");
  const uint64_t n_slots = c->st.n_slots();
  c->h_cols.resize(n_slots);
  uint32_t scal[2];
  CUDA_OK(cudaMemcpyAsync(scal, c->d_scalars.p, 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaMemcpyAsync(c->h_cols.data(), c->d_cols.p, n_slots * sizeof(ColumnOut), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  if (scal[1] > c->flagged_cap) throw std::runtime_error("flagged-slot list overflow");
  c->h_flagged.resize(scal[1]);
  if (scal[1]) CUDA_OK(cudaMemcpy(c->h_flagged.data(), c->d_flagged.p, (size_t)scal[1] * 4, cudaMemcpyDeviceToHost));
}

// What the evidence writer needs of the device results: the walk records of the event columns (compacted on the device:
// a few thousand of 4.6 M at C1), the flagged list, and the full 96-byte results of the flagged slots only.
void download_walk(brq_ctx* c, const double* prop, uint32_t n_targets) {
  if (!c->have_cols) throw std::runtime_error("brq_score_columns has not run");
  const PileupStream& st = c->st;
  const uint64_t n_slots = st.n_slots();
  uint32_t scal[2];
  CUDA_OK(cudaMemcpyAsync(scal, c->d_scalars.p, 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  if (scal[1] > c->flagged_cap) throw std::runtime_error("flagged-slot list overflow");
  uint32_t n_flagged = scal[1];
  if (!st.user_columns.empty()) {
    // user evidence: the slots the list meets are re-evaluated on the host whatever the kernels decided (their rows need the
    // full fit, identify_mutations.cpp:1914-2019): they join the flagged list here
    std::vector<uint32_t> have(n_flagged), extra;
    if (n_flagged) CUDA_OK(cudaMemcpy(have.data(), c->d_flagged.p, (size_t)n_flagged * 4, cudaMemcpyDeviceToHost));
    std::sort(have.begin(), have.end());
    for (const UserColumn& uc : st.user_columns) {
      if (uc.slot == ~0ull) continue;
      for (const auto& lv : uc.consumed) {
        uint64_t slot = uc.slot;
        if (lv.first > 0) {
          slot = ~0ull;
          const size_t j0 = std::lower_bound(st.ins_parent.begin(), st.ins_parent.end(), uc.slot) - st.ins_parent.begin();
          for (size_t j = j0; j < st.ins_parent.size() && st.ins_parent[j] == uc.slot; ++j) if (st.ins_count[j] == lv.first) slot = st.n_base + j;
          if (slot == ~0ull) continue;
        }
        if (!std::binary_search(have.begin(), have.end(), (uint32_t)slot) && std::find(extra.begin(), extra.end(), (uint32_t)slot) == extra.end()) extra.push_back((uint32_t)slot);
      }
    }
    if (!extra.empty()) {
      if (n_flagged + extra.size() > c->flagged_cap) throw std::runtime_error("flagged-slot list overflow");
      CUDA_OK(cudaMemcpy(c->d_flagged.p + n_flagged, extra.data(), extra.size() * 4, cudaMemcpyHostToDevice));
      n_flagged += (uint32_t)extra.size();
      CUDA_OK(cudaMemcpy(c->d_scalars.p + 1, &n_flagged, 4, cudaMemcpyHostToDevice));
    }
  }
  // segments of the visit order with their cutoffs
  std::vector<uint32_t> first, last;
  std::vector<double> sp;
  for (const Segment& sg : st.segments) {
    if (sg.hi <= sg.lo) continue;
    first.push_back((uint32_t)sg.slot0); last.push_back((uint32_t)(sg.slot0 + (uint64_t)(sg.hi - sg.lo) - 1));
    sp.push_back(prop[(size_t)sg.tid]);
  }
  c->h_events.clear();
  if (!first.empty() && st.n_base) {
    const uint32_t n_seg = (uint32_t)first.size();
    c->d_seg_first.ensure(n_seg); c->d_seg_last.ensure(n_seg); c->d_seg_prop.ensure(n_seg);
    c->d_mark.ensure(st.n_base); c->d_events.ensure(st.n_base); c->d_ins_parent.ensure(st.n_ins ? st.n_ins : 1);
    CUDA_OK(cudaMemcpyAsync(c->d_seg_first.p, first.data(), n_seg * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->d_seg_last.p, last.data(), n_seg * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->d_seg_prop.p, sp.data(), n_seg * 8, cudaMemcpyHostToDevice, c->stream));
    if (st.n_ins) CUDA_OK(cudaMemcpyAsync(c->d_ins_parent.p, st.ins_parent.data(), st.n_ins * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemsetAsync(c->d_scalars.p + 3, 0, 4, c->stream));  // scalars[3]: the fit kernel's hand-out counter, free again
    c->d_walk.ensure(st.n_base);
    launch_walk_events(c->d_cols.p, c->d_walk.p, st.n_base, c->d_seg_first.p, c->d_seg_last.p, c->d_seg_prop.p, n_seg, c->d_flagged.p,
                       c->d_scalars.p + 1, c->flagged_cap, c->d_ins_parent.p, c->d_mark.p, c->d_events.p, c->d_scalars.p + 3, c->stream);
    uint32_t n_events = 0;
    CUDA_OK(cudaMemcpyAsync(&n_events, c->d_scalars.p + 3, 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    c->h_events.resize(n_events);
    if (n_events) CUDA_OK(cudaMemcpyAsync(c->h_events.data(), c->d_events.p, (size_t)n_events * sizeof(WalkEvent), cudaMemcpyDeviceToHost, c->stream));
  }
  c->h_flagged.resize(n_flagged);
  c->h_fcols.resize(n_flagged);
  if (n_flagged) {
    c->d_fcols.ensure(n_flagged);
    launch_gather_columns(c->d_cols.p, c->d_flagged.p, n_flagged, c->d_fcols.p, c->stream);
    CUDA_OK(cudaMemcpyAsync(c->h_flagged.data(), c->d_flagged.p, (size_t)n_flagged * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->h_fcols.data(), c->d_fcols.p, (size_t)n_flagged * sizeof(ColumnOut), cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_OK(cudaStreamSynchronize(c->stream));
  (void)n_slots;
  if (st.device_built) gather_flagged_records(c->ds, st, c->d_flagged.p, n_flagged, c->flagged_records, c->stream, &c->d2h_bytes);
  c->d2h_bytes += 12 + c->h_events.size() * sizeof(WalkEvent) + (uint64_t)n_flagged * (4 + sizeof(ColumnOut));
  c->walk_prop.assign(prop, prop + n_targets);
  c->have_walk = true;
}

// where collect_evidence finds the flagged slots' records: in the host stream, or (device staging) in what download_walk gathered
const FlaggedRecords* flagged_view(brq_ctx* c, FlaggedRecords& fr) {
  if (!c->st.device_built) return nullptr;
  const FlaggedRecordsHost& h = c->flagged_records;
  fr.word_off = h.word_off.data(); fr.side_off = h.side_off.data(); fr.words = h.words.data(); fr.side = h.side.data(); fr.ref = h.ref.data();
  return &fr;
}

EvidenceCounts evidence(brq_ctx* c, const char* gd_file, const double* prop, const double* seed, uint32_t n_targets, int skip_mc) {
  const bool timing = getenv("BRQ_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto t0 = now();
  if (n_targets != c->hdr.target_names.size())
    throw std::runtime_error("Number of targets in BAM file [" + std::to_string(c->hdr.target_names.size()) +
                             "] does not match number in cutoff table [" + std::to_string(n_targets) + "].");
  if (!c->have_walk || c->walk_prop != std::vector<double>(prop, prop + n_targets)) download_walk(c, prop, n_targets);
  const auto t1 = now();
  EvidenceParams ep;
  ep.mutation_cutoff = c->last_params.mutation_cutoff;
  ep.polymorphism_cutoff = c->last_params.polymorphism_cutoff;
  ep.precision_decimal = c->last_params.polymorphism_precision_decimal;
  ep.precision_places = c->last_params.polymorphism_precision_places;
  ep.base_quality_cutoff = c->last_params.base_quality_cutoff;
  ep.log10_ref_length = c->sp.log10_ref_length;
  ep.skip_missing_coverage_prediction = skip_mc != 0;
  ep.polymorphism_prediction = (c->last_params.flags & BRQ_SCORE_POLYMORPHISM_PREDICTION) != 0;
  ep.deletion_propagation_cutoff.assign(prop, prop + n_targets);
  ep.deletion_seed_cutoff.assign(seed, seed + n_targets);
  ensure_host_lut(c);
  const auto t2 = now();
  FlaggedRecords fr;
  const EvidenceCounts k = write_evidence(gd_file, c->hdr, c->st, c->h_events, c->h_flagged, c->h_fcols, c->sp, c->h_lut, ep, flagged_view(c, fr));
  if (timing) fprintf(stderr, "[brq] evidence: download %.2f ms, lut %.2f ms, write_evidence %.2f ms (%zu flagged, %llu RA)\n",
                      ms(t0, t1), ms(t1, t2), ms(t2, now()), c->h_flagged.size(), (unsigned long long)k.ra);
  return k;
}

// This context's share of the evidence of a run sharded by reference range: event columns and RA rows, as bytes.
void evidence_export(brq_ctx* c, const double* prop, uint32_t n_targets) {
  if (n_targets != c->hdr.target_names.size())
    throw std::runtime_error("Number of targets in BAM file [" + std::to_string(c->hdr.target_names.size()) +
                             "] does not match number in cutoff table [" + std::to_string(n_targets) + "].");
  const bool timing = getenv("BRQ_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto t0 = now();
  if (!c->have_walk || c->walk_prop != std::vector<double>(prop, prop + n_targets)) download_walk(c, prop, n_targets);
  const auto t1 = now();
  EvidenceParams ep;
  ep.mutation_cutoff = c->last_params.mutation_cutoff;
  ep.polymorphism_cutoff = c->last_params.polymorphism_cutoff;
  ep.precision_decimal = c->last_params.polymorphism_precision_decimal;
  ep.precision_places = c->last_params.polymorphism_precision_places;
  ep.base_quality_cutoff = c->last_params.base_quality_cutoff;
  ep.log10_ref_length = c->sp.log10_ref_length;
  ep.skip_missing_coverage_prediction = false;
  ep.polymorphism_prediction = (c->last_params.flags & BRQ_SCORE_POLYMORPHISM_PREDICTION) != 0;
  ep.deletion_propagation_cutoff.assign(prop, prop + n_targets);
  ep.deletion_seed_cutoff.assign(n_targets, 0.0);
  ensure_host_lut(c);
  const auto t2 = now();
  FlaggedRecords fr;
  c->shard_blob = serialize_shard(collect_evidence(c->hdr, c->st, c->h_events, c->h_flagged, c->h_fcols, c->sp, c->h_lut, ep, flagged_view(c, fr)));
  if (timing) fprintf(stderr, "[brq] evidence_export: download %.2f ms, lut %.2f ms, collect %.2f ms (%zu flagged, %zu events, %zu record words, %zu side entries)\n",
                      ms(t0, t1), ms(t1, t2), ms(t2, now()), c->h_flagged.size(), c->h_events.size(), c->flagged_records.words.size(), c->flagged_records.side.size());
}

void write_pass1_files(brq_ctx* c, const char* output_dir, const char* error_rates_file, const char* const* readfiles,
                       uint32_t n_readfiles, int do_coverage, int do_errors, const char* counts_dump) {
  if (!c->host_hist_valid) download_hist(c);
  std::string dir = output_dir ? output_dir : ".";
  if (do_coverage) write_coverage_distributions(dir, c->h_cov, c->cov_stride, c->n_groups);
  if (do_errors && c->spec.per_position) {
    // covariates with ref_pos: the reference prints every position's counts during the pileup and writes no error rates
    // (error_count.cpp:105-111, 193-198, 255-275)
    const PileupStream view = host_view(c);
    write_count_table_per_position(dir + "/error_counts.tab", c->spec, view);
    return;
  }
  if (do_errors) {
    if (!c->have_table) throw std::runtime_error("brq_derive_error_table has not run");
    if (counts_dump && *counts_dump) write_count_table(counts_dump, c->spec, c->h_counts);
    host_table_ready(c);
    write_error_rates(error_rates_file && *error_rates_file ? error_rates_file : dir + "/error_rates.tab", c->spec, c->h_log10);
    std::vector<std::string> rf;
    for (uint32_t i = 0; i < n_readfiles; ++i) rf.push_back(readfiles[i]);
    if (c->spec.used[COV_READ_SET] && c->spec.used[COV_QUALITY] && c->spec.used[COV_REF_BASE] && c->spec.used[COV_OBS_BASE])
      write_base_qual_tables(dir + "/base_qual_error_prob.#.tab", c->spec, c->h_counts, rf);
  }
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

const char* brq_version(void) { return "breseq_b200 0.1 (sm_100a)"; }

brq_ctx* brq_create(const brq_config* cfg) {
  brq_ctx* c = new brq_ctx;
  c->device = cfg ? cfg->device : 0;
  c->threads = (cfg && cfg->threads > 0) ? cfg->threads : (int)std::max(1u, std::thread::hardware_concurrency());
  if (c->device >= 0) {
    try {
      CUDA_OK(cudaSetDevice(c->device));
      CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
      for (auto& e : c->ev) CUDA_OK(cudaEventCreate(&e));
      for (auto& e : c->user_ev) CUDA_OK(cudaEventCreate(&e));
      c->d_scalars.ensure(8);
      CUDA_OK(cudaMemset(c->d_scalars.p, 0, 32));
    } catch (const std::exception& e) {
      c->error = std::string("no usable CUDA device: ") + e.what();
      c->device = -2;  // poisoned: every compute call reports the error
    }
  }
  return c;
}

void brq_destroy(brq_ctx* c) {
  if (!c) return;
  if (c->device >= 0) cudaSetDevice(c->device);
  drop_stream(c);
  unpin_reads(c);
  if (c->device >= 0) {
    c->ds.release(); c->d_reads.release(); c->xs.release(); c->d_flagged.release(); c->d_worklist.release(); c->d_survivors.release(); c->d_tallyT.release(); c->d_coldT.release(); c->d_prob.release(); c->d_slot_mapq.release(); c->d_hotR.release(); c->d_scalars.release(); c->d_table_err.release(); c->d_score16.release(); c->d_score_exc.release(); c->d_score_exc_off.release();
    if (c->h_log10_pinned) { cudaFreeHost(c->h_log10_pinned); c->h_log10_pinned = nullptr; }
    c->d_hist.release(); c->d_coverage_columns.release();
    for (void* m : c->peer_mappings) cudaIpcCloseMemHandle(m);
    c->peer_mappings.clear();
    if (c->exchange_inbox) { cudaFree(c->exchange_inbox); c->exchange_inbox = nullptr; }
    c->d_exchange_done.release();
    c->d_log10.release(); c->d_lut.release(); c->d_cols.release(); c->d_fcols.release(); c->d_walk.release();
    c->d_events.release(); c->d_mark.release(); c->d_seg_first.release(); c->d_seg_last.release(); c->d_seg_prop.release(); c->d_ins_parent.release();
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    for (auto& e : c->user_ev) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
  }
  delete c;
}

const char* brq_last_error(const brq_ctx* c) { return c ? c->error.c_str() : "null context"; }

int brq_stage_bam(brq_ctx* c, const char* bam, const char* fasta, const brq_stage_options* opt) {
  return guarded(c, [&] {
    // (the options first: a sharded run's rank reads only its own range of the BAM)
    StageConfig wanted = c->stage_cfg;
    { const StageConfig kept = c->stage_cfg; apply_stage_options(c, opt); wanted = c->stage_cfg; c->stage_cfg = kept; }
    load_inputs(c, bam, fasta, &wanted);
    drop_stream(c);
    apply_stage_options(c, opt);
    do_stage(c);
  });
}

int brq_synth_write(brq_ctx* c, const brq_synth_spec* spec, const char* bam_out, const char* fasta_out) {
  return guarded(c, [&] {
    drop_stream(c);
    synth_into(c, spec);
    write_bam(bam_out, c->hdr, c->reads);
    write_fasta(fasta_out, c->ref);
  });
}

int brq_stage_synthetic(brq_ctx* c, const brq_synth_spec* spec, const brq_stage_options* opt) {
  return guarded(c, [&] {
    drop_stream(c);
    synth_into(c, spec);
    apply_stage_options(c, opt);
    do_stage(c);
  });
}

static void fill_stream_info(brq_ctx* c, const PileupStream& st, brq_stream_info* info, bool views) {
  memset(info, 0, sizeof *info);
  info->n_base = st.n_base; info->n_ins = st.n_ins; info->n_score_records = st.n_score; info->n_hist_records = st.n_hist;
  info->n_reads = c->reads.size();
  info->n_score_padded = st.n_score_padded;
  info->n_side = st.n_side; info->n_rounds = st.n_rounds;
  info->base_quality_cutoff = st.geo.cutoff; info->hot_mapq = st.geo.hot_mapq; info->table_q_lo = st.geo.q_lo; info->table_n_q = st.geo.n_q;
  info->table_n_st = st.geo.n_st; info->table_words = st.geo.n_words(); info->side_stride = st.geo.side_stride;
  info->device_built = st.device_built ? 1u : 0u;
  info->bytes_host = st.device_built ? st.bytes_uploaded : st.n_rounds * 392 + st.n_slots() * 4 + st.n_side * 4 * st.geo.side_stride + (st.n_slots() + 1) * 4 + (st.score16 ? st.n_score_padded * 2 + st.n_score_exc * 4 + (st.n_rounds * 32 + 1) * 4 : st.n_score_padded * 4) + (st.n_slots() + 1) * 8 + st.n_slots() + (st.hist16 ? st.n_hist16 * 2 + st.n_hist_exc * 4 : st.n_hist * st.hist_bytes) + (st.n_base + 1) * 8 + st.n_base;
  info->n_targets = (uint32_t)c->hdr.target_names.size(); info->pinned = st.pinned;
  info->hist_record_bytes = st.hist_bytes;
  info->n_hist16 = st.n_hist16; info->n_hist_exc = st.n_hist_exc; info->n_score_exc = st.n_score_exc;
  info->hist_compact = st.hist_compact ? 1u : 0u;
  if (!views) return;
  info->side_rec = st.side_rec; info->side_off = st.side_off;
  info->round_slot = st.round_slot; info->score_cnt = st.score_cnt; info->round_off = st.round_off;
  info->score_rec = st.score_rec; info->score_off = st.score_off; info->hist_rec = st.hist_rec; info->hist_off = st.hist_off;
  info->slot_ref = st.slot_ref; info->ins_parent = c->st.ins_parent.data(); info->ins_count = c->st.ins_count.data();
  info->hist16 = st.hist16; info->hist_exc = st.hist_exc;
  info->score16 = st.score16; info->score_exc = st.score_exc; info->score_exc_off = st.score_exc_off;
}

int brq_stream(brq_ctx* c, brq_stream_info* info) {
  return guarded(c, [&] {
    if (!c->staged) throw std::runtime_error("nothing staged");
    const PileupStream view = host_view(c);   // (a device-built stream is copied to the host here, once)
    fill_stream_info(c, view, info, true);
  });
}

int brq_stream_summary(brq_ctx* c, brq_stream_info* info) {
  return guarded(c, [&] {
    if (!c->staged) throw std::runtime_error("nothing staged");
    fill_stream_info(c, c->st, info, false);
  });
}

int brq_pin_reads(brq_ctx* c) {
  return guarded(c, [&] {
    c->need_device();
    if (!c->pinned_reads.empty()) return;
    static const bool verbose = getenv("BRQ_STAGE_TIMES") != nullptr;
    auto pin = [&](auto& vec) {
      if (vec.empty()) return;
      // in pieces: a registration the system refuses (locked-memory limits) leaves only its own piece pageable
      const size_t bytes = vec.size() * sizeof(vec[0]), piece = (size_t)1 << 30;
      char* base = reinterpret_cast<char*>(vec.data());
      for (size_t at = 0; at < bytes; at += piece) {
        const size_t len = std::min(piece, bytes - at);
        const cudaError_t e = cudaHostRegister(base + at, len, cudaHostRegisterDefault);
        if (e != cudaSuccess) { cudaGetLastError(); if (verbose) fprintf(stderr, "pin_reads: %zu bytes stay pageable (%s)\n", len, cudaGetErrorString(e)); continue; }
        c->pinned_reads.push_back(base + at);
      }
    };
    ReadBatch& R = c->reads;
    pin(R.tid); pin(R.pos); pin(R.flag); pin(R.mapq); pin(R.rg); pin(R.x1); pin(R.xl); pin(R.xr); pin(R.l_seq); pin(R.seq_off);
    pin(R.n_cigar); pin(R.cigar_off); pin(R.bases); pin(R.quals); pin(R.cigars);
  });
}

int brq_restage(brq_ctx* c) {
  return guarded(c, [&] {
    if (c->hdr.target_names.empty()) throw std::runtime_error("no reads to stage: call brq_stage_bam or brq_stage_synthetic first");
    drop_stream(c);
    do_stage(c);
  });
}

int brq_set_min_coverage_depth(brq_ctx* c, uint64_t depth) {
  if (!c) return 1;
  c->min_cov_depth = depth;
  return 0;
}

int brq_max_coverage_depth(brq_ctx* c, uint64_t* depth) {
  return guarded(c, [&] { if (!c->staged) throw std::runtime_error("nothing staged"); *depth = c->st.max_hist_depth; });
}

int brq_synth_shard_bounds(brq_ctx* c, const brq_synth_spec* sp, uint32_t n_shards, uint64_t* bounds) {
  return guarded(c, [&] {
    RefSet ref;
    if (sp->contig_lens) {
      std::vector<uint32_t> lens(sp->contig_lens, sp->contig_lens + sp->n_contigs);
      synth_reference(sp->seed, lens, sp->contig_prefix ? sp->contig_prefix : "contig", ref);
    } else {
      read_fasta(sp->fasta, ref);
    }
    const std::vector<uint64_t> b = synth_shard_bounds(synth_config(sp, c->threads), ref, n_shards);
    std::copy(b.begin(), b.end(), bounds);
  });
}

int brq_bam_shard_bounds(brq_ctx* c, const char* bam, uint32_t n_shards, uint64_t* bounds) {
  return guarded(c, [&] {
    BamHeader hdr; ReadBatch R;
    read_bam(bam, hdr, R, c->threads);
    // visit order = alphabetical target names (pileup_base.cpp:364-385); weights = aligned query bases by start column
    std::vector<size_t> order(hdr.target_names.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return hdr.target_names[a] < hdr.target_names[b]; });
    std::vector<uint64_t> slot0(order.size(), 0);
    uint64_t total = 0;
    for (size_t t : order) { slot0[t] = total; total += hdr.target_lens[t]; }
    const uint64_t bin = 256, n_bins = (total + bin - 1) / bin;
    std::vector<uint64_t> w(n_bins + 1, 0);
    uint64_t all = 0;
    for (size_t i = 0; i < R.size(); ++i) {
      if (R.tid[i] < 0 || (R.flag[i] & 4) || (size_t)R.tid[i] >= slot0.size()) continue;
      w[(slot0[(size_t)R.tid[i]] + (uint64_t)std::max(0, R.pos[i])) / bin] += R.l_seq[i]; all += R.l_seq[i];
    }
    bounds[0] = 0; bounds[n_shards] = total;
    uint64_t acc = 0, b = 0;
    for (uint32_t k = 1; k < n_shards; ++k) {
      const uint64_t want = all / n_shards * k;
      while (b < n_bins && acc + w[b] <= want) acc += w[b++];
      bounds[k] = std::max(bounds[k - 1], std::min(total, b * bin));
    }
  });
}

int brq_upload(brq_ctx* c) { return guarded(c, [&] { upload(c); }); }
int brq_sync(brq_ctx* c) {
  return guarded(c, [&] {
    c->need_device();
    if (c->hist_check_pending) finish_error_count(c);  // synchronises, and reports what pass 1's kernels flagged
    else CUDA_OK(cudaStreamSynchronize(c->stream));
  });
}

int brq_error_count(brq_ctx* c, const char* covariates, int do_coverage, int do_errors) {
  return guarded(c, [&] { error_count_device(c, covariates ? covariates : "", do_coverage != 0, do_errors != 0); });
}

int brq_preprocess_read_starts(brq_ctx* c, const uint64_t** counts, uint32_t* n_targets) {
  return guarded(c, [&] {
    if (!c->staged) throw std::runtime_error("nothing staged");
    if (c->st.read_start_counts.empty()) throw std::runtime_error("the stream was staged without brq_stage_options.preprocess_stage");
    *counts = c->st.read_start_counts.data(); *n_targets = (uint32_t)(c->st.read_start_counts.size() / 2);
  });
}

int brq_hist_device(brq_ctx* c, void** counts, uint64_t* n_bins, void** coverage, uint64_t* n_coverage) {
  return guarded(c, [&] {
    if (!c->have_counts) throw std::runtime_error("brq_error_count has not run");
    *counts = c->d_counts.p; *n_bins = c->spec.n_bins; *coverage = c->d_cov.p; *n_coverage = c->cov_stride * c->n_groups;
  });
}

int brq_hist_download(brq_ctx* c, const uint64_t** counts, uint64_t* n_bins, const uint64_t** coverage, uint64_t* stride, uint64_t* n_groups) {
  return guarded(c, [&] {
    download_hist(c);
    *counts = c->h_counts.data(); *n_bins = c->h_counts.size(); *coverage = c->h_cov.data(); *stride = c->cov_stride; *n_groups = c->n_groups;
  });
}

int brq_derive_error_table(brq_ctx* c) { return guarded(c, [&] { derive_table(c); }); }

int brq_error_table(brq_ctx* c, const double** log10_prob, uint64_t* n_bins) {
  return guarded(c, [&] {
    if (!c->have_table) throw std::runtime_error("no error table");
    host_table_ready(c);
    *log10_prob = c->h_log10.data(); *n_bins = c->h_log10.size();
  });
}

int brq_write_error_count_files(brq_ctx* c, const char* output_dir, const char* error_rates_file, const char* const* readfiles,
                                uint32_t n_readfiles, int do_coverage, int do_errors, const char* counts_dump_file) {
  return guarded(c, [&] { write_pass1_files(c, output_dir, error_rates_file, readfiles, n_readfiles, do_coverage, do_errors, counts_dump_file); });
}

int brq_load_error_table(brq_ctx* c, const char* error_rates_file) {
  return guarded(c, [&] {
    c->host_table_pending = false;
    read_error_rates(error_rates_file, c->spec, c->h_log10);
    c->have_spec = true;
    install_table(c);
  });
}

int brq_score_columns(brq_ctx* c, const brq_score_params* p) { return guarded(c, [&] { score_device(c, p); }); }

int brq_columns_download(brq_ctx* c, const brq_column** columns, uint64_t* n_slots, const uint32_t** flagged, uint32_t* n_flagged) {
  static_assert(sizeof(brq_column) == sizeof(ColumnOut), "brq_column layout");
  return guarded(c, [&] {
    download_columns(c);
    *columns = reinterpret_cast<const brq_column*>(c->h_cols.data()); *n_slots = c->h_cols.size();
    *flagged = c->h_flagged.data(); *n_flagged = (uint32_t)c->h_flagged.size();
  });
}

int brq_columns_device(brq_ctx* c, void** columns, uint64_t* n_slots) {
  return guarded(c, [&] {
    if (!c->have_cols) throw std::runtime_error("brq_score_columns has not run");
    *columns = c->d_cols.p; *n_slots = c->st.n_slots();
  });
}

int brq_write_evidence(brq_ctx* c, const char* gd_file, const double* prop, const double* seed, uint32_t n_targets, int skip_mc,
                       uint64_t* n_ra, uint64_t* n_mc, uint64_t* n_un) {
  return guarded(c, [&] {
    EvidenceCounts k = evidence(c, gd_file, prop, seed, n_targets, skip_mc);
    if (n_ra) *n_ra = k.ra;
    if (n_mc) *n_mc = k.mc;
    if (n_un) *n_un = k.un;
  });
}

int brq_run_error_count(brq_ctx* c, const char* bam, const char* fasta, const char* output_dir, const char* error_rates_file,
                        const char* const* readfiles, uint32_t n_readfiles, int do_coverage, int do_errors, const char* covariates,
                        const brq_stage_options* opt) {
  int rc = brq_stage_bam(c, bam, fasta, opt);
  if (rc) return rc;
  return guarded(c, [&] {
    error_count_device(c, covariates ? covariates : "", do_coverage != 0, do_errors != 0);
    if (do_errors && !c->spec.per_position) derive_table(c);
    write_pass1_files(c, output_dir, error_rates_file, readfiles, n_readfiles, do_coverage, do_errors, nullptr);
  });
}

int brq_run_identify_mutations(brq_ctx* c, const char* bam, const char* fasta, const char* error_rates_file, const char* gd_file,
                               const double* prop, const double* seed, uint32_t n_targets, const brq_score_params* p, int skip_mc,
                               const brq_stage_options* opt) {
  return guarded(c, [&] {
    // the error table names its covariates: a table with read_pos / base_repeat needs them in the staged records
    CovSpec spec;
    std::vector<double> log10_prob;
    read_error_rates(error_rates_file, spec, log10_prob);
    const bool fresh = load_inputs(c, bam, fasta);
    const StageConfig before = c->stage_cfg;
    apply_stage_options(c, opt);
    c->stage_cfg.use_read_pos = c->stage_cfg.use_read_pos || spec.used[COV_READ_POS];
    c->stage_cfg.use_base_repeat = c->stage_cfg.use_base_repeat || spec.used[COV_BASE_REPEAT];
    if (p) c->stage_cfg.base_quality_cutoff = p->base_quality_cutoff;   // the score parameters carry Settings::base_quality_cutoff
    if (!c->stage_cfg.user_evidence.empty()) c->stage_cfg.user_skip_cutoff.assign(prop, prop + n_targets);  // skipped targets never look at the user list
    // the stream error_count() staged from the same BAM with the same options is still there: no second staging
    if (fresh || !c->staged || !same_stage_config(before, c->stage_cfg)) { drop_stream(c); do_stage(c); }
    c->host_table_pending = false;
    c->spec = spec; c->h_log10 = log10_prob;
    c->have_spec = true;
    install_table(c);
    score_device(c, p);
    evidence(c, gd_file, prop, seed, n_targets, skip_mc);
  });
}

int brq_write_per_position_file(brq_ctx* c, const char* path, const double* prop, uint32_t n_targets) {
  return guarded(c, [&] {
    if (n_targets != c->hdr.target_names.size()) throw std::runtime_error("the cutoff table does not match the BAM targets");
    if (c->h_cols.size() != c->st.n_slots()) download_columns(c);
    const PileupStream view = host_view(c);
    write_per_position_file(path, c->hdr, view, c->h_cols, c->last_params.base_quality_cutoff, std::vector<double>(prop, prop + n_targets));
  });
}

int brq_write_coverage_tsv(brq_ctx* c, const char* pattern) {
  return guarded(c, [&] {
    if (c->h_cols.size() != c->st.n_slots()) download_columns(c);
    const PileupStream view = host_view(c);
    // a BAM with two or more read groups gets the three columns once more per group (identify_mutations.cpp:858-862,
    // 1588-1610, 2046-2050): pass 2's coverage tallies split by read group, from one walk per group over the reads in HBM
    std::vector<std::vector<CoverageColumn>> by_group;
    if (c->hdr.read_groups.ids.size() > 1) {
      if (!c->st.device_built) throw std::runtime_error("the per-read-group columns of the coverage TSV need reads staged on the device");
      by_group.resize(c->hdr.read_groups.ids.size());
      for (size_t g = 0; g < by_group.size(); ++g) {
        coverage_columns_on_device(c->xs, c->st.n_base, c->d_coverage_columns, (uint32_t)g, true, c->stream);
        by_group[g].resize(c->st.n_base);
        if (c->st.n_base) CUDA_OK(cudaMemcpyAsync(by_group[g].data(), c->d_coverage_columns.p, c->st.n_base * sizeof(CoverageColumn), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        c->d2h_bytes += c->st.n_base * sizeof(CoverageColumn);
      }
    }
    write_coverage_tsv(pattern, c->hdr, c->ref, view, c->h_cols, by_group);
  });
}

int brq_write_per_position_counts(brq_ctx* c, const char* covariates, const char* path) {
  return guarded(c, [&] {
    if (!c->staged) throw std::runtime_error("nothing staged");
    const CovSpec spec = parse_covariates(covariates ? covariates : "");
    const PileupStream view = host_view(c);
    write_count_table_per_position(path, spec, view);
  });
}

static int coverage_table(brq_ctx* c, const char* region, const char* path, uint32_t resolution, int total_only, int csv, int per_read_group,
                          const double* reference_average) {
  return guarded(c, [&] {
    c->need_device();
    if (!c->staged || !c->st.device_built) throw std::runtime_error("the coverage table needs reads staged on the device (brq_stage_options.staging = 0 or 2)");
    const size_t n_base = c->st.n_base;
    auto walk = [&](uint32_t group, std::vector<CoverageColumn>& cols) {
      coverage_columns_on_device(c->xs, n_base, c->d_coverage_columns, group, false, c->stream);
      cols.resize(n_base);
      if (n_base) CUDA_OK(cudaMemcpyAsync(cols.data(), c->d_coverage_columns.p, n_base * sizeof(CoverageColumn), cudaMemcpyDeviceToHost, c->stream));
      CUDA_OK(cudaStreamSynchronize(c->stream));
      CUDA_OK(cudaGetLastError());
      c->d2h_bytes += n_base * sizeof(CoverageColumn);
    };
    std::vector<CoverageColumn> cols;
    walk(COVERAGE_ALL_GROUPS, cols);
    // one more walk per read group: at least one set, like bam2cov --per-read-group (coverage_output.h:142-143)
    std::vector<std::vector<CoverageColumn>> by_group(per_read_group ? std::max<size_t>(c->hdr.read_groups.ids.size(), 1) : 0);
    for (size_t g = 0; g < by_group.size(); ++g) walk((uint32_t)g, by_group[g]);
    write_coverage_table(path, c->hdr, c->ref, c->st, cols, by_group, region ? region : "", resolution, total_only != 0, csv != 0, reference_average);
  });
}

int brq_write_coverage_table(brq_ctx* c, const char* region, const char* path, uint32_t resolution, int total_only, int csv, int per_read_group) {
  return coverage_table(c, region, path, resolution, total_only, csv, per_read_group, nullptr);
}

int brq_write_coverage_table_with_average(brq_ctx* c, const char* region, const char* path, uint32_t resolution, int total_only, int csv,
                                          int per_read_group, double reference_unique_average_cov) {
  return coverage_table(c, region, path, resolution, total_only, csv, per_read_group, &reference_unique_average_cov);
}

int brq_evidence_export(brq_ctx* c, const double* prop, uint32_t n_targets, const void** data, uint64_t* bytes) {
  return guarded(c, [&] {
    evidence_export(c, prop, n_targets);
    *data = c->shard_blob.data(); *bytes = c->shard_blob.size();
  });
}

int brq_write_evidence_merged(brq_ctx* c, const void* const* shards, const uint64_t* sizes, uint32_t n_shards, const char* gd_file,
                              const double* prop, const double* seed, uint32_t n_targets, int skip_mc,
                              uint64_t* n_ra, uint64_t* n_mc, uint64_t* n_un) {
  return guarded(c, [&] {
    std::vector<EvidenceShard> parsed;
    for (uint32_t i = 0; i < n_shards; ++i) parsed.push_back(parse_shard(shards[i], sizes[i]));
    std::vector<const EvidenceShard*> ptrs;
    for (const EvidenceShard& sh : parsed) ptrs.push_back(&sh);
    EvidenceParams ep;
    ep.mutation_cutoff = ep.polymorphism_cutoff = ep.precision_decimal = 0.0;
    ep.precision_places = 0; ep.base_quality_cutoff = 0; ep.log10_ref_length = 0.0;
    ep.skip_missing_coverage_prediction = skip_mc != 0;
    ep.deletion_propagation_cutoff.assign(prop, prop + n_targets);
    ep.deletion_seed_cutoff.assign(seed, seed + n_targets);
    const EvidenceCounts k = walk_evidence(ptrs, ep, gd_file);
    if (n_ra) *n_ra = k.ra;
    if (n_mc) *n_mc = k.mc;
    if (n_un) *n_un = k.un;
  });
}

namespace {
void fill_fit(const CoverageFit& f, brq_coverage_fit* out) {
  out->average = f.average; out->variance = f.variance; out->relative_variance = f.relative_variance;
  out->nbinom_size_parameter = f.nb_size; out->nbinom_mean_parameter = f.nb_mu;
  out->deletion_coverage_propagation_cutoff = f.deletion_coverage_propagation_cutoff;
  out->censor_start = f.censor_start; out->censor_end = f.censor_end;
}
// the context's parked threads take the fit's independent restarts
std::function<void(size_t, const std::function<void(size_t)>&)> fit_threads(brq_ctx* c) {
  if (!c->pool) c->pool.reset(new WorkerPool((size_t)std::max(1, std::min(c->threads, 16) - 1)));
  WorkerPool* pool = c->pool.get();
  return [pool](size_t n_jobs, const std::function<void(size_t)>& job) {
    pool->run([&](size_t part, size_t n_workers) { for (size_t j = part; j < n_jobs; j += n_workers) job(j); });
  };
}
}  // namespace

int brq_fit_coverage_distribution(brq_ctx* c, uint32_t group, double pr_cutoff, brq_coverage_fit* out) {
  return guarded(c, [&] {
    if (!c->have_counts && !c->host_hist_valid) throw std::runtime_error("brq_error_count has not run");
    if (!c->host_hist_valid) download_hist(c);
    if (group >= c->n_groups) throw std::runtime_error("no such coverage group");
    // the histogram as its file would be read back (error_count.cpp:239-253, coverage_distribution.cpp:34-65): depths 1 .. the deepest seen
    const uint64_t* h = c->h_cov.data() + (size_t)group * c->cov_stride;
    uint32_t N = 0;
    for (uint64_t j = 0; j < c->cov_stride; ++j) if (h[j]) N = (uint32_t)j;
    std::vector<double> n((size_t)N + 1, 0.0);
    for (uint32_t j = 1; j <= N; ++j) n[j] = (double)h[j];
    const auto threads = fit_threads(c);
    fill_fit(fit_coverage_distribution(n, N, pr_cutoff, &threads), out);
  });
}

int brq_fit_coverage_file(brq_ctx* c, const char* path, double pr_cutoff, brq_coverage_fit* out) {
  return guarded(c, [&] {
    std::vector<double> n;
    uint32_t N = 0;
    read_coverage_distribution(path, n, N);
    const auto threads = fit_threads(c);
    fill_fit(fit_coverage_distribution(n, N, pr_cutoff, &threads), out);
  });
}

// ---- the Output stage's RA filter (ra_filter.cpp)
void brq_ra_filter_defaults(int polymorphism_prediction, brq_ra_filter_options* out) {
  const RaFilterOptions o = ra_filter_defaults(polymorphism_prediction != 0);
  out->polymorphism_prediction = o.polymorphism_prediction;
  out->mutation_log10_e_value_cutoff = o.mutation_log10_e_value_cutoff;
  out->consensus_frequency_cutoff = o.consensus_frequency_cutoff;
  out->consensus_minimum_variant_coverage = o.consensus_minimum_variant_coverage;
  out->consensus_minimum_total_coverage = o.consensus_minimum_total_coverage;
  out->consensus_minimum_variant_coverage_each_strand = o.consensus_minimum_variant_coverage_each_strand;
  out->consensus_minimum_total_coverage_each_strand = o.consensus_minimum_total_coverage_each_strand;
  out->consensus_reject_indel_homopolymer_length = o.consensus_reject_indel_homopolymer_length;
  out->consensus_reject_surrounding_homopolymer_length = o.consensus_reject_surrounding_homopolymer_length;
  out->polymorphism_log10_e_value_cutoff = o.polymorphism_log10_e_value_cutoff;
  out->polymorphism_frequency_cutoff = o.polymorphism_frequency_cutoff;
  out->polymorphism_minimum_variant_coverage = o.polymorphism_minimum_variant_coverage;
  out->polymorphism_minimum_total_coverage = o.polymorphism_minimum_total_coverage;
  out->polymorphism_minimum_variant_coverage_each_strand = o.polymorphism_minimum_variant_coverage_each_strand;
  out->polymorphism_minimum_total_coverage_each_strand = o.polymorphism_minimum_total_coverage_each_strand;
  out->polymorphism_reject_indel_homopolymer_length = o.polymorphism_reject_indel_homopolymer_length;
  out->polymorphism_reject_surrounding_homopolymer_length = o.polymorphism_reject_surrounding_homopolymer_length;
  out->polymorphism_fisher_strand_p_value_cutoff = o.polymorphism_fisher_strand_p_value_cutoff;
  out->polymorphism_ks_quality_p_value_cutoff = o.polymorphism_ks_quality_p_value_cutoff;
  out->polymorphism_no_indels = o.polymorphism_no_indels;
}

void brq_binomial_frequency_bounds(double k, double n, double alpha, double* lower, double* upper) {
  *lower = binomial_frequency_lower_bound(k, n, alpha);
  *upper = binomial_frequency_upper_bound(k, n, alpha);
}

double brq_fisher_strand_p_value(uint32_t minor_top, uint32_t minor_bottom, uint32_t major_top, uint32_t major_bottom) {
  return fisher_strand_p_value(minor_top, minor_bottom, major_top, major_bottom);
}

int brq_test_ra_evidence(brq_ctx* c, const char* gd_in, const char* fasta, const brq_ra_filter_options* in, const char* gd_out,
                         uint32_t* counts5) {
  return guarded(c, [&] {
    RaFilterOptions o;
    o.polymorphism_prediction = in->polymorphism_prediction != 0;
    o.mutation_log10_e_value_cutoff = in->mutation_log10_e_value_cutoff;
    o.consensus_frequency_cutoff = in->consensus_frequency_cutoff;
    o.consensus_minimum_variant_coverage = in->consensus_minimum_variant_coverage;
    o.consensus_minimum_total_coverage = in->consensus_minimum_total_coverage;
    o.consensus_minimum_variant_coverage_each_strand = in->consensus_minimum_variant_coverage_each_strand;
    o.consensus_minimum_total_coverage_each_strand = in->consensus_minimum_total_coverage_each_strand;
    o.consensus_reject_indel_homopolymer_length = in->consensus_reject_indel_homopolymer_length;
    o.consensus_reject_surrounding_homopolymer_length = in->consensus_reject_surrounding_homopolymer_length;
    o.polymorphism_log10_e_value_cutoff = in->polymorphism_log10_e_value_cutoff;
    o.polymorphism_frequency_cutoff = in->polymorphism_frequency_cutoff;
    o.polymorphism_minimum_variant_coverage = in->polymorphism_minimum_variant_coverage;
    o.polymorphism_minimum_total_coverage = in->polymorphism_minimum_total_coverage;
    o.polymorphism_minimum_variant_coverage_each_strand = in->polymorphism_minimum_variant_coverage_each_strand;
    o.polymorphism_minimum_total_coverage_each_strand = in->polymorphism_minimum_total_coverage_each_strand;
    o.polymorphism_reject_indel_homopolymer_length = in->polymorphism_reject_indel_homopolymer_length;
    o.polymorphism_reject_surrounding_homopolymer_length = in->polymorphism_reject_surrounding_homopolymer_length;
    o.polymorphism_fisher_strand_p_value_cutoff = in->polymorphism_fisher_strand_p_value_cutoff;
    o.polymorphism_ks_quality_p_value_cutoff = in->polymorphism_ks_quality_p_value_cutoff;
    o.polymorphism_no_indels = in->polymorphism_no_indels != 0;
    RefSet ref;
    read_fasta(fasta, ref);
    normalise_reference(ref);
    const RaFilterCounts n = test_ra_evidence(gd_in, ref, o, gd_out);
    if (counts5) {
      counts5[0] = n.rows; counts5[1] = n.consensus; counts5[2] = n.polymorphism; counts5[3] = n.rejected_kept; counts5[4] = n.deleted;
    }
  });
}

int brq_predict_ra_mutations(brq_ctx* c, const char* gd_in, const char* fasta, int polymorphism_prediction, int targeted_sequencing,
                             int call_mutations_overlapping_missing_coverage, const char* gd_out, uint32_t* counts5) {
  return guarded(c, [&] {
    RefSet ref;
    read_fasta(fasta, ref);
    normalise_reference(ref);
    const RaMutationCounts n = predict_ra_mutations(gd_in, ref, polymorphism_prediction != 0, targeted_sequencing != 0,
                                                    call_mutations_overlapping_missing_coverage != 0, gd_out);
    if (counts5) { counts5[0] = n.snp; counts5[1] = n.del; counts5[2] = n.ins; counts5[3] = n.sub; counts5[4] = n.ra_marked_deleted; }
  });
}

// ---- the fused collective of pass 1 (exchange.cu)
int brq_hist_exchange_export(brq_ctx* c, void* handle64, uint64_t* capacity_words) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  return guarded(c, [&] {
    c->need_device();
    if (!c->exchange_inbox) {
      c->exchange_capacity = (uint64_t)1 << 20;   // words per copy: covariate bins + coverage groups x depths of any run so far
      const size_t bytes = 2 * c->exchange_capacity * 8 + 64;
      CUDA_OK(cudaMalloc(&c->exchange_inbox, bytes));
      CUDA_OK(cudaMemset(c->exchange_inbox, 0, bytes));
      c->d_exchange_done.ensure(1);
      CUDA_OK(cudaMemset(c->d_exchange_done.p, 0, 4));
    }
    cudaIpcMemHandle_t h;
    CUDA_OK(cudaIpcGetMemHandle(&h, c->exchange_inbox));
    memcpy(handle64, &h, 64);
    if (capacity_words) *capacity_words = c->exchange_capacity;
  });
}

int brq_hist_exchange_attach(brq_ctx* c, const void* handles, uint32_t world, uint32_t rank) {
  return guarded(c, [&] {
    c->need_device();
    if (!c->exchange_inbox) throw std::runtime_error("brq_hist_exchange_export has not run");
    if (world < 1 || world > 16 || rank >= world) throw std::runtime_error("brq_hist_exchange_attach: 1 to 16 ranks");
    for (void* m : c->peer_mappings) cudaIpcCloseMemHandle(m);
    c->peer_mappings.clear();
    HistPeers P{};
    P.world = world; P.rank = rank; P.capacity = c->exchange_capacity;
    for (uint32_t r = 0; r < world; ++r) {
      void* base = c->exchange_inbox;
      if (r != rank) {
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + (size_t)r * 64, 64);
        CUDA_OK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_mappings.push_back(base);
      }
      P.inbox[r] = static_cast<unsigned long long*>(base);
      P.arrived[r] = reinterpret_cast<uint32_t*>(static_cast<char*>(base) + 2 * c->exchange_capacity * 8);
    }
    c->peers = P;
    c->exchange_step = 0;
  });
}

int brq_cuda_stream(brq_ctx* c, void** stream) {
  return guarded(c, [&] { c->need_device(); *stream = (void*)c->stream; });
}

int brq_d2h_bytes(brq_ctx* c, uint64_t* bytes, int reset) {
  if (!c) return 1;
  if (bytes) *bytes = c->d2h_bytes;
  if (reset) c->d2h_bytes = 0;
  return 0;
}

int brq_launch_count(void) { return launch_count(); }

int brq_event_record(brq_ctx* c, int slot) {
  return guarded(c, [&] {
    c->need_device();
    if (slot < 0 || slot > 3) throw std::runtime_error("event slot out of range");
    CUDA_OK(cudaEventRecord(c->user_ev[slot], c->stream));
  });
}

int brq_event_elapsed_ms(brq_ctx* c, int a, int b, float* ms) {
  return guarded(c, [&] {
    c->need_device();
    if (a < 0 || a > 3 || b < 0 || b > 3) throw std::runtime_error("event slot out of range");
    CUDA_OK(cudaEventSynchronize(c->user_ev[b]));
    CUDA_OK(cudaEventElapsedTime(ms, c->user_ev[a], c->user_ev[b]));
  });
}

int brq_score_phase_ms(brq_ctx* c, float* tally_ms, float* fit_ms) {
  if (!c) return 1;
  if (tally_ms) *tally_ms = c->ms_tally;
  if (fit_ms) *fit_ms = c->ms_fit;
  return 0;
}

int brq_kernel_ms(brq_ctx* c, float* hist_ms, float* coverage_ms, float* derive_ms, float* score_ms) {
  if (!c) return 1;
  if (c->hist_check_pending) { const int rc = guarded(c, [&] { finish_error_count(c); }); if (rc) return rc; }
  if (hist_ms) *hist_ms = c->ms_hist;
  if (coverage_ms) *coverage_ms = c->ms_cov;
  if (!c->derive_timed && c->device >= 0) {  // derive_table() does not wait for its own kernels
    if (cudaEventSynchronize(c->ev[4]) == cudaSuccess && cudaEventElapsedTime(&c->ms_derive, c->ev[3], c->ev[4]) == cudaSuccess) c->derive_timed = true;
  }
  if (derive_ms) *derive_ms = c->ms_derive;
  if (score_ms) *score_ms = c->ms_score;
  return 0;
}

}  // extern "C"
