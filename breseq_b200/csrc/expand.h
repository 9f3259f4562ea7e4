// Device-side staging: the aligned reads in HBM -> the streams of brq_types.h, built by kernels (expand.cu, expand_core.h).
#pragma once
#include "bam_io.h"
#include "brq_types.h"
#include "devbuf.h"
#include "expand_core.h"
#include "staging.h"

#include <functional>

namespace brq {

// Pageable host memory -> HBM at PCIe speed: a ring of page-locked staging buffers that a few host threads fill (memcpy)
// while the copy engine drains the ones before (one cudaMemcpyAsync per piece; a plain cudaMemcpyAsync from pageable memory
// goes through the driver's single staging thread at a sixth of the link's speed, and page-locking gigabytes of existing
// memory costs more than the copy it speeds up).
struct UploadRing {
  static constexpr int SLOTS = 4;
  static constexpr size_t SLOT_BYTES = (size_t)16 << 20;
  // parallel_for(n_parts, body(part)): the context's parked worker threads
  std::function<void(size_t, const std::function<void(size_t)>&)> parallel_for;
  void copy(void* dst_device, const void* src_host, size_t bytes, cudaStream_t s);
  void release();
};

// The aligned reads of a BAM in HBM, structure-of-arrays in file order: what crosses PCIe on the device staging path
// (about 2.3 bytes per aligned base: base code, quality, and 60 bytes per read).
struct ReadsDev {
  DevBuf<int32_t> tid, pos, xl, xr;
  DevBuf<uint16_t> flag;
  DevBuf<uint8_t> mapq, rg, bases, quals;
  DevBuf<uint32_t> x1, l_seq, n_cigar, cigars;
  DevBuf<uint64_t> seq_off, cigar_off;
  uint64_t n = 0, bytes = 0;
  UploadRing ring;
  void upload(const ReadBatch& R, cudaStream_t s);
  RawReads view() const;
  void release();
};

// The staged stream in HBM (brq_types.h: PileupStream, the arrays the kernels read).
struct StreamDev {
  DevBuf<uint32_t> score_rec, side_rec, side_off, round_slot, score_cnt, round_side;
  DevBuf<uint64_t> score_off, hist_off, round_off;
  DevBuf<uint8_t> hist_rec, slot_ref, slot_group;
  size_t hist_exc_at = 0;   // byte offset of the exception records inside hist_rec (compact form)
  void release();
};

// Scratch of the expander, kept between staging calls.
struct ExpandScratch {
  DevBuf<ReadMeta> meta;
  DevBuf<ExpandSeg> seg;
  DevBuf<int32_t> max_span, seg_of_tid;
  DevBuf<uint8_t> ref, sub_k, col_red, col_qstart;
  DevBuf<uint32_t> part, stats, sub_first, red_cnt, side_cnt, side_red_cnt, hist_cnt, sub_cur, block_entries, block_base, round_vecs, scan_tmp32;
  DevBuf<uint64_t> ins_mask, geo_stats, scan_tmp, ins_parent, totals, read_starts;
  DevBuf<uint32_t> ins_count;
  DevBuf<uint8_t> hist_pos;   // positional histogram records before compaction
  ExpandArgs walk_args;       // the last staging's reads, segments and tiles (valid until the next staging)
  bool have_walk_args = false;
  void release();
};

// Builds the stream of `reads` (already in HBM; `host` supplies the per-target read ranges) for the visited targets / shard
// of `cfg` into `out`.  `st` receives the stream's metadata (segments, counts, geometry, the small ins_parent / ins_count
// arrays); its record arrays stay null and st.device_built is set.  Synchronises `stream` a few times (sizes of the arrays
// it allocates).  Throws what staging.cpp's stage() throws.
void expand_on_device(const BamHeader& hdr, const RefSet& ref, const ReadBatch& host, const ReadsDev& reads, const StageConfig& cfg,
                      ExpandScratch& scratch, StreamDev& out, PileupStream& st, cudaStream_t stream);

// Flagged slots' records for the host re-evaluation: per slot i of `slots`, its device words in record order and its side-list
// entries, contiguous (finalize.cpp: FlaggedRecords).
struct FlaggedRecordsHost {
  std::vector<uint64_t> word_off;   // [n + 1] into words
  std::vector<uint64_t> side_off;   // [n + 1] into side (entries, side_stride words each)
  std::vector<uint32_t> words, side;
  std::vector<uint8_t> ref;         // slot_ref of every slot
};
void gather_flagged_records(const StreamDev& ds, const PileupStream& st, const uint32_t* d_slots, uint32_t n, FlaggedRecordsHost& out,
                            cudaStream_t stream, uint64_t* d2h_bytes);

// BAM2COV (coverage_output.cpp:307-470): per base slot of the staged range, the coverage table's counts, from a walk over the
// reads of the last staging
// (group: COVERAGE_ALL_GROUPS, or the index of the one read group whose reads count; include_deleted: pass 2's notion of
// coverage, where a deletion over the column counts)
void coverage_columns_on_device(const ExpandScratch& scratch, uint64_t n_base, DevBuf<CoverageColumn>& out, uint32_t group, bool include_deleted,
                                cudaStream_t stream);

int expand_launch_count();

}  // namespace brq
