// Shared host-side types of the B200 read-alignment evidence pileup.
//
// Data flow:  BAM (or the synthetic generator)  ->  ReadBatch  ->  stage()  ->  PileupStream
//             (pinned, columnar, reference-position sorted)  ->  CUDA kernels.
//
// Vocabulary follows the reference: "column" = one reference position of one target,
// "insert sub-column" = the k-th inserted base after a column
// (/root/reference/src/breseq/identify_mutations.cpp:1359), "slot" = a column or a sub-column,
// "record" = one (read, slot) incidence.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace brq {

// Base indices: A,C,G,T,'.' = 0..4, N = 5 (/root/reference/src/breseq/common.h:202-316).
enum : uint8_t { kBaseA = 0, kBaseC = 1, kBaseG = 2, kBaseT = 3, kBaseGap = 4, kBaseN = 5, kBaseNul = 6 };

inline uint8_t nibble_to_index(uint8_t bam4) {  // BAM 4-bit code -> index; anything not ACGT is N
  switch (bam4) { case 1: return 0; case 2: return 1; case 4: return 2; case 8: return 3; default: return 5; }
}
inline uint8_t char_to_index(char c) {
  switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
               case '.': return 4; case 'N': return 5; default: return 255; }
}
inline char index_to_char(uint8_t b) { return b < 6 ? "ACGT.N"[b] : '?'; }

struct RefSet {
  std::vector<std::string> names;
  std::vector<std::string> seqs;  // as stored in the FASTA (no case folding, like fai_fetch)
  uint64_t total_length() const { uint64_t t = 0; for (auto& s : seqs) t += s.size(); return t; }
};

// One sequencing read file set = one SAM read group (@RG ID/LB = base name); a paired set owns
// two read files (/root/reference/src/breseq/alignment.h:576-591).
struct ReadGroups {
  std::vector<std::string> ids;
  std::vector<std::string> libraries;
};

// std::vector whose resize() leaves new elements uninitialised: the read arrays are gigabytes that the decode threads fill
// at once, and value-initialising them first is a single-threaded pass over all of that memory.
template <class T>
struct NoInit : std::allocator<T> {
  template <class U> struct rebind { using other = NoInit<U>; };
  NoInit() = default;
  template <class U> NoInit(const NoInit<U>&) {}
  template <class U, class... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
  template <class U> void construct(U* p) { ::new (static_cast<void*>(p)) U; }
};
template <class T> using RawVec = std::vector<T, NoInit<T>>;

// Reads of one BAM, structure-of-arrays, in file (coordinate) order.
struct ReadBatch {
  RawVec<int32_t> tid, pos;       // 0-based leftmost reference position
  RawVec<uint16_t> flag;
  RawVec<uint8_t> mapq;
  RawVec<uint8_t> rg;             // resolved read-group index (0 when unresolvable)
  RawVec<uint32_t> x1;            // X1:i redundancy; 1 when the tag is absent
  RawVec<int32_t> xl, xr;         // XL/XR:i trims; -1 when the tag is absent
  RawVec<int32_t> as;             // AS:i (carried for the writer only)
  RawVec<uint32_t> l_seq;
  RawVec<uint64_t> seq_off;       // into bases/quals
  RawVec<uint32_t> n_cigar;
  RawVec<uint64_t> cigar_off;     // into cigars
  RawVec<uint8_t> bases;          // one BAM 4-bit code per byte (1,2,4,8,15)
  RawVec<uint8_t> quals;          // raw phred
  RawVec<uint32_t> cigars;        // BAM encoding: len<<4 | op
  std::vector<std::string> names;      // optional (writer); may be empty
  size_t size() const { return tid.size(); }
};

// ---- packed stream records -------------------------------------------------------------

// ---- scoring (identify_mutations) records ------------------------------------------------
// One record per (read, slot) whose base at that slot is not N.  What a record MEANS is the
// "classic" word below; what the device stream HOLDS is the table-coordinate form after it.
//
// Classic word (host semantics; the side list and the host re-evaluation use it):
//   [2:0]   obs        base index 0..4 ('.' = 4)
//   [9:3]   qual       quality chosen by alignment_position_to_covariates (error_count.cpp:1049-1105)
//   [10]    top        1 = read on the top strand
//   [24]    unique     X1 == 1
//   [25]    trimmed    is_trimmed() (alignment.h:389-410)
//   [26]    ok         covariates resolvable (not past q_end, no N at the quality position)
//   [27]    match      obs equals the slot's reference base (side-list entries only)
//   unique records:    [15:11] read_set (flat read-file index), [23:16] mapq
//   redundant records: [23:11] redundancy (X1, saturated at 8191)
// A record SCORES when it is unique, untrimmed, ok and qual >= Settings::base_quality_cutoff.
constexpr uint32_t SR_OBS_SHIFT = 0, SR_QUAL_SHIFT = 3, SR_TOP_BIT = 1u << 10, SR_SET_SHIFT = 11, SR_MAPQ_SHIFT = 16,
                   SR_UNIQUE_BIT = 1u << 24, SR_TRIM_BIT = 1u << 25, SR_OK_BIT = 1u << 26, SR_MATCH_BIT = 1u << 27,
                   SR_RED_SHIFT = 11, SR_RED_MASK = 0x1FFF;

// Device stream word (score_rec), 4 bytes.  The staging layer has already classified the record and
// resolved it to (a) the counter it increments in the tally kernel's per-slot class histogram and (b) its
// cell in the shared-memory likelihood table (geometry chosen per stream: ScoreGeometry).  The kernel
// counts first and multiplies the counts into the table once per slot, so the common record costs one
// AND/OR for its address and a byte increment.
//   [12:0]  counter  the record's byte counter inside its lane's histogram: word * 128 + byte * 8, i.e. the word's byte
//                    offset in [12:7] and the counter's bit position inside the word in [4:0] (the kernel adds
//                    1 << (record & 31) to the word at (record & 0x1F80) | lane base).
//                    HOT record matching the slot's reference base: class sq = (read_set*2 + top) * n_q + qual - q_lo,
//                    word sq / 4, byte sq % 4.  Every other record counts in the special words that follow the
//                    class words (ScoreGeometry::special_counter).
//   [13]    top      1 = read on the top strand (all kinds)
//   [14]    (unused)
//   [15]    trimmed  REDUNDANT: is_trimmed() (only the per-position debug file looks at it)
//   [23:16] sq       HOT: class index (above)         REDUNDANT: [24:16] X1; 511 = the value is the slot's next
//   [26:24] obs      HOT: observed base A,C,G,T,'.'              SIDE_BIG entry of the side list; [27:25] observed base
//   [28]    match    HOT (every HOT record matches the slot's reference base)
//   [31:30] kind     0 HOT    scores; dominant MAPQ, quality inside the table window, observation = the slot's reference
//                             base ('.' included: a read without an inserted base is a '.' observation of the insert
//                             sub-column, and nearly every record of such a slot is one)
//                    1 IDLE   unique but does not score (trimmed, unresolvable, quality below the cutoff)
//                    2 COLD   scores, but not as a count of a shared-table class: another MAPQ, a quality outside the
//                             window, or an observation that does not match the reference base (a sequencing error
//                             or a variant, one record in a thousand: the presence bound of the tally kernel needs
//                             them one by one).  Its classic word is the slot's next cold entry of the side list
//                    3 REDUNDANT
// Within a slot the REDUNDANT records come first and the others follow, each part in arrival (BAM) order.
// The stream is ROUND-MAJOR AND LANE-INTERLEAVED.  The tally kernel works in rounds of 32 slots (round_slot: same
// reference base, similar depth), one slot per lane.  Round r owns words [round_off[r], round_off[r + 1]): one 1 KB
// "round vector" (ROUND_VECTOR_WORDS) per eight records of its deepest slot.  Round vector i holds records 8i .. 8i+7
// of every lane: lane l's records 8i .. 8i+3 at word i*256 + 4l, its records 8i+4 .. 8i+7 at word i*256 + 128 + 4l, so
// a warp reads a round vector with two fully coalesced 128-bit accesses per lane.  Lanes whose slot is shallower, and
// idle lanes, are filled with pad words (the trash counter, no other bit: never a real record).
// score_off[s] = first word of slot s (round_off[r] + 4 l), score_cnt[s] = its records; record j of slot s is word
// score_index(score_off[s], j).
//
// Side list (side_rec, CSR side_off per slot, counted in entries): in stream order, the classic words of the slot's
// COLD records, and for redundant records with X1 >= 511 an entry SIDE_BIG | X1.  A stream staged for the read_pos /
// base_repeat covariates (ScoreGeometry::side_stride = 2) has no shared table: every scoring record is COLD and every
// entry is two words, the classic word and [15:0] read_pos, [23:16] base_repeat of the record's quality position
// (error_count.cpp:1049-1105).
constexpr uint32_t DR_COUNTER_MASK = 0x1FFFu, DR_COUNTER_WORD_MASK = 0x1F80u, DR_TOP_BIT = 1u << 13, DR_SQ_SHIFT = 16, DR_SQ_MASK = 0xFFu,
                   DR_OBS_SHIFT = 24, DR_RED_TRIM_BIT = 1u << 15, DR_RED_OBS_SHIFT = 25, DR_X1_SHIFT = 16, DR_X1_MASK = 0x1FFu, DR_MATCH_BIT = 1u << 28,
                   DR_KIND_SHIFT = 30, DR_IDLE = 1u << 30, DR_COLD = 2u << 30, DR_REDUNDANT = 3u << 30,
                   SIDE_BIG = 1u << 31, SIDE_PAD = 0xFFFFFFFFu;  // SIDE_PAD fills a slot's side range to an even count
// special counters, in the two histogram words after the class words
enum : uint32_t { SC_IDLE_TOP = 0, SC_IDLE_BOT = 1, SC_COLD_TOP = 2, SC_COLD_BOT = 3, SC_SLOW_TOP = 4, SC_SLOW_BOT = 5, SC_TRASH = 6 };

// Table geometry baked into the device stream words of one staged stream.
struct ScoreGeometry {
  uint32_t cutoff = 3;     // Settings::base_quality_cutoff the stream was staged for
  uint32_t hot_mapq = 0;   // the MAPQ value whose classes the shared table holds
  uint32_t q_lo = 0, n_q = 0;   // quality window of the shared table (n_q is a multiple of 4)
  uint32_t n_st = 2;       // (read sets) x 2 strands
  uint32_t side_stride = 1;  // words per side-list entry: 2 when the records carry read_pos / base_repeat
  uint32_t n_sq() const { return n_st * n_q; }            // classes of the per-slot histogram (<= 248)
  uint32_t n_words() const { return n_sq() / 4 + 2; }     // 32-bit histogram words per lane: class words + 2 special words
  uint32_t n_hot() const { return n_sq() * 5; }           // cells of the shared likelihood table
  static uint32_t counter_of(uint32_t index) { return (index >> 2) * 128u + (index & 3u) * 8u; }
  uint32_t special_counter(uint32_t which) const { return counter_of(n_sq() + which); }
  uint32_t pad_word() const { return special_counter(SC_TRASH); }
};

// TRANSFER FORM of score_rec (score16 + score_exc): what crosses PCIe.  The low half of a device word decides the word
// for every kind of record but two: a HOT record matching its slot's reference base is its counter (class = the
// counter's word and byte, observation = the slot's base), IDLE / COLD / pad words are their special counter.  Staging
// sends the low halves (u16, same round-major geometry as score_rec) with bit 15 set on the records whose word does not
// come back that way (REDUNDANT ones), and those words in full, per lane of every round
// in record order (score_exc, CSR score_exc_off[round * 32 + lane]); expand_score_kernel rebuilds score_rec in HBM,
// bit for bit (staging checks every word), so the kernels and the host never see the transfer form.
struct ScoreRecon { uint32_t c_idle_top, c_idle_bot, c_cold_top, c_cold_bot, c_trash; };
inline ScoreRecon score_recon_of(const ScoreGeometry& g) {
  return {g.special_counter(SC_IDLE_TOP), g.special_counter(SC_IDLE_BOT), g.special_counter(SC_COLD_TOP), g.special_counter(SC_COLD_BOT),
          g.special_counter(SC_TRASH)};
}
constexpr uint32_t S16_EXCEPTION = 0x8000u;
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t score_word_from16(uint32_t lo, uint32_t ref, const ScoreRecon& g) {  // lo: 15 bits, ref: the slot's base index
  const uint32_t c = lo & DR_COUNTER_MASK, top = lo & DR_TOP_BIT;
  if (c == g.c_trash) return lo;
  if (c == (top ? g.c_idle_top : g.c_idle_bot)) return DR_IDLE | lo;
  if (c == (top ? g.c_cold_top : g.c_cold_bot)) return DR_COLD | lo;
  const uint32_t sq = ((c >> 7) << 2) | ((c >> 3) & 3u);
  return sq << DR_SQ_SHIFT | ref << DR_OBS_SHIFT | DR_MATCH_BIT | lo;
}

// classic word of a HOT device word
inline uint32_t classic_of_hot(uint32_t d, const ScoreGeometry& g) {
  const uint32_t sq = (d >> DR_SQ_SHIFT) & DR_SQ_MASK, obs = (d >> DR_OBS_SHIFT) & 7u, qual = g.q_lo + sq % g.n_q, st = sq / g.n_q;
  return obs | qual << SR_QUAL_SHIFT | st << 10 | g.hot_mapq << SR_MAPQ_SHIFT | SR_UNIQUE_BIT | SR_OK_BIT |
         ((d & DR_MATCH_BIT) ? SR_MATCH_BIT : 0u);
}

// Histogram (error_count) record, one per unique, non-deleted (read, column).  Both observations of
// cErrorTable::count_alignment_position (error_count.cpp:854-986) are resolved by the staging layer
// into table coordinates on the READ strand (bases complemented for reversed reads):
//   observation A, the aligned base:           [2:0] ref  [5:3] obs  [12:6] quality  [13] valid
//                                              (valid = neither the read base nor the reference base is N)
//   observation B, what follows it in the read: [16:14] ref [19:17] obs [26:20] quality [27] valid
//       next base also aligned      ('.', '.')         quality of the next base on the read strand
//       deletion of exactly 1 base  (ref base, '.')    quality of the next base on the read strand
//       insertion of exactly 1 base ('.', inserted)    quality of the inserted base
//   [30:28] read_set (low three bits)
//   [31]    fast: the dominant kind of record, which the kernel counts with ONE atomic on a joint histogram: observation A
//           valid with ref == obs, and observation B either ('.', '.') or absent (then [26:20] reads 127)
// That is the whole record (4 bytes) unless the run uses the read_pos / base_repeat covariates or has
// more than 8 read files; then records are 8 bytes and the high word adds
//   [15:0] read_pos of A (0-based query index)  [23:16] base_repeat of A  [28:24] base_repeat of B (saturated at 31)
//   [31:29] read_set bits 5:3
// Base indices A,C,G,T,'.' = 0..4.
// COMPACT FORM (hist16 + hist_exc).  All but ~0.3 % of the 4-byte records are `fast`, and a fast record is four small
// numbers: staging re-packs it into 16 bits
//   [1:0] base (ref == obs)   [7:2] quality of A   [13:8] quality of B, 63 = no observation B   [15:14] read_set
// and moves every other record (not fast, a quality above 62, a read_set above 3) unchanged into the exception stream
// hist_exc.  The device reads only these two streams: half the bytes over PCIe and out of HBM.  Order is not kept
// (a histogram does not need it); hist_rec / hist_off stay on the host as the positional form (tests, shard merges).
constexpr int HR_REFA = 0, HR_OBSA = 3, HR_QUALA = 6, HR_VALIDA = 13, HR_REFB = 14, HR_OBSB = 17, HR_QUALB = 20, HR_VALIDB = 27,
              HR_SET = 28, HR_FAST = 31, HR_RPOS = 32, HR_REPA = 48, HR_REPB = 56, HR_SET_HI = 61;

#ifdef __CUDACC__
#define BRQ_HD __host__ __device__
#else
#define BRQ_HD
#endif
// 16-bit form of a fast record -> the 4-byte record it stands for
BRQ_HD inline uint32_t hist16_expand(uint32_t r) {
  const uint32_t base = r & 3u, qa = (r >> 2) & 63u, qb = (r >> 8) & 63u, set = r >> 14;
  return base | base << HR_OBSA | qa << HR_QUALA | 1u << HR_VALIDA | set << HR_SET | 1u << HR_FAST |
         (qb == 63u ? 127u << HR_QUALB : (4u << HR_REFB | 4u << HR_OBSB | qb << HR_QUALB | 1u << HR_VALIDB));
}
// 4-byte record -> its 16-bit form, or 0x10000 when it has none (expansion must give the record back bit for bit)
inline uint32_t hist16_pack(uint32_t lo) {
  if (!(lo >> HR_FAST)) return 0x10000u;
  const uint32_t qa = (lo >> HR_QUALA) & 127u, qb = (lo >> HR_QUALB) & 127u, set = (lo >> HR_SET) & 7u;
  if (qa > 62u || (qb > 62u && qb != 127u) || set > 3u) return 0x10000u;
  const uint32_t r = (lo & 3u) | qa << 2 | (qb == 127u ? 63u : qb) << 8 | set << 14;
  return hist16_expand(r) == lo ? r : 0x10000u;
}

// User evidence (Settings::user_evidence_genome_diff_file_name, identify_mutations.cpp:879, 1013-1020): RA rows the user wants
// reported whatever the data says, stripped to their specification and sorted like cGenomeDiff::sort().
struct UserRa { std::string seq_id; uint32_t position = 0, insert_position = 0; std::string ref_base, new_base; };
// A column where the pileup meets the list: the insert sub-columns the list forces there (identify_mutations.cpp:1346-1355)
// and the entries it consumes, by insert level (:1914-2019).
struct UserColumn {
  int32_t tid = 0; uint32_t pos1 = 0;          // target and 1-based position
  uint64_t slot = 0;                            // base slot, ~0 when the column is outside this shard
  uint32_t force_max = 0;
  std::vector<std::pair<uint32_t, uint32_t>> consumed;  // (insert level, index into the list), in list order
};

// Columns [lo, hi) (0-based) of BAM target `tid` occupy base slots slot0 .. slot0 + (hi - lo).
struct Segment { int32_t tid, lo, hi; uint64_t slot0; };

// The staged, columnar, position-sorted stream handed to the device.
struct PileupStream {
  // geometry
  std::vector<Segment> segments;       // visited targets in visit (alphabetical seq id) order, clipped to this shard
  std::vector<Segment> visit_targets;  // every visited target, whole (slot0 = its first column in the concatenated visit order)
  uint64_t n_base = 0;                 // base columns (sum of segment lengths)
  uint64_t n_ins = 0;                  // insert sub-column slots, appended after the base slots
  std::vector<uint64_t> ins_parent;    // [n_ins] base slot of each sub-column
  std::vector<uint32_t> ins_count;     // [n_ins] insert_count (>= 1)
  // per slot
  uint8_t* slot_ref = nullptr;         // [n_base + n_ins] reference base index ('.' for sub-columns)
  uint64_t* score_off = nullptr;       // [n_base + n_ins + 1] first word of every slot in score_rec (see above; score_index())
  uint32_t* score_cnt = nullptr;       // [n_base + n_ins] records of every slot
  uint64_t* round_off = nullptr;       // [n_rounds + 1] first word of every round in score_rec
  uint32_t* round_side = nullptr;      // [n_rounds * 32 * 2] per lane: side-list begin | reference base << 29, side-list end (used entries)
  uint64_t* hist_off = nullptr;        // [n_base + 1] CSR into hist_rec; bit 63 of entry c = column c has a redundant read
  uint8_t* slot_group = nullptr;       // [n_base] coverage group of the column's target
  // records
  uint32_t* score_rec = nullptr;       // device stream words
  uint16_t* score16 = nullptr;         // transfer form of score_rec (above): low halves, [n_score_padded]
  uint32_t* score_exc = nullptr;       // the words the low halves do not determine, per (round, lane) in record order
  uint32_t* score_exc_off = nullptr;   // [n_rounds * 32 + 1] CSR into score_exc
  uint64_t n_score_exc = 0;
  uint32_t* side_rec = nullptr;        // side list: classic words of COLD records, SIDE_BIG | X1 of very redundant ones
  uint32_t* side_off = nullptr;        // [n_base + n_ins + 1] CSR into side_rec
  uint64_t n_side = 0;
  uint32_t* round_slot = nullptr;      // [n_rounds * 32] slot of every lane of every tally round, ROUND_NO_SLOT = idle lane
  uint64_t n_rounds = 0;
  ScoreGeometry geo;
  void* hist_rec = nullptr;            // n_hist records of hist_bytes (4 or 8) each
  uint32_t hist_bytes = 4;
  uint16_t* hist16 = nullptr;          // compact form of the fast records (above); null when hist_bytes == 8
  uint32_t* hist_exc = nullptr;        // the records without a 16-bit form, unchanged
  uint64_t n_hist16 = 0, n_hist_exc = 0;
  uint64_t n_score = 0, n_hist = 0;    // records (padding not counted)
  uint64_t n_score_padded = 0;         // words in score_rec
  uint32_t mapq_seen[8] = {0};         // 256-bit mask of MAPQ values present among scoring records
  uint64_t mapq_count[256] = {0};      // scoring records per MAPQ value
  uint64_t qual_count[128] = {0};      // scoring records per quality value
  uint64_t max_hist_depth = 0;         // deepest unique, non-deleted column (sizes the coverage histogram)
  uint32_t n_groups = 1;               // coverage groups present
  std::vector<UserRa> user_list;            // the user evidence the stream was staged with
  std::vector<UserColumn> user_columns;     // ... and where the pileup meets it, in visit order
  std::vector<uint64_t> read_start_counts;  // preprocess stage: [BAM tid][0 = without, 1 = with a read start] position-strand combinations of this shard
  uint32_t max_qual_seen = 0;
  uint32_t max_hist_qual = 0, max_hist_rpos = 0;  // largest quality / read position in a valid histogram observation
  uint32_t max_score_rpos = 0;                    // largest read position of a scoring record (streams staged with read_pos)
  uint32_t max_read_set_seen = 0;
  bool pinned = false;                 // buffers came from cudaHostAlloc
  bool device_built = false;           // built in HBM by the expander (expand.cu): the record arrays above are null on the host
  bool hist_compact = false;           // the device reads the compact histogram streams (hist16 + hist_exc)
  uint64_t bytes_uploaded = 0;         // device_built: bytes of reads and reference that crossed PCIe for it
  bool score_rec_plain = false, hist_rec_plain = false;  // ... except these two: plain memory when only their transfer / compact form is uploaded
  uint64_t n_slots() const { return n_base + n_ins; }
};

constexpr uint64_t HIST_OFF_REDUNDANT_BIT = 1ull << 63;
constexpr uint32_t ROUND_NO_SLOT = 0xFFFFFFFFu, ROUND_BLOCK = 4096, ROUND_VECTOR_WORDS = 256;

// word of record j of the slot whose first word is `base` (round-major, lane-interleaved stream)
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint64_t score_index(uint64_t base, uint64_t j) {
  return base + (j >> 3) * ROUND_VECTOR_WORDS + ((j >> 2) & 1u) * (ROUND_VECTOR_WORDS / 2) + (j & 3u);
}

// Classic words of one slot in stream order (redundant first): f(classic word, X1, ext).  word(j) = the slot's j-th device
// word, side = the slot's side-list entries (ss words each).  Redundant records come back with their full X1 (the classic
// field saturates at 8191); ext = read_pos | base_repeat << 16 of a scoring record of a stream staged with them, else 0.
template <class W, class F>
inline void for_each_classic_words(const ScoreGeometry& geo, uint64_t cnt, W&& word, const uint32_t* side_entries, F&& f) {
  const uint32_t ss = geo.side_stride;
  uint32_t side = 0;
  for (uint64_t j = 0; j < cnt; ++j) {
    const uint32_t d = word(j), kind = d >> DR_KIND_SHIFT, top = (d & DR_TOP_BIT) ? SR_TOP_BIT : 0u;
    if (kind == 0) f(classic_of_hot(d, geo), 1u, 0u);
    else if (kind == 1) f(top | SR_UNIQUE_BIT | SR_TRIM_BIT, 1u, 0u);   // does not score; why is not kept
    else if (kind == 2) { f(side_entries[(size_t)side * ss], 1u, ss == 2 ? side_entries[(size_t)side * ss + 1] : 0u); ++side; }
    else {
      uint32_t x1 = (d >> DR_X1_SHIFT) & DR_X1_MASK;
      if (x1 == DR_X1_MASK) x1 = side_entries[(size_t)(side++) * ss] & ~SIDE_BIG;
      f(top | ((d >> DR_RED_OBS_SHIFT) & 7u) | ((d & DR_RED_TRIM_BIT) ? SR_TRIM_BIT : 0u) | (x1 < SR_RED_MASK ? x1 : SR_RED_MASK) << SR_RED_SHIFT, x1, 0u);
    }
  }
}
// ... of slot s of a stream whose arrays are on the host
template <class F>
inline void for_each_classic(const PileupStream& st, uint64_t s, F&& f) {
  const uint64_t base = st.score_off[s];
  for_each_classic_words(st.geo, st.score_cnt[s], [&](uint64_t j) { return st.score_rec[score_index(base, j)]; },
                         st.side_rec + (size_t)st.side_off[s] * st.geo.side_stride, f);
}

}  // namespace brq
