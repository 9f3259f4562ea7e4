// TEST INFRASTRUCTURE -- part of oracle/, never linked into the product library.
//
// htslib-compatible shim over zlib (see htslib/sam.h in this directory for scope and the
// "unpinned at the htslib boundary" caveat).  Restates, from the SAM/BAM specification and
// from knowledge of htslib 1.2x sam.c:
//   * BGZF member inflate + BAM header / record parse        (SAMv1 section 4)
//   * bam_aux_get / bam_aux2i / bam_aux2Z                     (aux TLV walk)
//   * faidx over a plain FASTA
//   * sam_itr_queryi / sam_itr_next as a linear scan of a coordinate-sorted BAM
//   * the pileup engine: bam_plp_push / bam_plp_next / resolve_cigar2 semantics
//     (qpos on deletions = first read base AFTER the deletion; indel look-ahead merging of
//      consecutive D / I runs with P skipped; only BAM_FUNMAP and tid<0 dropped at push).
//     THE FLAG MASK (decided, not guessed twice): bam_plp_init() still stores BAM_DEF_MASK (UNMAP | SECONDARY | QCFAIL | DUP)
//     in iter->flag_mask, but since htslib 1.0 bam_plp_push() tests only BAM_FUNMAP -- its own comment reads "Skip only
//     unmapped reads here, any additional filtering must be done in iter->func" -- which is why samtools mpileup filters
//     the other three flags in its read callback (mplp_func, --excl-flags).  breseq's read callbacks (pileup_base.cpp:225-236,
//     290-301) filter nothing, so SECONDARY / QCFAIL / DUP records reach both pileup callbacks and are counted.  The
//     product (csrc/staging.cpp in_pileup, csrc/expand_core.h PILEUP_FLAG_MASK) does the same; tests/test_pileup_semantics.py
//     feeds flagged reads to the oracle and the staging layer and checks the hand-derived counts.
// The whole BAM is inflated into memory at hts_open(): this is a checker, not a product.
#include "htslib/sam.h"
#include "htslib/faidx.h"

#include <zlib.h>
#include <cassert>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

[[noreturn]] void die(const char* msg) {
  fprintf(stderr, "hts_shim: %s\n", msg);
  abort();
}

bool read_file(const char* fn, std::vector<uint8_t>& out) {
  FILE* f = fopen(fn, "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize((size_t)n);
  size_t got = n ? fread(out.data(), 1, (size_t)n, f) : 0;
  fclose(f);
  return got == (size_t)n;
}

// Inflate every BGZF member of a file image (SAMv1 4.1: gzip members with a 'BC' extra subfield
// holding BSIZE; payload is raw deflate; ISIZE in the trailer).
bool bgzf_inflate_all(const std::vector<uint8_t>& in, std::vector<uint8_t>& out) {
  size_t p = 0;
  while (p + 18 <= in.size()) {
    if (in[p] != 31 || in[p + 1] != 139 || in[p + 2] != 8 || !(in[p + 3] & 4)) return false;
    uint16_t xlen = in[p + 10] | (in[p + 11] << 8);
    size_t x = p + 12, xend = x + xlen;
    int bsize = -1;
    while (x + 4 <= xend) {
      uint16_t slen = in[x + 2] | (in[x + 3] << 8);
      if (in[x] == 'B' && in[x + 1] == 'C' && slen == 2) bsize = in[x + 4] | (in[x + 5] << 8);
      x += 4 + slen;
    }
    if (bsize < 0) return false;
    size_t block_len = (size_t)bsize + 1;
    if (p + block_len > in.size()) return false;
    const uint8_t* cdata = &in[xend];
    size_t clen = block_len - (xend - p) - 8;
    uint32_t isize;
    memcpy(&isize, &in[p + block_len - 4], 4);
    size_t o = out.size();
    out.resize(o + isize);
    if (isize) {
      z_stream zs;
      memset(&zs, 0, sizeof zs);
      if (inflateInit2(&zs, -15) != Z_OK) return false;
      zs.next_in = const_cast<Bytef*>(cdata);
      zs.avail_in = (uInt)clen;
      zs.next_out = &out[o];
      zs.avail_out = isize;
      int r = inflate(&zs, Z_FINISH);
      inflateEnd(&zs);
      if (r != Z_STREAM_END) return false;
    }
    p += block_len;
  }
  return p == in.size();
}

}  // namespace

extern "C" {

struct htsFile {
  std::vector<uint8_t> u;                 // uncompressed BAM image
  size_t first_rec = 0;                   // offset of the first alignment record
  size_t cur = 0;                         // sequential read cursor
  std::vector<std::vector<size_t>> by_tid;  // record offsets per tid (file order)
  bool indexed = false;
  int n_ref = 0;
};
struct hts_idx_t { htsFile* fp; };
struct hts_itr_t { htsFile* fp; int tid; hts_pos_t beg, end; size_t next; };

int kputs(const char* p, kstring_t* s) {
  size_t l = strlen(p);
  if (s->l + l + 1 > s->m) { s->m = (s->l + l + 1) * 2; s->s = (char*)realloc(s->s, s->m); }
  memcpy(s->s + s->l, p, l + 1);
  s->l += l;
  return (int)l;
}
int kputc(int c, kstring_t* s) {
  char b[2] = {(char)c, 0};
  kputs(b, s);
  return c;
}

bam1_t* bam_init1(void) { return (bam1_t*)calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t* b) {
  if (!b) return;
  free(b->data);
  free(b);
}
bam1_t* bam_copy1(bam1_t* d, const bam1_t* s) {
  if ((uint32_t)s->l_data > d->m_data) {
    d->m_data = (uint32_t)s->l_data;
    d->data = (uint8_t*)realloc(d->data, d->m_data);
  }
  memcpy(d->data, s->data, (size_t)s->l_data);
  d->l_data = s->l_data;
  d->core = s->core;
  d->id = s->id;
  return d;
}
int64_t bam_cigar2qlen(int n, const uint32_t* c) {
  int64_t l = 0;
  for (int k = 0; k < n; ++k)
    if (bam_cigar_type(bam_cigar_op(c[k])) & 1) l += bam_cigar_oplen(c[k]);
  return l;
}
hts_pos_t bam_cigar2rlen(int n, const uint32_t* c) {
  hts_pos_t l = 0;
  for (int k = 0; k < n; ++k)
    if (bam_cigar_type(bam_cigar_op(c[k])) & 2) l += bam_cigar_oplen(c[k]);
  return l;
}
hts_pos_t bam_endpos(const bam1_t* b) {
  hts_pos_t rlen = (b->core.flag & BAM_FUNMAP) ? 0 : bam_cigar2rlen((int)b->core.n_cigar, bam_get_cigar(b));
  if (rlen == 0) rlen = 1;
  return b->core.pos + rlen;
}

static int aux_type_size(int t) {
  switch (t) {
    case 'A': case 'c': case 'C': return 1;
    case 's': case 'S': return 2;
    case 'i': case 'I': case 'f': return 4;
    case 'd': return 8;
    default: return 0;
  }
}
static const uint8_t* aux_skip(const uint8_t* s, const uint8_t* end) {
  // s points at the type byte
  int t = *s++;
  int sz = aux_type_size(t);
  if (sz) return s + sz;
  if (t == 'Z' || t == 'H') {
    while (s < end && *s) ++s;
    return s + 1;
  }
  if (t == 'B') {
    int st = *s++;
    uint32_t n;
    memcpy(&n, s, 4);
    return s + 4 + (size_t)n * aux_type_size(st);
  }
  die("bad aux type");
}
uint8_t* bam_aux_get(const bam1_t* b, const char tag[2]) {
  const uint8_t* s = bam_get_aux(b);
  const uint8_t* end = b->data + b->l_data;
  while (s + 3 <= end) {
    if (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]) return const_cast<uint8_t*>(s + 2);
    s = aux_skip(s + 2, end);
  }
  return NULL;
}
int64_t bam_aux2i(const uint8_t* s) {
  int t = *s++;
  switch (t) {
    case 'c': return (int8_t)*s;
    case 'C': return *s;
    case 's': { int16_t v; memcpy(&v, s, 2); return v; }
    case 'S': { uint16_t v; memcpy(&v, s, 2); return v; }
    case 'i': { int32_t v; memcpy(&v, s, 4); return v; }
    case 'I': { uint32_t v; memcpy(&v, s, 4); return v; }
    default: return 0;
  }
}
char* bam_aux2Z(const uint8_t* s) {
  int t = *s++;
  if (t == 'Z' || t == 'H') return (char*)const_cast<uint8_t*>(s);
  return NULL;
}
int bam_aux_append(bam1_t*, const char[2], char, int, const uint8_t*) { die("bam_aux_append: writer stub"); }
int bam_aux_del(bam1_t*, uint8_t*) { die("bam_aux_del: writer stub"); }

htsFile* hts_open(const char* fn, const char* mode) {
  if (!mode || mode[0] != 'r') die("hts_open: only reading is supported by the shim");
  std::vector<uint8_t> raw;
  if (!read_file(fn, raw)) return NULL;
  htsFile* fp = new htsFile;
  if (!bgzf_inflate_all(raw, fp->u)) { delete fp; return NULL; }
  if (fp->u.size() < 12 || memcmp(fp->u.data(), "BAM\1", 4) != 0) { delete fp; return NULL; }
  return fp;
}
int hts_close(htsFile* fp) { delete fp; return 0; }
int hts_set_opt(htsFile*, int, ...) { return 0; }

static int32_t rd_i32(const std::vector<uint8_t>& u, size_t p) { int32_t v; memcpy(&v, &u[p], 4); return v; }

sam_hdr_t* sam_hdr_read(samFile* fp) {
  const std::vector<uint8_t>& u = fp->u;
  size_t p = 4;
  int32_t l_text = rd_i32(u, p); p += 4;
  sam_hdr_t* h = (sam_hdr_t*)calloc(1, sizeof(sam_hdr_t));
  h->text = (char*)malloc((size_t)l_text + 1);
  memcpy(h->text, &u[p], (size_t)l_text);
  h->text[l_text] = 0;
  h->l_text = strlen(h->text);
  p += (size_t)l_text;
  h->n_targets = rd_i32(u, p); p += 4;
  h->target_name = (char**)calloc((size_t)h->n_targets + 1, sizeof(char*));
  h->target_len = (uint32_t*)calloc((size_t)h->n_targets + 1, sizeof(uint32_t));
  for (int i = 0; i < h->n_targets; ++i) {
    int32_t l_name = rd_i32(u, p); p += 4;
    h->target_name[i] = (char*)malloc((size_t)l_name);
    memcpy(h->target_name[i], &u[p], (size_t)l_name);
    p += (size_t)l_name;
    h->target_len[i] = (uint32_t)rd_i32(u, p); p += 4;
  }
  fp->first_rec = fp->cur = p;
  fp->n_ref = h->n_targets;
  return h;
}
void sam_hdr_destroy(sam_hdr_t* h) {
  if (!h) return;
  for (int i = 0; i < h->n_targets; ++i) free(h->target_name[i]);
  free(h->target_name);
  free(h->target_len);
  free(h->text);
  free(h);
}
sam_hdr_t* sam_hdr_parse(size_t, const char*) { die("sam_hdr_parse: writer stub"); }
int sam_hdr_write(samFile*, const sam_hdr_t*) { die("sam_hdr_write: writer stub"); }
int sam_write1(samFile*, const sam_hdr_t*, const bam1_t*) { die("sam_write1: writer stub"); }
int sam_parse1(kstring_t*, sam_hdr_t*, bam1_t*) { die("sam_parse1: writer stub"); }

// @XX header line access (text scan; the reference calls these once per BAM open:
// /root/reference/src/breseq/alignment.cpp:545-563).
static std::vector<std::string> hdr_lines(sam_hdr_t* h, const char* type) {
  std::vector<std::string> out;
  const char* s = h->text;
  while (s && *s) {
    const char* e = strchr(s, '\n');
    std::string line = e ? std::string(s, e) : std::string(s);
    if (line.size() >= 3 && line[0] == '@' && line[1] == type[0] && line[2] == type[1]) out.push_back(line);
    s = e ? e + 1 : NULL;
  }
  return out;
}
static bool line_tag(const std::string& line, const char* key, std::string& val) {
  size_t p = 3;
  while (p < line.size()) {
    size_t q = line.find('\t', p + 1);
    if (line[p] == '\t') {
      std::string f = line.substr(p + 1, (q == std::string::npos ? line.size() : q) - p - 1);
      if (f.size() >= 3 && f[0] == key[0] && f[1] == key[1] && f[2] == ':') { val = f.substr(3); return true; }
    }
    if (q == std::string::npos) break;
    p = q;
  }
  return false;
}
int sam_hdr_count_lines(sam_hdr_t* h, const char* type) { return (int)hdr_lines(h, type).size(); }
const char* sam_hdr_line_name(sam_hdr_t* h, const char* type, int pos) {
  static thread_local std::string keep;
  std::vector<std::string> l = hdr_lines(h, type);
  if (pos < 0 || pos >= (int)l.size()) return NULL;
  const char* key = (type[0] == 'S' && type[1] == 'Q') ? "SN" : (type[0] == 'R' && type[1] == 'G') ? "ID"
                    : (type[0] == 'P' && type[1] == 'G') ? "ID" : NULL;
  if (!key || !line_tag(l[(size_t)pos], key, keep)) return NULL;
  return keep.c_str();
}
int sam_hdr_find_tag_pos(sam_hdr_t* h, const char* type, int pos, const char* key, kstring_t* ks) {
  std::vector<std::string> l = hdr_lines(h, type);
  if (pos < 0 || pos >= (int)l.size()) return -2;
  std::string v;
  if (!line_tag(l[(size_t)pos], key, v)) return -1;
  ks->l = 0;
  kputs(v.c_str(), ks);
  return 0;
}

// One BAM record at uncompressed offset p -> b. Returns bytes consumed, 0 at EOF.
static size_t parse_record(const std::vector<uint8_t>& u, size_t p, bam1_t* b) {
  if (p + 4 > u.size()) return 0;
  int32_t block = rd_i32(u, p);
  if (p + 4 + (size_t)block > u.size()) die("truncated BAM record");
  const uint8_t* x = &u[p + 4];
  bam1_core_t& c = b->core;
  int32_t tid, pos, l_seq, mtid, mpos, tlen;
  uint8_t l_read_name, mapq;
  uint16_t bin, n_cigar, flag;
  memcpy(&tid, x, 4); memcpy(&pos, x + 4, 4);
  l_read_name = x[8]; mapq = x[9];
  memcpy(&bin, x + 10, 2); memcpy(&n_cigar, x + 12, 2); memcpy(&flag, x + 14, 2);
  memcpy(&l_seq, x + 16, 4); memcpy(&mtid, x + 20, 4); memcpy(&mpos, x + 24, 4); memcpy(&tlen, x + 28, 4);
  c.tid = tid; c.pos = pos; c.l_qname = l_read_name; c.qual = mapq; c.bin = bin; c.n_cigar = n_cigar;
  c.flag = flag; c.l_qseq = l_seq; c.mtid = mtid; c.mpos = mpos; c.isize = tlen; c.l_extranul = 0;
  int l_data = block - 32;
  if ((uint32_t)l_data > b->m_data) { b->m_data = (uint32_t)l_data; b->data = (uint8_t*)realloc(b->data, b->m_data); }
  memcpy(b->data, x + 32, (size_t)l_data);
  b->l_data = l_data;
  return 4 + (size_t)block;
}

int sam_read1(samFile* fp, sam_hdr_t*, bam1_t* b) {
  size_t n = parse_record(fp->u, fp->cur, b);
  if (!n) return -1;
  fp->cur += n;
  return 0;
}

hts_idx_t* sam_index_load(htsFile* fp, const char*) {
  // "Index" = one pass recording each record's offset under its tid.
  if (!fp->indexed) {
    if (fp->first_rec == 0) { sam_hdr_t* h = sam_hdr_read(fp); sam_hdr_destroy(h); }
    fp->by_tid.assign((size_t)fp->n_ref, std::vector<size_t>());
    size_t p = fp->first_rec;
    while (p + 4 <= fp->u.size()) {
      int32_t block = rd_i32(fp->u, p);
      int32_t tid = rd_i32(fp->u, p + 4);
      if (tid >= 0 && tid < fp->n_ref) fp->by_tid[(size_t)tid].push_back(p);
      p += 4 + (size_t)block;
    }
    fp->indexed = true;
  }
  hts_idx_t* idx = new hts_idx_t;
  idx->fp = fp;
  return idx;
}
void hts_idx_destroy(hts_idx_t* idx) { delete idx; }

hts_itr_t* sam_itr_queryi(const hts_idx_t* idx, int tid, hts_pos_t beg, hts_pos_t end) {
  if (!idx || tid < 0 || tid >= idx->fp->n_ref) return NULL;
  hts_itr_t* it = new hts_itr_t;
  it->fp = idx->fp; it->tid = tid; it->beg = beg; it->end = end; it->next = 0;
  return it;
}
int sam_itr_next(htsFile* fp, hts_itr_t* it, bam1_t* b) {
  const std::vector<size_t>& v = fp->by_tid[(size_t)it->tid];
  while (it->next < v.size()) {
    parse_record(fp->u, v[it->next++], b);
    if (b->core.pos >= it->end) { it->next = v.size(); return -1; }  // sorted input: nothing further overlaps
    if (bam_endpos(b) > it->beg) return 0;
  }
  return -1;
}
void hts_itr_destroy(hts_itr_t* it) { delete it; }

// ---------------------------------------------------------------- faidx
struct faidx_t {
  std::vector<std::string> names;
  std::map<std::string, std::string> seqs;
};
faidx_t* fai_load(const char* fn) {
  FILE* f = fopen(fn, "r");
  if (!f) return NULL;
  faidx_t* fai = new faidx_t;
  std::string cur;
  char* line = NULL;
  size_t cap = 0;
  ssize_t n;
  while ((n = getline(&line, &cap, f)) >= 0) {
    while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
    if (line[0] == '>') {
      cur.assign(line + 1);
      size_t sp = cur.find_first_of(" \t");
      if (sp != std::string::npos) cur.resize(sp);
      fai->names.push_back(cur);
      fai->seqs[cur] = "";
    } else if (!cur.empty()) {
      fai->seqs[cur].append(line, (size_t)n);
    }
  }
  free(line);
  fclose(f);
  return fai;
}
void fai_destroy(faidx_t* fai) { delete fai; }
char* fai_fetch(const faidx_t* fai, const char* reg, int* len) {
  std::map<std::string, std::string>::const_iterator it = fai->seqs.find(reg);
  if (it == fai->seqs.end()) { *len = -2; return NULL; }
  *len = (int)it->second.size();
  char* s = (char*)malloc(it->second.size() + 1);
  memcpy(s, it->second.c_str(), it->second.size() + 1);
  return s;
}
int faidx_nseq(const faidx_t* fai) { return (int)fai->names.size(); }
const char* faidx_iseq(const faidx_t* fai, int i) { return fai->names[(size_t)i].c_str(); }
int faidx_seq_len(const faidx_t* fai, const char* seq) {
  std::map<std::string, std::string>::const_iterator it = fai->seqs.find(seq);
  return it == fai->seqs.end() ? -1 : (int)it->second.size();
}

// ---------------------------------------------------------------- pileup engine
struct cstate_t { int k; hts_pos_t x, y, end; };
struct lbnode_t {
  bam1_t b;
  hts_pos_t beg, end;
  cstate_t s;
};
struct bam_plp_s {
  std::vector<lbnode_t*> list;  // active reads, arrival order (htslib: singly linked list head..tail)
  bam1_t* b;
  bam_plp_auto_f func;
  void* data;
  int tid, max_tid;
  hts_pos_t pos, max_pos;
  bool is_eof;
  int maxcnt;
  uint64_t id;
  std::vector<bam_pileup1_t> plp;
};

bam_plp_t bam_plp_init(bam_plp_auto_f func, void* data) {
  bam_plp_s* it = new bam_plp_s;
  it->b = bam_init1();
  it->func = func;
  it->data = data;
  it->tid = 0; it->pos = 0;
  it->max_tid = -1; it->max_pos = -1;
  it->is_eof = false;
  it->maxcnt = 8000;
  it->id = 0;
  return it;
}
void bam_plp_destroy(bam_plp_t it) {
  for (size_t i = 0; i < it->list.size(); ++i) { free(it->list[i]->b.data); delete it->list[i]; }
  bam_destroy1(it->b);
  delete it;
}
void bam_plp_set_maxcnt(bam_plp_t it, int maxcnt) { it->maxcnt = maxcnt; }

#define _cop(c) ((c) & BAM_CIGAR_MASK)
#define _cln(c) ((c) >> BAM_CIGAR_SHIFT)
static inline bool op_is_match(int op) { return op == BAM_CMATCH || op == BAM_CEQUAL || op == BAM_CDIFF; }
static inline bool op_is_refwalk(int op) { return op_is_match(op) || op == BAM_CDEL || op == BAM_CREF_SKIP; }

// s->k: index of the current reference-consuming op; s->x: its reference start; s->y: query
// bases consumed before it.
static void resolve_cigar(bam_pileup1_t* p, hts_pos_t pos, cstate_t* s) {
  bam1_t* b = p->b;
  bam1_core_t* c = &b->core;
  uint32_t* cigar = bam_get_cigar(b);
  int k;
  if (s->k == -1) {  // first column for this read: find the first M/D/N/=/X
    p->qpos = 0;
    if (c->n_cigar == 1) {
      if (op_is_match(_cop(cigar[0]))) { s->k = 0; s->x = c->pos; s->y = 0; }
    } else {
      for (k = 0, s->x = c->pos, s->y = 0; k < (int)c->n_cigar; ++k) {
        int op = _cop(cigar[k]);
        int l = _cln(cigar[k]);
        if (op_is_refwalk(op)) break;
        else if (op == BAM_CINS || op == BAM_CSOFT_CLIP) s->y += l;
      }
      assert(k < (int)c->n_cigar);
      s->k = k;
    }
  } else {
    int l = _cln(cigar[s->k]);
    if (pos - s->x >= l) {  // advance to the next reference-consuming op
      assert(s->k < (int)c->n_cigar);
      if (op_is_match(_cop(cigar[s->k]))) s->y += l;
      s->x += l;
      for (k = s->k + 1; k < (int)c->n_cigar; ++k) {
        int op = _cop(cigar[k]);
        l = _cln(cigar[k]);
        if (op_is_refwalk(op)) break;
        else if (op == BAM_CINS || op == BAM_CSOFT_CLIP) s->y += l;
      }
      s->k = k;
      assert(s->k < (int)c->n_cigar);
    }
  }
  {
    int op = _cop(cigar[s->k]);
    int l = _cln(cigar[s->k]);
    p->is_del = p->indel = p->is_refskip = 0;
    if (s->x + l - 1 == pos && s->k + 1 < (int)c->n_cigar) {  // last column of this op: peek ahead
      int op2 = _cop(cigar[s->k + 1]);
      int l2 = _cln(cigar[s->k + 1]);
      if (op2 == BAM_CDEL && op != BAM_CDEL) {
        p->indel = -(int)l2;
        for (k = s->k + 2; k < (int)c->n_cigar; ++k) {
          op2 = _cop(cigar[k]); l2 = _cln(cigar[k]);
          if (op2 == BAM_CDEL) p->indel -= l2;
          else break;
        }
      } else if (op2 == BAM_CINS) {
        p->indel = l2;
        for (k = s->k + 2; k < (int)c->n_cigar; ++k) {
          op2 = _cop(cigar[k]); l2 = _cln(cigar[k]);
          if (op2 == BAM_CINS) p->indel += l2;
          else if (op2 != BAM_CPAD) break;
        }
      } else if (op2 == BAM_CPAD && s->k + 2 < (int)c->n_cigar) {
        int l3 = 0;
        for (k = s->k + 2; k < (int)c->n_cigar; ++k) {
          op2 = _cop(cigar[k]); l2 = _cln(cigar[k]);
          if (op2 == BAM_CINS) l3 += l2;
          else if (op_is_refwalk(op2)) break;
        }
        if (l3 > 0) p->indel = l3;
      }
    }
    if (op_is_match(op)) {
      p->qpos = (int32_t)(s->y + (pos - s->x));
    } else if (op == BAM_CDEL || op == BAM_CREF_SKIP) {
      p->is_del = 1;
      p->qpos = (int32_t)s->y;
      p->is_refskip = (op == BAM_CREF_SKIP);
    }
    p->is_head = (pos == c->pos);
    p->is_tail = (pos == s->end);
  }
  p->cigar_ind = s->k;
}

static int plp_push(bam_plp_s* it, const bam1_t* b) {
  if (!b) { it->is_eof = true; return 0; }
  if (b->core.tid < 0) return 0;
  if (b->core.flag & BAM_FUNMAP) return 0;
  hts_pos_t beg = b->core.pos, end = bam_endpos(b);
  if (b->core.tid < it->max_tid || (b->core.tid == it->max_tid && beg < it->max_pos)) die("pileup: unsorted input");
  it->max_tid = b->core.tid;
  it->max_pos = beg;
  if (end > it->pos || b->core.tid > it->tid) {
    lbnode_t* n = new lbnode_t;
    memset(&n->b, 0, sizeof(bam1_t));
    bam_copy1(&n->b, b);
    n->b.id = it->id++;
    n->beg = beg; n->end = end;
    n->s.k = -1; n->s.x = 0; n->s.y = 0; n->s.end = end - 1;
    it->list.push_back(n);
  }
  return 0;
}

static const bam_pileup1_t* plp_next(bam_plp_s* it, int* _tid, int* _pos, int* _n_plp) {
  *_n_plp = 0;
  if (it->is_eof && it->list.empty()) return NULL;
  while (it->is_eof || it->max_tid > it->tid || (it->max_tid == it->tid && it->max_pos > it->pos)) {
    it->plp.clear();
    size_t w = 0;
    for (size_t r = 0; r < it->list.size(); ++r) {
      lbnode_t* p = it->list[r];
      if (p->b.core.tid < it->tid || (p->b.core.tid == it->tid && p->end <= it->pos)) {
        free(p->b.data);
        delete p;
        continue;
      }
      it->list[w++] = p;
      if (p->b.core.tid == it->tid && p->beg <= it->pos) {
        bam_pileup1_t e;
        memset(&e, 0, sizeof e);
        e.b = &p->b;
        resolve_cigar(&e, it->pos, &p->s);
        it->plp.push_back(e);
      }
    }
    it->list.resize(w);
    int n_plp = (int)it->plp.size();
    *_n_plp = n_plp; *_tid = it->tid; *_pos = (int)it->pos;
    if (!it->list.empty()) {
      lbnode_t* head = it->list[0];
      if (it->tid > head->b.core.tid) die("pileup: unsorted input");
      if (it->tid < head->b.core.tid) { it->tid = head->b.core.tid; it->pos = head->beg; }
      else if (it->pos < head->beg) it->pos = head->beg;
      else ++it->pos;
    } else {
      ++it->pos;
    }
    if (n_plp) return it->plp.data();
    if (it->is_eof && it->list.empty()) break;
  }
  return NULL;
}

const bam_pileup1_t* bam_plp_auto(bam_plp_t it, int* _tid, int* _pos, int* _n_plp) {
  const bam_pileup1_t* plp;
  if ((plp = plp_next(it, _tid, _pos, _n_plp)) != 0) return plp;
  *_n_plp = 0;
  if (it->is_eof) return 0;
  while (it->func(it->data, it->b) >= 0) {
    plp_push(it, it->b);
    if ((plp = plp_next(it, _tid, _pos, _n_plp)) != 0) return plp;
  }
  plp_push(it, 0);
  if ((plp = plp_next(it, _tid, _pos, _n_plp)) != 0) return plp;
  return 0;
}

}  // extern "C"
