/* TEST INFRASTRUCTURE -- part of oracle/, never linked into the product library.
 *
 * Minimal htslib-compatible surface (htslib 1.23.1 is the reference's pinned, un-vendored
 * dependency: /root/reference/dev-environment.yml; it is NOT installed in this image).
 * Only the ~45 symbols the reference touches (sole include site:
 * /root/reference/src/breseq/common.h:88-91) are declared.  The implementation in
 * ../hts_shim.cpp restates the PUBLISHED behaviour of htslib's BGZF/BAM reader,
 * aux-tag access, faidx and the pileup engine (sam.c: bam_plp_push / bam_plp_next /
 * resolve_cigar2) from the format specification (SAMv1.pdf) and from knowledge of
 * htslib's sources; it could not be diffed against a real htslib here, so every
 * parity claim that passes through it is "unpinned at the htslib boundary"
 * (see DESIGN.md).
 */
#ifndef ORACLE_HTS_SHIM_SAM_H
#define ORACLE_HTS_SHIM_SAM_H

#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t hts_pos_t;

/* ---- kstring ---- */
typedef struct kstring_t { size_t l, m; char *s; } kstring_t;
#define KS_INITIALIZE { 0, 0, NULL }
static inline void ks_free(kstring_t *s) { if (s) { free(s->s); s->l = s->m = 0; s->s = NULL; } }
int kputs(const char *p, kstring_t *s);
int kputc(int c, kstring_t *s);

/* ---- CIGAR / flags ---- */
#define BAM_CMATCH 0
#define BAM_CINS 1
#define BAM_CDEL 2
#define BAM_CREF_SKIP 3
#define BAM_CSOFT_CLIP 4
#define BAM_CHARD_CLIP 5
#define BAM_CPAD 6
#define BAM_CEQUAL 7
#define BAM_CDIFF 8
#define BAM_CBACK 9
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK 0xf
#define BAM_CIGAR_TYPE 0x3C1A7
#define bam_cigar_op(c) ((c) & BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c) >> BAM_CIGAR_SHIFT)
#define bam_cigar_type(o) (BAM_CIGAR_TYPE >> ((o) << 1) & 3)

#define BAM_FPAIRED 1
#define BAM_FPROPER_PAIR 2
#define BAM_FUNMAP 4
#define BAM_FMUNMAP 8
#define BAM_FREVERSE 16
#define BAM_FMREVERSE 32
#define BAM_FREAD1 64
#define BAM_FREAD2 128
#define BAM_FSECONDARY 256
#define BAM_FQCFAIL 512
#define BAM_FDUP 1024
#define BAM_FSUPPLEMENTARY 2048

/* ---- records ---- */
typedef struct bam1_core_t {
  hts_pos_t pos;
  int32_t tid;
  uint16_t bin;
  uint8_t qual;
  uint8_t l_extranul;
  uint16_t flag;
  uint16_t l_qname;
  uint32_t n_cigar;
  int32_t l_qseq;
  int32_t mtid;
  hts_pos_t mpos;
  hts_pos_t isize;
} bam1_core_t;

typedef struct bam1_t {
  bam1_core_t core;
  uint64_t id;
  uint8_t *data;
  int l_data;
  uint32_t m_data;
  uint32_t mempolicy;
} bam1_t;

#define bam_is_rev(b) (((b)->core.flag & BAM_FREVERSE) != 0)
#define bam_is_mrev(b) (((b)->core.flag & BAM_FMREVERSE) != 0)
#define bam_get_qname(b) ((char *)(b)->data)
#define bam_get_cigar(b) ((uint32_t *)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_get_aux(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1) + (b)->core.l_qseq)
#define bam_get_l_aux(b) ((b)->l_data - ((b)->core.n_cigar << 2) - (b)->core.l_qname - (b)->core.l_qseq - (((b)->core.l_qseq + 1) >> 1))
#define bam_seqi(s, i) ((s)[(i) >> 1] >> ((~(i) & 1) << 2) & 0xf)

bam1_t *bam_init1(void);
void bam_destroy1(bam1_t *b);
bam1_t *bam_copy1(bam1_t *bdst, const bam1_t *bsrc);
int64_t bam_cigar2qlen(int n_cigar, const uint32_t *cigar);
hts_pos_t bam_cigar2rlen(int n_cigar, const uint32_t *cigar);
hts_pos_t bam_endpos(const bam1_t *b);

uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]);
int64_t bam_aux2i(const uint8_t *s);
char *bam_aux2Z(const uint8_t *s);
int bam_aux_append(bam1_t *b, const char tag[2], char type, int len, const uint8_t *data);
int bam_aux_del(bam1_t *b, uint8_t *s);

/* ---- header ---- */
typedef struct sam_hdr_t {
  int32_t n_targets;
  int32_t ignore_sam_err;
  size_t l_text;
  uint32_t *target_len;
  const int8_t *cigar_tab;
  char **target_name;
  char *text;
  void *sdict;
  void *hrecs;
  uint32_t ref_count;
} sam_hdr_t;
typedef sam_hdr_t bam_hdr_t;

/* ---- files / iterators ---- */
typedef struct htsFile htsFile;
typedef htsFile samFile;
typedef struct hts_idx_t hts_idx_t;
typedef struct hts_itr_t hts_itr_t;

htsFile *hts_open(const char *fn, const char *mode);
int hts_close(htsFile *fp);
int hts_set_opt(htsFile *fp, int opt, ...);
hts_idx_t *sam_index_load(htsFile *fp, const char *fn);
void hts_idx_destroy(hts_idx_t *idx);
sam_hdr_t *sam_hdr_read(samFile *fp);
void sam_hdr_destroy(sam_hdr_t *h);
sam_hdr_t *sam_hdr_parse(size_t l_text, const char *text);
int sam_hdr_write(samFile *fp, const sam_hdr_t *h);
int sam_hdr_count_lines(sam_hdr_t *h, const char *type);
const char *sam_hdr_line_name(sam_hdr_t *h, const char *type, int pos);
int sam_hdr_find_tag_pos(sam_hdr_t *h, const char *type, int pos, const char *key, kstring_t *ks);
int sam_read1(samFile *fp, sam_hdr_t *h, bam1_t *b);
int sam_write1(samFile *fp, const sam_hdr_t *h, const bam1_t *b);
int sam_parse1(kstring_t *s, sam_hdr_t *h, bam1_t *b);
hts_itr_t *sam_itr_queryi(const hts_idx_t *idx, int tid, hts_pos_t beg, hts_pos_t end);
int sam_itr_next(htsFile *fp, hts_itr_t *iter, bam1_t *b);
void hts_itr_destroy(hts_itr_t *iter);

/* ---- pileup ---- */
typedef union bam_pileup_cd { void *p; int64_t i; double f; } bam_pileup_cd;
typedef struct bam_pileup1_t {
  bam1_t *b;
  int32_t qpos;
  int indel, level;
  uint32_t is_del : 1, is_head : 1, is_tail : 1, is_refskip : 1, aux : 28;
  bam_pileup_cd cd;
  int cigar_ind;
} bam_pileup1_t;

typedef int (*bam_plp_auto_f)(void *data, bam1_t *b);
typedef struct bam_plp_s *bam_plp_t;
bam_plp_t bam_plp_init(bam_plp_auto_f func, void *data);
void bam_plp_destroy(bam_plp_t iter);
void bam_plp_set_maxcnt(bam_plp_t iter, int maxcnt);
const bam_pileup1_t *bam_plp_auto(bam_plp_t iter, int *_tid, int *_pos, int *_n_plp);

#ifdef __cplusplus
}
#endif
#endif
