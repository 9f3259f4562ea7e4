/* TEST INFRASTRUCTURE -- see sam.h in this directory. faidx subset over a plain FASTA. */
#ifndef ORACLE_HTS_SHIM_FAIDX_H
#define ORACLE_HTS_SHIM_FAIDX_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct faidx_t faidx_t;
faidx_t *fai_load(const char *fn);
void fai_destroy(faidx_t *fai);
/* region is a bare sequence name here (the only form the reference's hot path uses:
 * /root/reference/src/breseq/pileup_base.cpp:41). Returns a malloc'd NUL-terminated copy. */
char *fai_fetch(const faidx_t *fai, const char *reg, int *len);
int faidx_nseq(const faidx_t *fai);
const char *faidx_iseq(const faidx_t *fai, int i);
int faidx_seq_len(const faidx_t *fai, const char *seq);
#ifdef __cplusplus
}
#endif
#endif
