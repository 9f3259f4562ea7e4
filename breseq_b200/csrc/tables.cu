// Likelihood tables of pass 2, built on the device from the error table.
//
// One thread per record class (read set, strand, MAPQ present in the stream, quality, observed base):
//   pr[b] = (1 - e) * P[class with ref = b, complemented on the bottom strand] + e / 5,  e = 10^(-MAPQ/10)
//   L[b] = log10 pr[b],  M = max_b L[b],  r[b] = 10^(L[b] - M)          identify_mutations.cpp:3359-3384
// and writes the class into every table that holds it: the full table (fit kernel, global), the
// {L, M} table of all MAPQ values (tally kernel, global), and for the dominant MAPQ the fit
// kernel's shared-memory image {r, M} and the tally kernel's image [obs][set, strand, quality] x {L, M}.
// CUDA's log10/pow are within a couple of ulp of glibc's, i.e. 1e-16 relative on every term, far inside
// the 1e-9 bar of the log-likelihood sums; slots whose decisions are close are re-evaluated on the
// host with glibc in arrival order (finalize.cpp).
#include "kernels.h"
#include "brq_types.h"

namespace brq {

void note_launches(int n);

__global__ void __launch_bounds__(256) build_tables_kernel(TableBuildArgs a, ScoreParams p) {
  const uint32_t W = p.n_rpos * p.n_rep;
  const uint32_t n_cls = a.n_st * a.n_mapq_slots * a.Q * W * 5u;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cls) return;
  const uint32_t obs = i % 5u, rr = (i / 5u) % W, q = (i / (5u * W)) % a.Q, ms = (i / (5u * W * a.Q)) % a.n_mapq_slots,
                 st = i / (5u * W * a.Q * a.n_mapq_slots);
  const uint32_t rpos = rr / p.n_rep, rpt = rr % p.n_rep;
  const uint32_t set = st >> 1, top = st & 1u;
  const uint32_t mapq = a.slot_mapq[ms];
  const double incorrect = pow(10.0, -(double)mapq / 10.0), correct = 1.0 - incorrect, uniform = 1.0 / 5.0;
  const uint32_t o = top ? obs : (obs < 4u ? 3u - obs : 4u);
  double L[5], r[5], M = -1.7976931348623157e308;
#pragma unroll
  for (uint32_t b = 0; b < 5; ++b) {
    const uint32_t rf = top ? b : (b < 4u ? 3u - b : 4u);
    double pr = correct * a.prob[set * a.off_set + rf * a.off_ref + o * a.off_obs + q * a.off_qual + rpos * a.off_rpos + rpt * a.off_rep] + incorrect * uniform;
    if (pr < 0.0) pr = 0.0;
    L[b] = log10(pr);
    M = fmax(M, L[b]);
  }
#pragma unroll
  for (int b = 0; b < 5; ++b) r[b] = pow(10.0, L[b] - M);

  ClassTerms t;
#pragma unroll
  for (int b = 0; b < 5; ++b) { t.L[b] = L[b]; t.r[b] = r[b]; }
  t.r2 = 0.0; t.M = M;
  a.lut[i] = t;
  // the class's best likelihood ratio against its own observation, max over b != obs of 10^(L[b] - L[obs]): what the tally
  // kernel's presence bound needs of a record besides its terms (score_slots.cu)
  double rho = 0.0;
#pragma unroll
  for (uint32_t b = 0; b < 5; ++b) if (b != obs) rho = fmax(rho, pow(10.0, L[b] - L[obs]));
  HotTerms c;
#pragma unroll
  for (int b = 0; b < 5; ++b) c.L[b] = L[b];
  c.M = rho; c.pad[0] = 0.0; c.pad[1] = 0.0;
  a.coldT[((((size_t)st * p.n_mq + (mapq - p.mq_min)) * a.Q + q) * W + rr) * 5u + obs] = c;
  if (ms != a.hot_slot) return;
  if (p.n_hot) {
    HotRatios h;
#pragma unroll
    for (int b = 0; b < 5; ++b) h.r[b] = r[b];
    h.M = M;
    a.hotR[((size_t)st * a.Q + q) * 5u + obs] = h;
  }
  if (q >= p.t_qlo && q < p.t_qlo + p.t_nq) {
    double* d = reinterpret_cast<double*>(reinterpret_cast<char*>(a.tallyT) + (size_t)obs * p.t_stride + ((size_t)st * p.t_nq + (q - p.t_qlo)) * 64u);
#pragma unroll
    for (int b = 0; b < 5; ++b) d[b] = L[b];
    d[5] = rho; d[6] = top ? 1.0 : 0.0; d[7] = top ? 0.0 : 1.0;  // the last two columns count the class by strand
  }
}

void launch_build_tables(const TableBuildArgs& a, const ScoreParams& p, cudaStream_t s) {
  const uint32_t n_cls = a.n_st * a.n_mapq_slots * a.Q * p.n_rpos * p.n_rep * 5u;
  if (!n_cls) return;
  build_tables_kernel<<<(n_cls + 255) / 256, 256, 0, s>>>(a, p);
  note_launches(1);
}

}  // namespace brq
