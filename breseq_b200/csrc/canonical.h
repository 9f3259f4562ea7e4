// The value a double has after the reference's text round trip: write_log10_prob_table streams it with six
// significant digits ("%.6g", error_count.cpp:660-690) and read_log10_prob_table reads it back with strtod (:629-654).
//
// text_canonical_6g(x) computes that value without the text, in exact arithmetic, so the same function runs on the
// device (between table derivation and the likelihood-table build, no host round trip inside a step) and on the
// host (tests/canonical_check.cpp compares it with snprintf + strtod on millions of values, half-way cases included):
//   |x| = m * 10^X with 1 <= m < 10;   D = the integer nearest to |x| * 10^(5 - X), ties to even ON THE EXACT PRODUCT
//   (printf rounds the exact binary value);   result = D / 10^(5 - X), one correctly rounded division of two exactly
//   representable numbers, which is what a correctly rounding strtod returns for the decimal D * 10^(X - 5).
// Supported range: 10^-17 <= |x| < 10^6 (5 - X in [0, 22]: 10^(5 - X) is exact); outside it `ok` is cleared and x
// comes back unchanged.  Table values are log10 probabilities, |x| in (1e-13, 10).
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define BRQ_CANON_HD __host__ __device__
#else
#define BRQ_CANON_HD
#endif

namespace brq {

BRQ_CANON_HD inline double pow10_exact(int k) {  // 10^k, k in [0, 22]: exactly representable
  const double t[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20,
                        1e21, 1e22};
  return t[k];
}

// the integer nearest to a * 10^k (a > 0, k in [0, 22]), ties to even on the exact product
BRQ_CANON_HD inline double scaled_round_exact(double a, int k) {
  const double s = pow10_exact(k);
  const double hi = a * s, lo = fma(a, s, -hi);  // a * s = hi + lo exactly
  double D = rint(hi);
  const double r = hi - D;                       // exact; |r| <= 1/2, and |lo| is below the spacing of r's possible values
  if (r == 0.5 && lo > 0.0) D += 1.0;            // hi sits on a half: the exact product decides
  else if (r == -0.5 && lo < 0.0) D -= 1.0;
  return D;
}

BRQ_CANON_HD inline double text_canonical_6g(double x, bool* ok) {
  if (x == 0.0 || !(x == x)) return x;
  const double a = fabs(x);
  if (!(a >= 1e-17 && a < 1e6)) { if (ok) *ok = false; return x; }
  int X = (int)floor(log10(a));  // may be off by one next to a power of ten: corrected by the range of D
  double D = 0.0;
  for (int it = 0; it < 4; ++it) {
    if (X > 5) X = 5;
    if (X < -17) X = -17;
    D = scaled_round_exact(a, 5 - X);
    if (D >= 1e6) { if (X == 5) break; ++X; }
    else if (D < 1e5) { if (X == -17) break; --X; }
    else break;
  }
  if (!(D >= 1e5 && D < 1e6)) { if (ok) *ok = false; return x; }
  const double v = D / pow10_exact(5 - X);
  return x < 0.0 ? -v : v;
}

}  // namespace brq
