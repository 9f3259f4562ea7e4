// Negative-binomial fit of the unique-only coverage histogram (see coverage_fit.h).
//
// What has to be bit-identical to the reference, and is: the objective (same terms in the same order through the same
// log-gamma, coverage_distribution.cpp:224-244, stats.h:92-100), the simplex search (stats.cpp:2289-2392) and the restart grid
// (:256-291), so that nb_mu / nb_size come out as the reference's doubles.  What only has to be accurate: the cumulative
// distribution behind the quantile (the Cephes incomplete beta function, stats.cpp:1385-1480): its results are only compared with a probability, but for
// Poisson-like histograms the comparison hangs on rounding noise, so that function is restated operation for operation, too.
#include "coverage_fit.h"
#include "stats_math.h"

#include <algorithm>
#include <cmath>
#include <fstream>
#include <limits>
#include <stdexcept>

namespace brq {

namespace {

// dnbinom(k, size = size, mu = mu) as stats.h:92-100 evaluates it
double nbinom_pmf(double k, double size, double mu) {
  const double p = size / (size + mu);
  const double log_pmf = log_gamma(k + size) - log_gamma(size) - log_gamma(k + 1) + size * log(p) + k * log(1.0 - p);
  return exp(log_pmf);
}

// ---- The regularized incomplete beta function I_x(a, b) behind the cumulative distribution, as the Cephes library computes it
// (incbet.c / gamma.c, Moshier 1984-2000: the reference carries a copy, stats.cpp:255-400, 409-470, 1385-1745).  The published
// algorithm is restated step for step, operation order included: a fit with a very large size parameter (a Poisson-like
// histogram) sends arguments here for which log Gamma(a + b) - log Gamma(a) cancels to a few digits, and the quantile that
// comes out depends on which digits; the reference's digits are the specification.  Positive arguments only.
namespace cephes {
const double MACHEP = 1.11022302462515654042E-16, MAXLOG = 7.09782712893383996732E2, MINLOG = -7.451332191019412076235E2;
const double MAXGAM = 171.624376956302725, BIG = 4.503599627370496e15, BIGINV = 2.22044604925031308085e-16;

double horner(double x, const double* c, int degree) {          // polevl: c[0] x^degree + ... + c[degree]
  double r = c[0];
  for (int i = 1; i <= degree; ++i) r = r * x + c[i];
  return r;
}
double horner_monic(double x, const double* c, int degree) {    // p1evl: leading coefficient 1, c[0..degree-1] follow
  double r = x + c[0];
  for (int i = 1; i < degree; ++i) r = r * x + c[i];
  return r;
}

double lgam(double x) {   // x > 0
  static const double A[5] = {8.11614167470508450300E-4, -5.95061904284301438324E-4, 7.93650340457716943945E-4,
                              -2.77777777730099687205E-3, 8.33333333333331927722E-2};
  static const double B[6] = {-1.37825152569120859100E3, -3.88016315134637840924E4, -3.31612992738871184744E5,
                              -1.16237097492762307383E6, -1.72173700820839662146E6, -8.53555664245765465627E5};
  static const double C[6] = {-3.51815701436523470549E2, -1.70642106651881159223E4, -2.20528590553854454839E5,
                              -1.13933444367982507207E6, -2.53252307177582951285E6, -2.01889141433532773231E6};
  if (x < 13.0) {
    double z = 1.0, shift = 0.0, u = x;
    while (u >= 3.0) { shift -= 1.0; u = x + shift; z *= u; }
    while (u < 2.0) {
      if (u == 0.0) return INFINITY;
      z /= u; shift += 1.0; u = x + shift;
    }
    if (z < 0.0) z = -z;
    if (u == 2.0) return log(z);
    shift -= 2.0;
    const double y = x + shift;
    return log(z) + y * horner(y, B, 5) / horner_monic(y, C, 6);
  }
  if (x > 2.556348e305) return INFINITY;
  double q = (x - 0.5) * log(x) - x + 0.91893853320467274178;
  if (x > 1.0e8) return q;
  const double p = 1.0 / (x * x);
  if (x >= 1000.0) q += ((7.9365079365079365079365e-4 * p - 2.7777777777777777777778e-3) * p + 0.0833333333333333333333) / x;
  else q += horner(p, A, 4) / x;
  return q;
}

double stirling(double x) {   // Gamma(x) for x > 33
  static const double S[5] = {7.87311395793093628397E-4, -2.29549961613378126380E-4, -2.68132617805781232825E-3,
                              3.47222221605458667310E-3, 8.33333333333482257126E-2};
  double w = 1.0 / x;
  w = 1.0 + w * horner(w, S, 4);
  double y = exp(x);
  if (x > 143.01608) { const double v = pow(x, 0.5 * x - 0.25); y = v * (v / y); }   // pow() alone would overflow
  else y = pow(x, x - 0.5) / y;
  return 2.50662827463100050242E0 * y * w;
}

double gamma(double x) {   // x > 0
  static const double P[7] = {1.60119522476751861407E-4, 1.19135147006586384913E-3, 1.04213797561761569935E-2, 4.76367800457137231464E-2,
                              2.07448227648435975150E-1, 4.94214826801497100753E-1, 9.99999999999999996796E-1};
  static const double Q[8] = {-2.31581873324120129819E-5, 5.39605580493303397842E-4, -4.45641913851797240494E-3, 1.18139785222060435552E-2,
                              3.58236398605498653373E-2, -2.34591795718243348568E-1, 7.14304917030273074085E-2, 1.00000000000000000320E0};
  if (fabs(x) > 33.0) return stirling(x);
  double z = 1.0;
  while (x >= 3.0) { x -= 1.0; z *= x; }
  while (x < 2.0) {
    if (x < 1.e-9) return x == 0.0 ? INFINITY : z / ((1.0 + 0.5772156649015329 * x) * x);
    z /= x; x += 1.0;
  }
  if (x == 2.0) return z;
  x -= 2.0;
  return z * horner(x, P, 6) / horner(x, Q, 7);
}

// the two continued fractions: `second` = false is expansion #1 (in x), true is expansion #2 (in x / (1 - x))
double fraction(double a, double b, double x, bool second) {
  double k1 = a, k2 = second ? b - 1.0 : a + b, k3 = a, k4 = a + 1.0, k5 = 1.0, k6 = second ? a + b : b - 1.0, k7 = a + 1.0, k8 = a + 2.0;
  const double z = second ? x / (1.0 - x) : x;
  double pkm2 = 0.0, qkm2 = 1.0, pkm1 = 1.0, qkm1 = 1.0, ans = 1.0, r = 1.0;
  const double thresh = 3.0 * MACHEP;
  int n = 0;
  do {
    double xk = -(z * k1 * k2) / (k3 * k4);
    double pk = pkm1 + pkm2 * xk, qk = qkm1 + qkm2 * xk;
    pkm2 = pkm1; pkm1 = pk; qkm2 = qkm1; qkm1 = qk;
    xk = (z * k5 * k6) / (k7 * k8);
    pk = pkm1 + pkm2 * xk; qk = qkm1 + qkm2 * xk;
    pkm2 = pkm1; pkm1 = pk; qkm2 = qkm1; qkm1 = qk;
    if (qk != 0) r = pk / qk;
    double t;
    if (r != 0) { t = fabs((ans - r) / r); ans = r; } else t = 1.0;
    if (t < thresh) break;
    k1 += 1.0; k2 += second ? -1.0 : 1.0; k3 += 2.0; k4 += 2.0; k5 += 1.0; k6 += second ? 1.0 : -1.0; k7 += 2.0; k8 += 2.0;
    if ((fabs(qk) + fabs(pk)) > BIG) { pkm2 *= BIGINV; pkm1 *= BIGINV; qkm2 *= BIGINV; qkm1 *= BIGINV; }
    if ((fabs(qk) < BIGINV) || (fabs(pk) < BIGINV)) { pkm2 *= BIG; pkm1 *= BIG; qkm2 *= BIG; qkm1 *= BIG; }
  } while (++n < 300);
  return ans;
}

double power_series(double a, double b, double x) {   // for small b x, x not close to 1
  const double ai = 1.0 / a;
  double u = (1.0 - b) * x, v = u / (a + 1.0);
  const double t1 = v, z = MACHEP * ai;
  double t = u, n = 2.0, s = 0.0;
  while (fabs(v) > z) {
    u = (n - b) * x / n;
    t *= u;
    v = t / (a + n);
    s += v;
    n += 1.0;
  }
  s += t1;
  s += ai;
  u = a * log(x);
  if ((a + b) < MAXGAM && fabs(u) < MAXLOG) {
    t = gamma(a + b) / (gamma(a) * gamma(b));
    return s * t * pow(x, a);
  }
  t = lgam(a + b) - lgam(a) - lgam(b) + u + log(s);
  return t < MINLOG ? 0.0 : exp(t);
}

double incbet(double aa, double bb, double xx) {
  if (aa <= 0.0 || bb <= 0.0) return 0.0;
  if (xx <= 0.0 || xx >= 1.0) return xx == 1.0 ? 1.0 : 0.0;
  if ((bb * xx) <= 1.0 && xx <= 0.95) return power_series(aa, bb, xx);
  double w = 1.0 - xx, a, b, x, xc, t;
  const bool flipped = xx > (aa / (aa + bb));   // work on the side of the mean where the expansions converge
  if (flipped) { a = bb; b = aa; xc = xx; x = w; } else { a = aa; b = bb; xc = w; x = xx; }
  if (flipped && (b * x) <= 1.0 && x <= 0.95) {
    t = power_series(a, b, x);
  } else {
    double y = x * (a + b - 2.0) - (a - 1.0);
    w = y < 0.0 ? fraction(a, b, x, false) : fraction(a, b, x, true) / xc;
    // times x^a (1 - x)^b Gamma(a + b) / (a Gamma(a) Gamma(b)), directly or through logarithms
    y = a * log(x);
    t = b * log(xc);
    if ((a + b) < MAXGAM && fabs(y) < MAXLOG && fabs(t) < MAXLOG) {
      t = pow(xc, b);
      t *= pow(x, a);
      t /= a;
      t *= w;
      t *= gamma(a + b) / (gamma(a) * gamma(b));
    } else {
      y += t + lgam(a + b) - lgam(a) - lgam(b);
      y += log(w / a);
      t = y < MINLOG ? 0.0 : exp(y);
    }
  }
  if (flipped) t = t <= MACHEP ? 1.0 - MACHEP : 1.0 - t;
  return t;
}

// ---- the inverse of incbet in x (Cephes incbi.c; the reference's copy: stats.cpp:1076-1318) and the inverse normal it starts
// from (ndtri.c; :1747-1796).  incbi alternates two searches until one of them is satisfied: bracketing with an adaptive split
// ("halve") and Newton steps from inside the bracket ("newton", entered once); the bracket may be mirrored (x -> 1 - x, a <-> b)
// when it closes in on 1.  Written as a three-state loop; every arithmetic step in the order the published code takes it.
double ndtri(double p) {
  static const double P0[5] = {-5.99633501014107895267E1, 9.80010754185999661536E1, -5.66762857469070293439E1, 1.39312609387279679503E1,
                               -1.23916583867381258016E0};
  static const double Q0[8] = {1.95448858338141759834E0, 4.67627912898881538453E0, 8.63602421390890590575E1, -2.25462687854119370527E2,
                               2.00260212380060660359E2, -8.20372256168333339912E1, 1.59056225126211695515E1, -1.18331621121330003142E0};
  static const double P1[9] = {4.05544892305962419923E0, 3.15251094599893866154E1, 5.71628192246421288162E1, 4.40805073893200834700E1,
                               1.46849561928858024014E1, 2.18663306850790267539E0, -1.40256079171354495875E-1, -3.50424626827848203418E-2,
                               -8.57456785154685413611E-4};
  static const double Q1[8] = {1.57799883256466749731E1, 4.53907635128879210584E1, 4.13172038254672030440E1, 1.50425385692907503408E1,
                               2.50464946208309415979E0, -1.42182922854787788574E-1, -3.80806407691578277194E-2, -9.33259480895457427372E-4};
  static const double P2[9] = {3.23774891776946035970E0, 6.91522889068984211695E0, 3.93881025292474443415E0, 1.33303460815807542389E0,
                               2.01485389549179081538E-1, 1.23716634817820021358E-2, 3.01581553508235416007E-4, 2.65806974686737550832E-6,
                               6.23974539184983293730E-9};
  static const double Q2[8] = {6.02427039364742014255E0, 3.67983563856160859403E0, 1.37702099489081330271E0, 2.16236993594496635890E-1,
                               1.34204006088543189037E-2, 3.28014464682127739104E-4, 2.89247864745380683936E-6, 6.79019408009981274425E-9};
  const double MAXNUM = 1.79769313486231570815E308, EXP_M2 = 0.13533528323661269189, SQRT_2PI = 2.50662827463100050242E0;
  if (p <= 0.0) return -MAXNUM;
  if (p >= 1.0) return MAXNUM;
  bool lower_tail = true;
  double y = p;
  if (y > 1.0 - EXP_M2) { y = 1.0 - y; lower_tail = false; }
  if (y > EXP_M2) {   // the middle: a rational function of (p - 1/2)^2
    y = y - 0.5;
    const double y2 = y * y;
    double x = y + y * (y2 * horner(y2, P0, 4) / horner_monic(y2, Q0, 8));
    return x * SQRT_2PI;
  }
  double x = sqrt(-2.0 * log(y));
  const double x0 = x - log(x) / x, z = 1.0 / x;
  const double x1 = x < 8.0 ? z * horner(z, P1, 8) / horner_monic(z, Q1, 8) : z * horner(z, P2, 8) / horner_monic(z, Q2, 8);
  x = x0 - x1;
  return lower_tail ? -x : x;
}

double incbi(double aa, double bb, double yy0) {
  if (yy0 <= 0) return 0.0;
  if (yy0 >= 1.0) return 1.0;
  double x0 = 0.0, yl = 0.0, x1 = 1.0, yh = 1.0;   // the bracket and incbet at its ends
  double a, b, y0, x = 0, y = 0, tolerance;
  bool mirrored = false, newton_used = false;
  enum { HALVE, NEWTON, DONE } state;
  auto mirror = [&](bool on) {
    mirrored = on;
    if (on) { a = bb; b = aa; y0 = 1.0 - yy0; } else { a = aa; b = bb; y0 = yy0; }
  };
  if (aa <= 1.0 || bb <= 1.0) {
    tolerance = 1.0e-6;
    mirror(false);
    x = a / (a + b);
    y = incbet(a, b, x);
    state = HALVE;
  } else {
    tolerance = 1.0e-4;
    double yp = -ndtri(yy0);   // a normal approximation of the answer
    if (yy0 > 0.5) { mirror(true); yp = -yp; } else { mirror(false); }
    const double lgm = (yp * yp - 3.0) / 6.0;
    x = 2.0 / (1.0 / (2.0 * a - 1.0) + 1.0 / (2.0 * b - 1.0));
    double d = yp * sqrt(x + lgm) / x - (1.0 / (2.0 * b - 1.0) - 1.0 / (2.0 * a - 1.0)) * (lgm + 5.0 / 6.0 - 2.0 / (3.0 * x));
    d = 2.0 * d;
    if (d < MINLOG) {
      x = 0.0;
      state = DONE;
    } else {
      x = a / (a + b * exp(d));
      y = incbet(a, b, x);
      yp = (y - y0) / y0;
      state = fabs(yp) < 0.2 ? NEWTON : HALVE;
    }
  }
  while (state != DONE) {
    if (state == HALVE) {
      int dir = 0;          // how many steps in a row went the same way
      double di = 0.5;      // where the bracket is split
      int i = 0;
      for (; i < 100; ++i) {
        if (i != 0) {
          x = x0 + di * (x1 - x0);
          if (x == 1.0) x = 1.0 - MACHEP;
          if (x == 0.0) {
            di = 0.5;
            x = x0 + di * (x1 - x0);
            if (x == 0.0) { state = DONE; break; }
          }
          y = incbet(a, b, x);
          double yp = (x1 - x0) / (x1 + x0);
          if (fabs(yp) < tolerance) { state = NEWTON; break; }
          yp = (y - y0) / y0;
          if (fabs(yp) < tolerance) { state = NEWTON; break; }
        }
        if (y < y0) {
          x0 = x;
          yl = y;
          if (dir < 0) { dir = 0; di = 0.5; }
          else if (dir > 3) di = 1.0 - (1.0 - di) * (1.0 - di);
          else if (dir > 1) di = 0.5 * di + 0.5;
          else di = (y0 - y) / (yh - yl);
          dir += 1;
          if (x0 > 0.75) {   // closing in on 1: continue on the mirrored problem, from a fresh bracket
            mirror(!mirrored);
            x = 1.0 - x;
            y = incbet(a, b, x);
            x0 = 0.0; yl = 0.0; x1 = 1.0; yh = 1.0;
            dir = 0; di = 0.5; i = -1;
            continue;
          }
        } else {
          x1 = x;
          if (mirrored && x1 < MACHEP) { x = 0.0; state = DONE; break; }
          yh = y;
          if (dir > 0) { dir = 0; di = 0.5; }
          else if (dir < -3) di = di * di;
          else if (dir < -1) di = 0.5 * di;
          else di = (y - y0) / (yh - yl);
          dir -= 1;
        }
      }
      if (i == 100) {   // the bracket did not close
        if (x0 >= 1.0) { x = 1.0 - MACHEP; state = DONE; }
        else if (x <= 0.0) { x = 0.0; state = DONE; }
        else state = NEWTON;
      }
    } else {   // NEWTON
      if (newton_used) { state = DONE; break; }
      newton_used = true;
      const double lgm = lgam(a + b) - lgam(a) - lgam(b);
      state = HALVE;   // unless a step below is small enough
      for (int i = 0; i < 8; ++i) {
        if (i != 0) y = incbet(a, b, x);
        if (y < yl) { x = x0; y = yl; }
        else if (y > yh) { x = x1; y = yh; }
        else if (y < y0) { x0 = x; yl = y; }
        else { x1 = x; yh = y; }
        if (x == 1.0 || x == 0.0) break;
        double d = (a - 1.0) * log(x) + (b - 1.0) * log(1.0 - x) + lgm;   // log of the density at x
        if (d < MINLOG) { state = DONE; break; }
        if (d > MAXLOG) break;
        d = exp(d);
        d = (y - y0) / d;
        double xt = x - d;
        if (xt <= x0) {
          y = (x - x0) / (x1 - x0);
          xt = x0 + 0.5 * y * (x - x0);
          if (xt <= 0.0) break;
        }
        if (xt >= x1) {
          y = (x1 - x) / (x1 - x0);
          xt = x1 - 0.5 * y * (x1 - x);
          if (xt >= 1.0) break;
        }
        x = xt;
        if (fabs(d / x) < 128.0 * MACHEP) { state = DONE; break; }
      }
      if (state == HALVE) tolerance = 256.0 * MACHEP;
    }
  }
  if (mirrored) x = x <= MACHEP ? 1.0 - MACHEP : 1.0 - x;
  return x;
}
}  // namespace cephes

// Nelder-Mead over two parameters with the reference's coefficients, start simplex, ordering and stopping rule
// (stats.cpp:2289-2392; max 1000 iterations, spread of the objective below 1e-8)
struct Vertex { double x[2]; double f; };
template <class F>
bool simplex_minimize(F&& objective, const double start[2], double best[2]) {
  Vertex v[3];
  for (int i = 0; i < 3; ++i) { v[i].x[0] = start[0]; v[i].x[1] = start[1]; }
  for (int i = 0; i < 2; ++i) v[i + 1].x[i] += (start[i] != 0.0) ? 0.05 * fabs(start[i]) : 0.00025;
  for (int i = 0; i < 3; ++i) v[i].f = objective(v[i].x);
  bool converged = false;
  for (uint32_t iter = 0; iter < 1000; ++iter) {
    {  // ascending by value, through the same std::sort of an index permutation as the reference (ties fall the same way)
      size_t order[3] = {0, 1, 2};
      std::sort(order, order + 3, [&v](size_t l, size_t r) { return v[l].f < v[r].f; });
      const Vertex s[3] = {v[order[0]], v[order[1]], v[order[2]]};
      v[0] = s[0]; v[1] = s[1]; v[2] = s[2];
    }
    if ((v[2].f - v[0].f) < 1e-8) { converged = true; break; }
    double centroid[2] = {0.0, 0.0};
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) centroid[j] += v[i].x[j];
    for (int j = 0; j < 2; ++j) centroid[j] /= 2.0;
    Vertex reflected;
    for (int j = 0; j < 2; ++j) reflected.x[j] = centroid[j] + 1.0 * (centroid[j] - v[2].x[j]);
    reflected.f = objective(reflected.x);
    if (reflected.f < v[0].f) {
      Vertex expanded;
      for (int j = 0; j < 2; ++j) expanded.x[j] = centroid[j] + 2.0 * (reflected.x[j] - centroid[j]);
      expanded.f = objective(expanded.x);
      v[2] = expanded.f < reflected.f ? expanded : reflected;
    } else if (reflected.f < v[1].f) {
      v[2] = reflected;
    } else {
      Vertex contracted;
      for (int j = 0; j < 2; ++j) contracted.x[j] = centroid[j] + 0.5 * (v[2].x[j] - centroid[j]);
      contracted.f = objective(contracted.x);
      if (contracted.f < v[2].f) {
        v[2] = contracted;
      } else {
        for (int i = 1; i < 3; ++i) {
          for (int j = 0; j < 2; ++j) v[i].x[j] = v[0].x[j] + 0.5 * (v[i].x[j] - v[0].x[j]);
          v[i].f = objective(v[i].x);
        }
      }
    }
  }
  int b = 0;
  for (int i = 1; i < 3; ++i) if (v[i].f < v[b].f) b = i;
  best[0] = v[b].x[0]; best[1] = v[b].x[1];
  return converged && std::isfinite(v[b].f);
}

}  // namespace

// stats.cpp:2394-2414
double binomial_frequency_lower_bound(double k, double n, double alpha) {
  if (!(alpha > 0.0) || !(alpha < 1.0) || !(n > 0.0)) return 0.0;
  if (k < 0.0) k = 0.0;
  if (k > n) k = n;
  if (k <= 0.0) return 0.0;
  if (k >= n) return pow(alpha, 1.0 / k);
  return cephes::incbi(k, n - k + 1.0, alpha);
}

double binomial_frequency_upper_bound(double k, double n, double alpha) {
  if (!(alpha > 0.0) || !(alpha < 1.0) || !(n > 0.0)) return 1.0;
  if (k < 0.0) k = 0.0;
  if (k > n) k = n;
  if (k >= n) return 1.0;
  if (k <= 0.0) return 1.0 - pow(alpha, 1.0 / n);
  return cephes::incbi(k + 1.0, n - k, 1.0 - alpha);
}

double nbinom_cdf(double k, double size, double mu) {
  return cephes::incbet(size, k + 1.0, size / (size + mu));
}

uint32_t nbinom_quantile(double target_pr, double size, double mu) {
  if (target_pr <= 0.0) return 0;
  uint64_t hi = 1;
  while (nbinom_cdf((double)hi, size, mu) < target_pr) hi *= 2;
  uint64_t lo = hi / 2;
  while (lo < hi) {
    const uint64_t mid = lo + (hi - lo) / 2;
    if (nbinom_cdf((double)mid, size, mu) < target_pr) lo = mid + 1; else hi = mid;
  }
  return (uint32_t)lo;
}

void read_coverage_distribution(const std::string& path, std::vector<double>& n, uint32_t& N) {
  std::ifstream in(path.c_str());
  if (!in) throw std::runtime_error("Could not open coverage distribution file: " + path);
  std::string header;
  std::getline(in, header);
  std::vector<std::pair<uint32_t, double>> rows;
  uint32_t coverage;
  double count;
  while (in >> coverage >> count) rows.emplace_back(coverage, count);
  N = 0;
  for (const auto& r : rows) N = std::max(N, r.first);
  n.assign((size_t)N + 1, 0.0);
  for (const auto& r : rows) n[r.first] = r.second;
}

CoverageFit fit_coverage_distribution(const std::vector<double>& n, uint32_t N, double pr_cutoff,
                                      const std::function<void(size_t, const std::function<void(size_t)>&)>* parallel_for) {
  CoverageFit out;
  // ---- moments, the peak of the smoothed histogram and the window around it (coverage_distribution.cpp:119-169)
  double positions = 0;
  for (uint32_t i = 1; i <= N; ++i) positions += n[i];
  bool have_window = positions != 0;
  double mean = 0, var = 0;
  uint32_t w_lo = 0, w_hi = 0;
  if (have_window) {
    for (uint32_t i = 1; i <= N; ++i) mean += i * n[i];
    mean /= positions;
    if (positions > 1) {
      for (uint32_t i = 1; i <= N; ++i) var += n[i] * (i - mean) * (i - mean);
      var /= (positions - 1);
    }
    out.average = mean; out.variance = var; out.relative_variance = mean > 0 ? var / mean : 0.0;
    const uint32_t from = std::max<uint32_t>((uint32_t)(mean / 4.0), 1u);
    double top = 0;
    uint32_t peak = 0;
    for (uint32_t i = from; i <= N; ++i) {
      double smooth;
      if (N >= 5) {  // five-point centred mean, undefined near either end
        if (i < 3 || i + 2 > N) continue;
        smooth = (n[i - 2] + n[i - 1] + n[i] + n[i + 1] + n[i + 2]) / 5.0;
      } else {
        smooth = n[i];
      }
      if (smooth > top) { top = smooth; peak = i; }
    }
    w_lo = std::max<uint32_t>((uint32_t)floor(peak * 0.5), 1u);
    w_hi = std::min<uint32_t>((uint32_t)ceil(peak * 1.5), N);
    have_window = w_lo != w_hi;
  }

  double fit_mu = 0, fit_size = 0;
  if (have_window) {
    out.censor_start = w_lo; out.censor_end = w_hi;
    // ---- coarse bins when the window is wider than ~2000 depths (:171-199)
    uint32_t per_bin = (w_hi - w_lo) / 1000;
    std::vector<double> x;
    uint32_t lo, hi;
    if (per_bin > 1) {
      lo = w_lo / per_bin;
      hi = (uint32_t)ceil((double)w_hi / per_bin);
      x.assign((size_t)hi + 1, 0.0);
      for (uint32_t i = lo; i <= hi; ++i)
        for (uint32_t j = 1; j <= per_bin; ++j) { const uint32_t at = i * per_bin + j; if (at <= N) x[i] += n[at]; }
    } else {
      per_bin = 1; lo = w_lo; hi = w_hi;
      x.assign((size_t)hi + 1, 0.0);
      for (uint32_t i = 1; i <= hi; ++i) x[i] = n[i];
    }
    double inside = 0;
    for (uint32_t i = lo; i <= hi; ++i) inside += x[i];
    double num = 0, den = 0;
    for (uint32_t i = 1; i <= hi; ++i) { num += i * x[i]; den += x[i]; }
    const double mean_estimate = num / den;

    // ---- objective: squared differences of proportions over the window, parameters in log space (:224-244)
    auto objective = [&](const double par[2]) -> double {
      const double mu = exp(par[0]), size = exp(par[1]);
      if (!std::isfinite(mu) || !std::isfinite(size)) return 1e10;
      std::vector<double> pmf((size_t)hi + 1, 0.0);
      double total = 0;
      for (uint32_t i = lo; i <= hi; ++i) { pmf[i] = nbinom_pmf((double)i, size, mu); total += pmf[i]; }
      if (!(total > 0) || !std::isfinite(total)) return 1e10;
      double l = 0;
      for (uint32_t i = lo; i <= hi; ++i) { const double diff = (x[i] / inside) - (pmf[i] / total); l += diff * diff; }
      return std::isfinite(l) ? l : 1e10;
    };

    // ---- restarts: six starting means x eight starting sizes; the lowest converged objective wins, the first on ties (:256-291)
    const double means[6] = {mean_estimate, (double)hi, (double)lo, 1.0 * (hi + lo) / 4.0, 2.0 * (hi + lo) / 4.0, 3.0 * (hi + lo) / 4.0};
    struct Restart { bool ok; double f, par[2]; };
    std::vector<Restart> runs(48);
    auto one = [&](size_t r) {
      double size0 = 100000;
      for (size_t k = 0; k <= r % 8; ++k) size0 /= 10.0;   // (the reference divides step by step, too)
      const double start[2] = {log(means[r / 8]), log(size0)};
      Restart& R = runs[r];
      R.ok = simplex_minimize(objective, start, R.par);
      R.f = R.ok ? objective(R.par) : 0.0;
    };
    if (parallel_for) (*parallel_for)(runs.size(), one);
    else for (size_t r = 0; r < runs.size(); ++r) one(r);
    double best = HUGE_VAL;
    for (const Restart& R : runs)
      if (R.ok && R.f < best) { best = R.f; fit_mu = exp(R.par[0]); fit_size = exp(R.par[1]); }
    if (!(best < HUGE_VAL)) { fit_mu = 0; fit_size = 0; }

    // ---- a fit that puts under 1 % of its mass inside the window is no fit (:299-316)
    double inside_fraction = 0;
    if (fit_mu > 0) {
      inside_fraction = nbinom_cdf((double)hi, fit_size, fit_mu) - nbinom_cdf((double)lo, fit_size, fit_mu);
      if (inside_fraction >= 0.01 && per_bin > 1) fit_mu = fit_mu * per_bin;
    }
    if (inside_fraction < 0.01) { fit_mu = 0; fit_size = 0; }
  }
  out.nb_mu = fit_mu; out.nb_size = fit_size;

  // ---- the coverage below which a deletion propagates (:371-398)
  double cutoff;
  if (fit_mu > 0) {
    cutoff = (double)nbinom_quantile(pr_cutoff, fit_size, fit_mu);
  } else {
    const double size_estimate = (1.0 / (out.variance - out.average)) * out.average * out.average;  // variance = mu + mu^2 / size
    if (size_estimate > 0 && std::isfinite(size_estimate) && out.average > 0) cutoff = (double)nbinom_quantile(pr_cutoff, size_estimate, out.average);
    else cutoff = -1;
    if (!(cutoff >= 1)) cutoff = out.average * 0.1;
  }
  if (cutoff < 1) cutoff = 1;                                   // one read does not make a region present
  if (fit_mu <= 3 && out.average <= 3) cutoff = -1;             // the reference sequence itself is missing
  out.deletion_coverage_propagation_cutoff = cutoff;
  return out;
}

}  // namespace brq
