// Small statistical functions shared by the host finalisation and the coverage fit.
#pragma once
#include <cmath>

namespace brq {

// log Gamma for x > 0: Cephes lgam as the reference carries it (stats.cpp:534-650).
inline double log_gamma(double x) {
  if (x < 13.0) {
    double z = 1.0, shift = 0.0, u = x;
    while (u >= 3.0) { shift -= 1.0; u = x + shift; z *= u; }
    while (u < 2.0) { z /= u; shift += 1.0; u = x + shift; }
    if (z < 0) z = -z;
    if (u == 2.0) return log(z);
    shift -= 2.0;
    const double y = x + shift;
    static const double num[6] = {-1378.25152569120859100, -38801.6315134637840924, -331612.992738871184744,
                                  -1162370.97492762307383, -1721737.00820839662146, -853555.664245765465627};
    static const double den[7] = {1.0, -351.815701436523470549, -17064.2106651881159223, -220528.590553854454839,
                                  -1139334.44367982507207, -2532523.07177582951285, -2018891.41433532773231};
    double b = num[0], c = den[0];
    for (int i = 1; i < 6; ++i) b = num[i] + y * b;
    for (int i = 1; i < 7; ++i) c = den[i] + y * c;
    return log(z) + y * b / c;
  }
  double q = (x - 0.5) * log(x) - x + 0.91893853320467274178;
  if (x > 100000000) return q;
  const double p = 1 / (x * x);
  if (x >= 1000.0) return q + ((7.9365079365079365079365 * 0.0001 * p - 2.7777777777777777777778 * 0.001) * p + 0.0833333333333333333333) / x;
  double a = 8.11614167470508450300 * 0.0001;
  a = -5.95061904284301438324 * 0.0001 + p * a;
  a = 7.93650340457716943945 * 0.0001 + p * a;
  a = -2.77777777730099687205 * 0.001 + p * a;
  a = 8.33333333333331927722 * 0.01 + p * a;
  return q + a / x;
}

}  // namespace brq
