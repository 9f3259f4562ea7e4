// Test-side check of csrc/canonical.h: text_canonical_6g(x) against the reference's text round trip
// (snprintf "%.6g" + strtod, error_count.cpp:629-690) on random values and on half-way cases.
// Usage: canonical_check <n_random>   prints "ok <count>" or the first mismatches and exits 1.
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include "canonical.h"

static uint64_t s = 0x9E3779B97F4A7C15ull;
static uint64_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }

static int bad = 0;
static uint64_t checked = 0;
static void check(double x) {
  if (fabs(x) >= 999999.5 || (x != 0.0 && fabs(x) < 1e-17)) return;  // outside the supported range (the function says so through `ok`)
  char buf[64];
  snprintf(buf, sizeof buf, "%.6g", x);
  const double want = strtod(buf, nullptr);
  bool ok = true;
  const double got = brq::text_canonical_6g(x, &ok);
  ++checked;
  if (!ok || memcmp(&want, &got, 8) != 0) {
    if (bad < 10) printf("mismatch: x=%.17g text=%s want=%.17g got=%.17g ok=%d\n", x, buf, want, got, (int)ok);
    ++bad;
  }
}

int main(int argc, char** argv) {
  const uint64_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1000000;
  // log-uniform magnitudes over the table's range and beyond, both signs
  for (uint64_t i = 0; i < n; ++i) {
    const double u = (double)(next() >> 11) / 9007199254740992.0, v = (double)(next() >> 11) / 9007199254740992.0;
    const double x = pow(10.0, -16.0 + 21.5 * u) * (1.0 + 9.0 * v) / 10.0;
    check((next() & 1) ? x : -x);
  }
  // what the table holds: log10(c + 1) - log10(S + 5)
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t S = next() % (1ull << (8 + next() % 34)), c = S ? next() % (S + 1) : 0;
    check(log10((double)c + 1.0) - log10((double)S + 5.0));
  }
  // half-way cases: the doubles around (D + 1/2) * 10^-k, and around exact six-digit decimals and powers of ten
  for (int k = 0; k <= 22; ++k) {
    for (uint64_t i = 0; i < n / 20 + 1000; ++i) {
      const uint64_t D = 100000 + next() % 900000;
      for (double frac : {0.5, 0.0}) {
        double x = ((double)D + frac) / brq::pow10_exact(k);
        for (int step = -3; step <= 3; ++step) {
          double y = x;
          for (int j = 0; j < (step < 0 ? -step : step); ++j) y = nextafter(y, step < 0 ? 0.0 : 1e300);
          check(y); check(-y);
        }
      }
    }
    for (int step = -4; step <= 4; ++step) {
      double y = 1.0 / brq::pow10_exact(k);
      for (int j = 0; j < (step < 0 ? -step : step); ++j) y = nextafter(y, step < 0 ? 0.0 : 1e300);
      if (y >= 1e-17) { check(y); check(-y); }
      y = 999999.5 / brq::pow10_exact(k);
      for (int j = 0; j < (step < 0 ? -step : step); ++j) y = nextafter(y, step < 0 ? 0.0 : 1e300);
      if (y >= 1e-17) { check(y); check(-y); }
    }
  }
  check(0.0);
  if (bad) { printf("FAILED %d of %" PRIu64 "\n", bad, checked); return 1; }
  printf("ok %" PRIu64 "\n", checked);
  return 0;
}
