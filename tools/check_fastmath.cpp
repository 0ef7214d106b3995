// Accuracy check of csrc/fastmath64.cuh on the host against long double libm.
//   g++ -O2 -std=c++17 -o check_fastmath check_fastmath.cpp && ./check_fastmath
#include <cmath>
#include <cstdio>
#include <random>
#include "../gnss_ins_sim_b200/csrc/fastmath64.cuh"

static double ulp_err(double got, long double want) {
  if (want == 0.0L) return std::fabs(got) == 0 ? 0 : 1e9;
  int e;
  std::frexp((double)want, &e);
  long double ulp = std::ldexp(1.0L, e - 53);
  return (double)(fabsl((long double)got - want) / ulp);
}

int main() {
  std::mt19937_64 rng(1);
  const long double PI = 3.141592653589793238462643383279502884L;
  double w_sin = 0, w_cos = 0, w_sinpi = 0, w_cospi = 0, w_log = 0, abs_sincos = 0;
  std::uniform_real_distribution<double> ua(-64.0, 64.0), ub(-3.2, 3.2), u01(0.0, 1.0);
  for (int i = 0; i < 4000000; ++i) {
    double x = (i & 1) ? ua(rng) : ub(rng);
    double s, c;
    b2ins::sincos_bounded(x, &s, &c);
    long double ws = sinl((long double)x), wc = cosl((long double)x);
    w_sin = std::fmax(w_sin, ulp_err(s, ws));
    w_cos = std::fmax(w_cos, ulp_err(c, wc));
    abs_sincos = std::fmax(abs_sincos, (double)fmaxl(fabsl(s - ws), fabsl(c - wc)));
    // Box-Muller angle: u = m * 2^-52
    uint64_t m = rng() >> 12;
    double u = (double)m * 0x1p-52;
    b2ins::sincospi_2u(2.0 * u, &s, &c);
    long double a = 2.0L * PI * (long double)u;
    w_sinpi = std::fmax(w_sinpi, (double)fabsl(s - sinl(a)));
    w_cospi = std::fmax(w_cospi, (double)fabsl(c - cosl(a)));
    double u1 = 1.0 - u;  // (0,1]
    w_log = std::fmax(w_log, ulp_err(b2ins::log_unit(u1), logl((long double)u1)));
    if (i < 64) {  // tiny arguments near the tail
      double tiny = std::ldexp(1.0 + u, -52 + (i % 52));
      if (tiny <= 1.0) w_log = std::fmax(w_log, ulp_err(b2ins::log_unit(tiny), logl((long double)tiny)));
    }
  }
  // exact points
  double s, c;
  b2ins::sincospi_2u(0.0, &s, &c);
  bool exact = (s == 0.0 && c == 1.0);
  b2ins::sincospi_2u(0.5, &s, &c);
  exact = exact && (s == 1.0 && std::fabs(c) < 1e-16);
  b2ins::sincospi_2u(1.0, &s, &c);
  exact = exact && (std::fabs(s) < 1e-15 && c == -1.0);
  exact = exact && (b2ins::log_unit(1.0) == 0.0);
  std::printf("{\"sin_ulp\": %.3f, \"cos_ulp\": %.3f, \"sincos_abs\": %.3e, \"sinpi_abs\": %.3e, "
              "\"cospi_abs\": %.3e, \"log_ulp\": %.3f, \"exact_points\": %s}\n",
              w_sin, w_cos, abs_sincos, w_sinpi, w_cospi, w_log, exact ? "true" : "false");
  return (w_sin < 2.0 && w_cos < 2.0 && w_log < 2.0 && w_sinpi < 3e-16 && w_cospi < 3e-16 && exact) ? 0 : 1;
}
