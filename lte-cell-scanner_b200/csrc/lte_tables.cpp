// lte_tables.cpp - host-side LTE constant tables and small numeric helpers of the product path.
// (Independent of oracle/: the oracle re-derives the same quantities loop-for-loop from the
// reference; here they are written for table generation.)
#include <cmath>

#include "lcs_internal.hpp"

namespace lcs {

static const double kPi = 3.14159265358979323846;

// Zadoff-Chu roots 25/29/34, d_u(n)=exp(-j*pi*u*n(n+1)/63), DC element (n=31) removed.
// Reference: src/lte_lib.cpp:155-161.
void pss_fd(int n_id_2, cd out[62]) {
  static const int root[3] = {25, 29, 34};
  int o = 0;
  for (int n = 0; n < 63; n++) {
    if (n == 31) continue;
    const double ph = -kPi * root[n_id_2] * (double)(n * (n + 1)) / 63.0;
    out[o++] = cd(std::cos(ph), std::sin(ph));
  }
}

// 128-point time-domain PSS with a 9-sample cyclic prefix (137 taps), scaled so that
// sigpower(td)==sigpower of the 62 occupied bins spread over 128.  Reference: src/lte_lib.cpp:177-188
// (idft(...)*sqrt(128/62), idft = ifft*sqrt(N)).  Direct O(N^2) inverse DFT in double.
static void pss_td_compute(int n_id_2, cd out[137]) {
  cd fd[62], X[128];
  pss_fd(n_id_2, fd);
  for (int i = 0; i < 128; i++) X[i] = 0;
  for (int i = 0; i < 31; i++) {
    X[1 + i] = fd[31 + i];   // positive frequencies 1..31
    X[97 + i] = fd[i];       // negative frequencies -31..-1
  }
  cd td[128];
  const double sc = std::sqrt(128.0) * std::sqrt(128.0 / 62.0) / 128.0;
  for (int n = 0; n < 128; n++) {
    cd s = 0;
    for (int k = 0; k < 128; k++) {
      if (X[k] == cd(0, 0)) continue;
      const int r = (n * k) & 127;  // exact argument reduction
      const double ph = 2 * kPi * r / 128.0;
      s += X[k] * cd(std::cos(ph), std::sin(ph));
    }
    td[n] = s * sc;
  }
  for (int i = 0; i < 9; i++) out[i] = td[119 + i];
  for (int i = 0; i < 128; i++) out[9 + i] = td[i];
}
// The three sequences are constants: computed once (a plan is built per centre frequency of a sweep).
void pss_td(int n_id_2, cd out[137]) {
  struct Tab { cd v[3][137]; Tab() { for (int t = 0; t < 3; t++) pss_td_compute(t, v[t]); } };
  static const Tab tab;
  for (int i = 0; i < 137; i++) out[i] = tab.v[n_id_2][i];
}

// SSS in the frequency domain as +-1 integers.  Reference: src/lte_lib.cpp:199-257 (3GPP 36.211 6.11.2).
void sss_fd(int n_id_1, int n_id_2, int slot, int out[62]) {
  struct Tab {      // the three m-sequences, built once (thread-safe static initialisation)
    int s[31], c[31], z[31];
    Tab() {
      int x[31];
      auto gen = [&](int* dst, auto rec) {
        for (int i = 0; i < 5; i++) x[i] = (i == 4);
        for (int i = 0; i < 26; i++) x[i + 5] = rec(x, i) & 1;
        for (int i = 0; i < 31; i++) dst[i] = 1 - 2 * x[i];
      };
      gen(s, [](const int* v, int i) { return v[i + 2] + v[i]; });                       // x^5+x^2+1
      gen(c, [](const int* v, int i) { return v[i + 3] + v[i]; });                       // x^5+x^3+1
      gen(z, [](const int* v, int i) { return v[i + 4] + v[i + 2] + v[i + 1] + v[i]; });  // x^5+x^4+x^2+x+1
    }
  };
  static const Tab tab;
  const int *s_t = tab.s, *c_t = tab.c, *z_t = tab.z;
  const int qp = n_id_1 / 30;
  const int q = (n_id_1 + qp * (qp + 1) / 2) / 30;
  const int mp = n_id_1 + q * (q + 1) / 2;
  const int m0 = mp % 31, m1 = (m0 + mp / 31 + 1) % 31;
  for (int n = 0; n < 31; n++) {
    const int s0 = s_t[(n + m0) % 31], s1 = s_t[(n + m1) % 31];
    const int c0 = c_t[(n + n_id_2) % 31], c1 = c_t[(n + n_id_2 + 3) % 31];
    const int z10 = z_t[(n + (m0 % 8)) % 31], z11 = z_t[(n + (m1 % 8)) % 31];
    if (slot == 0) { out[2 * n] = s0 * c0; out[2 * n + 1] = s1 * c1 * z10; }
    else { out[2 * n] = s1 * c0; out[2 * n + 1] = s0 * c1 * z11; }
  }
}

// Length-31 Gold sequence of 36.211 7.2 (Nc=1600), 32-bit shift registers.  Reference: src/lte_lib.cpp:41-147.
std::vector<uint8_t> lte_pn(uint32_t c_init, uint32_t len) {
  uint32_t x1 = 1, x2 = c_init & 0x7fffffffu;
  auto clk = [&]() {
    const uint32_t n1 = ((x1 >> 3) ^ x1) & 1u;
    const uint32_t n2 = ((x2 >> 3) ^ (x2 >> 2) ^ (x2 >> 1) ^ x2) & 1u;
    x1 = (x1 >> 1) | (n1 << 30);
    x2 = (x2 >> 1) | (n2 << 30);
  };
  for (int i = 0; i < 1600; i++) clk();
  std::vector<uint8_t> c(len);
  for (uint32_t i = 0; i < len; i++) {
    c[i] = (uint8_t)((x1 ^ x2) & 1u);
    clk();
  }
  return c;
}

// chi2cdf_inv(p,k) = 2*gamma_p_inv(k/2,p)  (include/dsp.h:188-193).  Newton iteration on the
// regularised incomplete gamma function evaluated by Lentz's continued fraction / power series,
// working on whichever tail is small.
static double reg_gamma_q(double a, double x) {  // Q(a,x)
  if (x <= 0) return 1.0;
  const double lg = std::lgamma(a);
  if (x < a + 1) {
    double term = 1.0 / a, sum = term;
    for (int n = 1; n < 10000; n++) {
      term *= x / (a + n);
      sum += term;
      if (term < sum * 1e-17) break;
    }
    return 1.0 - sum * std::exp(a * std::log(x) - x - lg);
  }
  const double tiny = 1e-300;
  double b = x + 1 - a, c = 1 / tiny, d = 1 / b, h = d;
  for (int i = 1; i < 10000; i++) {
    const double an = -i * (i - a);
    b += 2;
    d = an * d + b; if (std::fabs(d) < tiny) d = tiny;
    c = b + an / c; if (std::fabs(c) < tiny) c = tiny;
    d = 1 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1) < 1e-16) break;
  }
  return h * std::exp(a * std::log(x) - x - lg);
}
double chi2cdf_inv(double p, double k) {
  const double a = k / 2, q = 1 - p;
  // Wilson-Hilferty start, then bisection-safeguarded Newton on log Q (upper tail) or log P.
  const bool upper = p > 0.5;
  double lo = 0, hi = a + 10 * std::sqrt(a) + 50;
  while ((upper ? reg_gamma_q(a, hi) > q : 1 - reg_gamma_q(a, hi) < p)) hi *= 2;
  double x = 0.5 * (lo + hi);
  for (int it = 0; it < 300; it++) {
    const double Q = reg_gamma_q(a, x);
    const bool too_low = upper ? (Q > q) : (1 - Q < p);
    if (too_low) lo = x; else hi = x;
    // Newton step using dQ/dx = -x^(a-1) e^-x / Gamma(a)
    const double dens = std::exp((a - 1) * std::log(x) - x - std::lgamma(a));
    double xn = upper ? x + (Q - q) / dens : x - ((1 - Q) - p) / dens;
    if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
    if (std::fabs(xn - x) <= 1e-15 * x) { x = xn; break; }
    x = xn;
  }
  return 2 * x;
}

}  // namespace lcs
