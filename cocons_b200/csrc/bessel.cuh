// Modified Bessel function K_nu(x) and the Matern correlation factor for the
// pairwise-assembly kernel.  Replaces boost::math::cyl_bessel_k at
// src/cocons_full.cpp:294,450,573 (reference) with a double-precision device
// routine from the same algorithm family (Temme 1975 series, Steed/Thompson-
// Barnett CF2, Hankel asymptotic tail), reorganised for SIMT execution:
//   * the branch taken is decided on a warp vote, so that a warp whose lanes sit
//     in the same band runs exactly one of the three bodies;
//   * everything is expressed through e^x K_nu(x) so that the e^-Q factor is
//     applied once, with an exactly representable argument;
//   * 1/Gamma(nu) comes out of the same gamma1/gamma2 polynomials the Temme
//     series needs, so no tgamma() call is made.
// The file also compiles as plain C++ (no CUDA) so tests can check it on the
// CPU against mpmath; that build is test scaffolding, never a product path.
#ifndef COCONS_BESSEL_CUH
#define COCONS_BESSEL_CUH

#include <math.h>

#ifdef __CUDACC__
#define COCONS_HD __host__ __device__ __forceinline__
#else
#define COCONS_HD inline
#endif

#include "gamma_coeffs.inc"

namespace cocons {

constexpr double kPi = 3.141592653589793238462643383279502884;
constexpr double kHalfPi = 1.570796326794896619231321691639751442;
constexpr double kBesselEps = 1.0e-16;  // series / continued-fraction stopping level
constexpr double kHankelX = 25.0;       // Hankel tail used for x >= kHankelX and nu <= kHankelNuMax
constexpr double kHankelNuMax = 3.0;    // (14 terms reach 8.8e-16 there)
constexpr int kHankelTerms = 14;

// gamma1, gamma2, 1/Gamma(1+mu), 1/Gamma(1-mu) for |mu| <= 1/2
struct TemmeGammas {
  double g1, g2, rgp, rgm;
};

COCONS_HD TemmeGammas temme_gammas(double mu) {
  const double c1[] = COCONS_GAMMA1_COEFFS;
  const double c2[] = COCONS_GAMMA2_COEFFS;
  const double t = fma(8.0 * mu, mu, -1.0);
  double g1 = c1[9], g2 = c2[9];
#pragma unroll
  for (int k = 8; k >= 0; --k) {
    g1 = fma(g1, t, c1[k]);
    g2 = fma(g2, t, c2[k]);
  }
  TemmeGammas r;
  r.g1 = g1;
  r.g2 = g2;
  r.rgp = fma(-mu, g1, g2);  // 1/Gamma(1+mu) = gamma2 - mu gamma1
  r.rgm = fma(mu, g1, g2);   // 1/Gamma(1-mu) = gamma2 + mu gamma1
  return r;
}

// 1/Gamma(nu) for nu = nl + mu, nu > 0, from 1/Gamma(1+mu)
COCONS_HD double rgamma_from(double rgp, double mu, int nl) {
  if (nl == 0) return rgp * mu;  // Gamma(mu) = Gamma(1+mu)/mu
  double prod = 1.0;
  for (int k = 1; k < nl; ++k) prod *= (mu + (double)k);
  return rgp / prod;
}

// K_mu(x), K_{mu+1}(x), UNSCALED, for 0 < x <= 2, |mu| <= 1/2  (Temme's series)
COCONS_HD void bessel_k_temme(double mu, double x, const TemmeGammas& G, double& kmu, double& kmu1) {
  const double mu2 = mu * mu;
  const double pimu = kPi * mu;
  const double fact = (fabs(pimu) < 1e-8) ? 1.0 : pimu / sin(pimu);
  const double d0 = -log(0.5 * x);
  const double e0 = mu * d0;
  const double e2 = e0 * e0;
  const double shc = (fabs(e0) < 1e-3) ? fma(e2, fma(e2, 1.0 / 120.0, 1.0 / 6.0), 1.0) : sinh(e0) / e0;
  double ff = fact * fma(G.g1, cosh(e0), G.g2 * shc * d0);
  double sum = ff;
  const double ee = exp(e0);
  double p = 0.5 * ee / G.rgp;
  double q = 0.5 / (ee * G.rgm);
  double c = 1.0;
  const double d = 0.25 * x * x;
  double sum1 = p;
  for (int i = 1; i < 64; ++i) {
    const double fi = (double)i;
    ff = (fma(fi, ff, p) + q) / fma(fi, fi, -mu2);
    c *= d / fi;
    p /= (fi - mu);
    q /= (fi + mu);
    const double del = c * ff;
    sum += del;
    sum1 = fma(c, fma(-fi, ff, p), sum1);
    if (fabs(del) < fabs(sum) * kBesselEps) break;
  }
  kmu = sum;
  kmu1 = sum1 * (2.0 / x);
}

// e^x K_mu(x), e^x K_{mu+1}(x) for x > 2, |mu| <= 1/2  (Steed's algorithm for CF2)
COCONS_HD void bessel_k_cf2_scaled(double mu, double x, double& kmu, double& kmu1) {
  const double mu2 = mu * mu;
  double b = 2.0 * (1.0 + x);
  double d = 1.0 / b;
  double h = d, delh = d;
  double q1 = 0.0, q2 = 1.0;
  const double a1 = 0.25 - mu2;
  double q = a1, c = a1;
  double a = -a1;
  double s = fma(q, delh, 1.0);
  for (int i = 2; i < 512; ++i) {
    a -= (double)(2 * (i - 1));
    c = -a * c / (double)i;
    const double qnew = fma(-b, q2, q1) / a;
    q1 = q2;
    q2 = qnew;
    q = fma(c, qnew, q);
    b += 2.0;
    d = 1.0 / fma(a, d, b);
    delh = fma(b, d, -1.0) * delh;
    h += delh;
    const double dels = q * delh;
    s += dels;
    if (fabs(dels) < fabs(s) * kBesselEps) break;
  }
  h = a1 * h;
  kmu = sqrt(kHalfPi / x) / s;
  kmu1 = kmu * (mu + x + 0.5 - h) / x;
}

// e^x K_nu(x) from the Hankel expansion, x >= kHankelX, 0 < nu <= kHankelNuMax
COCONS_HD double bessel_k_hankel_scaled(double nu, double x) {
  const double four_nu2 = 4.0 * nu * nu;
  const double r8x = 1.0 / (8.0 * x);
  double term = 1.0, sum = 1.0;
#pragma unroll
  for (int k = 1; k <= kHankelTerms; ++k) {
    const double odd = (double)(2 * k - 1);
    term *= (four_nu2 - odd * odd) * r8x * (1.0 / (double)k);
    sum += term;
  }
  return sqrt(kHalfPi / x) * sum;
}

// upward recurrence K_{mu+k+1} = K_{mu+k-1} + 2(mu+k)/x K_{mu+k}, nl steps
COCONS_HD double bessel_k_recur(double kmu, double kmu1, double mu, double x, int nl) {
  const double two_over_x = 2.0 / x;
  for (int i = 1; i <= nl; ++i) {
    const double nxt = fma((mu + (double)i) * two_over_x, kmu1, kmu);
    kmu = kmu1;
    kmu1 = nxt;
  }
  return kmu;
}

// which body a given (nu, x) belongs to: 0 Temme, 1 CF2, 2 Hankel
COCONS_HD int bessel_band(double nu, double x) {
  if (x <= 2.0) return 0;
  if (x >= kHankelX && nu <= kHankelNuMax) return 2;
  return 1;
}

// K_nu(x) itself (unscaled) - used by tests and by callers that want the plain value
COCONS_HD double bessel_k(double nu, double x) {
  const int nl = (int)(nu + 0.5);
  const double mu = nu - (double)nl;
  const int band = bessel_band(nu, x);
  if (band == 2) return bessel_k_hankel_scaled(nu, x) * exp(-x);
  double kmu, kmu1;
  if (band == 0) {
    const TemmeGammas G = temme_gammas(mu);
    bessel_k_temme(mu, x, G, kmu, kmu1);
    return bessel_k_recur(kmu, kmu1, mu, x, nl);
  }
  bessel_k_cf2_scaled(mu, x, kmu, kmu1);
  return bessel_k_recur(kmu, kmu1, mu, x, nl) * exp(-x);
}

// Matern correlation factor  2^{1-nu}/Gamma(nu) * Q^nu * K_nu(Q)   (eps < Q < 706),
// the quantity formed at src/cocons_full.cpp:293-294.
COCONS_HD double matern_corr(double nu, double Q) {
  const int nl = (int)(nu + 0.5);
  const double mu = nu - (double)nl;
  const TemmeGammas G = temme_gammas(mu);
  const double two_rgamma = 2.0 * rgamma_from(G.rgp, mu, nl);
  const double powfac = exp(nu * log(0.5 * Q));  // (Q/2)^nu
  const int band = bessel_band(nu, Q);
  if (band == 0) {
    double kmu, kmu1;
    bessel_k_temme(mu, Q, G, kmu, kmu1);
    return two_rgamma * powfac * bessel_k_recur(kmu, kmu1, mu, Q, nl);
  }
  double ks;
  if (band == 2) {
    ks = bessel_k_hankel_scaled(nu, Q);
  } else {
    double kmu, kmu1;
    bessel_k_cf2_scaled(mu, Q, kmu, kmu1);
    ks = bessel_k_recur(kmu, kmu1, mu, Q, nl);
  }
  return two_rgamma * powfac * ks * exp(-Q);
}

// The reference's own tail formula for Q >= 706 (src/cocons_full.cpp:299-305):
// leading Hankel term only.
COCONS_HD double matern_corr_tail(double nu, double Q) {
  const int nl = (int)(nu + 0.5);
  const double mu = nu - (double)nl;
  const TemmeGammas G = temme_gammas(mu);
  const double two_rgamma = 2.0 * rgamma_from(G.rgp, mu, nl);
  return two_rgamma * exp(nu * log(0.5 * Q)) * sqrt(kHalfPi / Q) * exp(-Q);
}

}  // namespace cocons

#endif
