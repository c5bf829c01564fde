// TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of the reference's dense
// covariance assembly (and of the tapered one, src/cocons_taper.cpp) on plain arrays.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library;
// the product path (cocons_b200/) never does.
//
// Parity status: PINNED against outputs of the reference's own source compiled
// here (oracle/_ref/libcocons_ref.so, built by oracle/Makefile from
// /root/reference/src/cocons_full.cpp against the Rcpp/BH stand-ins in
// oracle/shim/) - tests/test_oracle.py demands bit equality with it - and
// against 50-digit mpmath evaluations of the mathematical formula.  NOT pinned
// by reference-published golden vectors: the reference's tests hold none for
// this path, and Boost's cyl_bessel_k is replaced by
// libstdc++'s std::cyl_bessel_k evaluated in long double (BH is unpinned and
// absent).
//
// Each function cites the reference lines whose arithmetic (operation order,
// fma placement, branch thresholds) it follows.
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <vector>

namespace {

// src/cocons_types.h:12-17 - sequential fma dot product x_i . b
inline double lin_pred(const double* X, long n, long p, long i, const double* b) {
  double acc = 0.0;
  for (long k = 0; k < p; ++k) acc = std::fma(X[k * n + i], b[k], acc);
  return acc;
}
// src/cocons_types.h:17 - the log link written as 1/exp(-eta)
inline double link_exp(double eta) { return 1 / std::exp(-1 * eta); }
// src/cocons_types.h:49-54 - a*b - c*d with the rounding error of c*d removed
inline double diff_of_products(double a, double b, double c, double d) {
  double cd = c * d;
  double err = std::fma(c, d, -cd);
  double res = std::fma(a, b, -cd);
  return res - err;
}
// Boost stand-in, see oracle/shim/boost/math/special_functions/bessel.hpp
inline double bessel_k(double nu, double x) {
  return (double)std::cyl_bessel_k((long double)nu, (long double)x);
}

struct Site {
  double tilt, r, a, D, sigma, nug, snu, var;
};

enum SmoothRule { kGeoMeanOfLimits = 0, kClassicExp = 1 };

// per-site stage: src/cocons_full.cpp:98-107 (cov_rns), :374-405 (pred), :517-527 (classic)
void site_stage(long n, long p, const double* X, const double* theta6, const double* limits, SmoothRule rule,
                bool fill_smooth, std::vector<Site>& out) {
  const double* sd = theta6;
  const double* scale = theta6 + p;
  const double* aniso = theta6 + 2 * p;
  const double* tilt = theta6 + 3 * p;
  const double* smooth = theta6 + 4 * p;
  const double* nugget = theta6 + 5 * p;
  std::vector<double> two_scale(p), root(p), half_sd(p);
  for (long k = 0; k < p; ++k) {
    double sj = (k == 0) ? 0.0 : scale[k];  // scale_je, :64
    two_scale[k] = 2 * sj;                  // `2 * scale_je`, :101
    root[k] = 2 * sj + aniso[k];            // sqrt_vector, :66
    half_sd[k] = 0.5 * sd[k];               // :104
  }
  out.resize(n);
  for (long i = 0; i < n; ++i) {
    Site s;
    s.tilt = M_PI / (1 + std::exp(-1 * lin_pred(X, n, p, i, tilt)));  // types.h:46
    s.r = link_exp(lin_pred(X, n, p, i, two_scale.data()));
    s.a = link_exp(lin_pred(X, n, p, i, aniso));
    s.D = link_exp(lin_pred(X, n, p, i, root.data()));
    s.sigma = link_exp(lin_pred(X, n, p, i, half_sd.data()));
    s.nug = link_exp(lin_pred(X, n, p, i, nugget));
    s.var = link_exp(lin_pred(X, n, p, i, sd));
    s.snu = 0.0;
    if (fill_smooth) {
      double eta = lin_pred(X, n, p, i, smooth);
      if (rule == kClassicExp)
        s.snu = link_exp(eta);  // :524
      else                      // types.h:27 then sqrt, :93 / :381
        s.snu = std::sqrt((limits[1] - limits[0]) / (1 + std::exp(-1 * eta)) + limits[0]);
    }
    out[i] = s;
  }
}

struct PairGeom {
  double det, Q;
};

// local-kernel averaging and scaled distance: src/cocons_full.cpp:260-281
inline PairGeom pair_geometry(const Site& si, const Site& sj, double dx, double dy, double nu, double global_range) {
  double s11 = (si.r + sj.r) * 0.5;
  double s22 = diff_of_products(si.r, si.a * si.a, -sj.r, sj.a * sj.a) * 0.5;
  double s12 = diff_of_products(si.r * si.a, std::cos(si.tilt), -1 * sj.r * sj.a, std::cos(sj.tilt)) * 0.5;
  double det = diff_of_products(s11, s22, s12, s12);
  double Q = std::sqrt(8 * nu / (global_range * det)) *
             std::sqrt(std::fma(diff_of_products(s22, dx * dx, -s11, dy * dy), 1, -2 * s12 * dx * dy));
  return PairGeom{det, Q};
}

// sigma_i sigma_j sqrt(D_i sin t_i D_j sin t_j)/sqrt(det) is applied left to right
// after the correlation factor, exactly as written at :293-297
inline double scaled(double corr, const Site& si, const Site& sj, double det) {
  return corr * si.sigma * sj.sigma * std::sqrt(si.D * std::sin(si.tilt) * sj.D * std::sin(sj.tilt)) /
         std::sqrt(det);
}

// general Matern branch: :291-307
inline double matern_general(double nu, double Q, const Site& si, const Site& sj, double det) {
  if (Q < 706.0)
    return scaled(std::pow(2.0, -(nu - 1)) / std::tgamma(nu) * std::pow(Q, nu) * bessel_k(nu, Q), si, sj, det);
  return scaled(std::pow(2.0, -(nu - 1)) / std::tgamma(nu) * std::pow(Q, nu) * std::sqrt(M_PI / (2.0 * Q)) *
                    std::exp(-Q),
                si, sj, det);
}

}  // namespace

extern "C" {

// cov_rns, src/cocons_full.cpp:40-321.  out: n x n column-major, full symmetric.
int oracle_cov_rns(long n, long p, const double* locs, const double* X, const double* theta6, const double* limits,
                   double* out) {
  const double eps = DBL_EPSILON;
  const double* smooth = theta6 + 4 * p;
  const double global_range = 1 / std::exp(-2 * theta6[p]);  // :62
  bool slopes_zero = true;                                   // types.h:56-63
  for (long k = 1; k < p; ++k)
    if (smooth[k] != 0) slopes_zero = false;
  int path = 0;  // 0 general, 1/2/3 closed forms
  double nu_fixed = 0.0;
  bool fixed = slopes_zero && (limits[0] == limits[1]);  // :85
  if (fixed) {
    nu_fixed = limits[0];
    if (std::fabs(nu_fixed - 0.5) < 1e-6) path = 1;  // types.h:65-70
    else if (std::fabs(nu_fixed - 1.5) < 1e-6) path = 2;
    else if (std::fabs(nu_fixed - 2.5) < 1e-6) path = 3;
  }
  std::vector<Site> S;
  // quirk: with fixed nu the per-site smooth vector stays 0
  site_stage(n, p, X, theta6, limits, kGeoMeanOfLimits, !fixed, S);
  for (size_t k = 0; k < (size_t)n * (size_t)n; ++k) out[k] = 0.0;
  for (long i = 0; i < n; ++i) out[(size_t)i * n + i] = S[i].var + S[i].nug;  // :110-112
  for (long i = 0; i < n; ++i) {
    for (long j = i + 1; j < n; ++j) {
      double dx = locs[i] - locs[j], dy = locs[n + i] - locs[n + j];
      double nu = path ? nu_fixed : S[i].snu * S[j].snu;  // :274
      PairGeom g = pair_geometry(S[i], S[j], dx, dy, nu, global_range);
      double v;
      if (g.Q <= eps) {
        v = S[i].var + S[i].nug;  // :284-286 (row-i value)
      } else if (path == 1) {
        v = scaled(std::exp(-g.Q), S[i], S[j], g.det);  // :150
      } else if (path == 2) {
        v = scaled((1 + g.Q) * std::exp(-g.Q), S[i], S[j], g.det);  // :196
      } else if (path == 3) {
        v = scaled((1 + g.Q + g.Q * g.Q / 3) * std::exp(-g.Q), S[i], S[j], g.det);  // :242
      } else {
        v = matern_general(nu, g.Q, S[i], S[j], g.det);
      }
      out[(size_t)j * n + i] = v;
      out[(size_t)i * n + j] = v;
    }
  }
  return 0;
}

// cov_rns_pred, src/cocons_full.cpp:334-471.  out: m x n column-major (pred sites are rows).
int oracle_cov_rns_pred(long n, long m, long p, const double* locs, const double* locs_pred, const double* X,
                        const double* X_pred, const double* theta6, const double* limits, double* out) {
  const double eps = DBL_EPSILON;
  const double global_range = 1 / std::exp(-2 * theta6[p]);  // :351
  std::vector<Site> S, P;
  site_stage(n, p, X, theta6, limits, kGeoMeanOfLimits, true, S);
  site_stage(m, p, X_pred, theta6, limits, kGeoMeanOfLimits, true, P);
  for (long i = 0; i < m; ++i) {
    for (long j = 0; j < n; ++j) {
      double v;
      if (locs_pred[i] == locs[j] && locs_pred[m + i] == locs[n + j]) {  // :410
        v = P[i].var + P[i].nug;
      } else {
        double dx = locs_pred[i] - locs[j], dy = locs_pred[m + i] - locs[n + j];
        double nu = P[i].snu * S[j].snu;  // :431
        PairGeom g = pair_geometry(P[i], S[j], dx, dy, nu, global_range);
        v = (g.Q <= eps) ? P[i].var + P[i].nug : matern_general(nu, g.Q, P[i], S[j], g.det);
      }
      out[(size_t)j * m + i] = v;
    }
  }
  return 0;
}

// cov_rns_classic, src/cocons_full.cpp:480-594.
int oracle_cov_rns_classic(long n, long p, const double* locs, const double* X, const double* theta6, double* out) {
  const double eps = DBL_EPSILON;
  const double global_range = 1 / std::exp(-2 * theta6[p]);  // :501
  std::vector<Site> S;
  site_stage(n, p, X, theta6, nullptr, kClassicExp, true, S);
  for (long i = 0; i < n; ++i) {
    out[(size_t)i * n + i] = S[i].var + S[i].nug;  // :532-536
    for (long j = i + 1; j < n; ++j) {
      double dx = locs[i] - locs[j], dy = locs[n + i] - locs[n + j];
      double nu = (S[i].snu + S[j].snu) / 2;  // :554
      PairGeom g = pair_geometry(S[i], S[j], dx, dy, nu, global_range);
      double v = (g.Q <= eps) ? S[i].var + S[i].nug : matern_general(nu, g.Q, S[i], S[j], g.det);
      out[(size_t)j * n + i] = v;
      out[(size_t)i * n + j] = v;
    }
  }
  return 0;
}

// ---- sparse (tapered) model: src/cocons_taper.cpp ---------------------------------------------
// Isotropic nonstationary Matern on the entries of a CSR pattern (spam layout, 1-based indices
// handed over as doubles, as Rcpp hands them to the reference).  No anisotropy / tilt here.
struct TaperSite {
  double range, sigma, snu;
};

// per-site stage: src/cocons_taper.cpp:54-70 (pred) and :195-209
static void taper_site_stage(long n, long p, const double* X, const double* theta6, const double* limits,
                             bool fill_smooth, std::vector<TaperSite>& out) {
  const double* sd = theta6;
  const double* scale = theta6 + p;
  const double* smooth = theta6 + 4 * p;
  std::vector<double> two_scale(p), half_sd(p);
  for (long k = 0; k < p; ++k) {
    two_scale[k] = 2 * scale[k];  // the intercept is kept (no global range in this file)
    half_sd[k] = 0.5 * sd[k];
  }
  out.resize(n);
  for (long i = 0; i < n; ++i) {
    TaperSite s;
    s.range = link_exp(lin_pred(X, n, p, i, two_scale.data()));
    s.sigma = link_exp(lin_pred(X, n, p, i, half_sd.data()));
    s.snu = 0.0;
    if (fill_smooth)
      s.snu = std::sqrt((limits[1] - limits[0]) / (1 + std::exp(-1 * lin_pred(X, n, p, i, smooth))) + limits[0]);
    out[i] = s;
  }
}

// one off-diagonal entry: src/cocons_taper.cpp:96-128 / :384-417 (path 0) and the closed forms :233-343.
// Returns false when Q <= eps (the caller stores the row site's variance + nugget).
static bool taper_pair(const TaperSite& a, const TaperSite& b, double dx, double dy, int path, double nu_fixed,
                       double* v) {
  const double nu = path ? nu_fixed : a.snu * b.snu;
  const double prefactor = (2 * std::pow(a.range, 0.5) * std::pow(b.range, 0.5)) / (a.range + b.range);
  const double avg = (a.range + b.range) / 2;
  const double Q = std::sqrt(8 * nu) * std::sqrt(std::pow(dx, 2) + std::pow(dy, 2)) / std::sqrt(avg);
  if (Q <= DBL_EPSILON) return false;
  if (path == 1)
    *v = prefactor * std::exp(-Q) * a.sigma * b.sigma;
  else if (path == 2)
    *v = prefactor * (1 + Q) * std::exp(-Q) * a.sigma * b.sigma;
  else if (path == 3)
    *v = prefactor * (1 + Q + Q * Q / 3) * std::exp(-Q) * a.sigma * b.sigma;
  else if (Q < 706.0)
    *v = prefactor * std::pow(2.0, -(nu - 1)) / std::tgamma(nu) * std::pow(Q, nu) * bessel_k(nu, Q) * a.sigma *
         b.sigma;
  else
    *v = prefactor * std::pow(2.0, -(nu - 1)) / std::tgamma(nu) * std::pow(Q, nu) * std::sqrt(M_PI / (2.0 * Q)) *
         std::exp(-Q) * a.sigma * b.sigma;
  return true;
}

// cov_rns_taper, src/cocons_taper.cpp:151-433.  out: nnz entries in CSR order.
int oracle_cov_rns_taper(long n, long p, const double* locs, const double* X, const double* theta6,
                         const double* limits, const double* colindices, const double* rowpointers, long nnz,
                         double* out) {
  const double* sd = theta6;
  const double* smooth = theta6 + 4 * p;
  const double* nugget = theta6 + 5 * p;
  bool slopes_zero = true;  // types.h:56-63
  for (long k = 1; k < p; ++k)
    if (smooth[k] != 0) slopes_zero = false;
  int path = 0;
  double nu_fixed = 0.0;
  const bool fixed = slopes_zero && (limits[0] == limits[1]);  // :187
  if (fixed) {
    nu_fixed = limits[0];
    if (std::fabs(nu_fixed - 0.5) < 1e-6) path = 1;
    else if (std::fabs(nu_fixed - 1.5) < 1e-6) path = 2;
    else if (std::fabs(nu_fixed - 2.5) < 1e-6) path = 3;
  }
  std::vector<TaperSite> S;
  taper_site_stage(n, p, X, theta6, limits, !fixed, S);  // fixed non-half-integer nu: smooth vector stays 0
  long e = 0;
  for (long i = 0; i < n; ++i) {
    for (long w = (long)(rowpointers[i] - 1); w < (long)(rowpointers[i + 1] - 1); ++w, ++e) {
      const long j = (long)(colindices[w] - 1);
      const double dv = link_exp(lin_pred(X, n, p, i, sd)) + link_exp(lin_pred(X, n, p, i, nugget));
      double v;
      if (i == j || !taper_pair(S[i], S[j], locs[i] - locs[j], locs[n + i] - locs[n + j], path, nu_fixed, &v)) v = dv;
      out[e] = v;
    }
  }
  return e == nnz ? 0 : -1;
}

// cov_rns_taper_pred, src/cocons_taper.cpp:17-139.  Rows of the pattern are prediction sites.
int oracle_cov_rns_taper_pred(long n, long m, long p, const double* locs, const double* locs_pred, const double* X,
                              const double* X_pred, const double* theta6, const double* limits,
                              const double* colindices, const double* rowpointers, long nnz, double* out) {
  const double* nugget = theta6 + 5 * p;
  std::vector<TaperSite> S, P;
  taper_site_stage(n, p, X, theta6, limits, true, S);
  taper_site_stage(m, p, X_pred, theta6, limits, true, P);
  long e = 0;
  for (long i = 0; i < m; ++i) {
    for (long w = (long)(rowpointers[i] - 1); w < (long)(rowpointers[i + 1] - 1); ++w, ++e) {
      const long j = (long)(colindices[w] - 1);
      // coincident / Q <= eps value: sigma_pred^2 + nugget (:91, :111) - sigma squared, not E(std.dev)
      const double dv = P[i].sigma * P[i].sigma + link_exp(lin_pred(X_pred, m, p, i, nugget));
      double v;
      const bool same = locs_pred[i] == locs[j] && locs_pred[m + i] == locs[n + j];  // :89
      if (same || !taper_pair(P[i], S[j], locs_pred[i] - locs[j], locs_pred[m + i] - locs[n + j], 0, 0.0, &v)) v = dv;
      out[e] = v;
    }
  }
  return e == nnz ? 0 : -1;
}

// sumsmoothlone, src/cocons_full.cpp:12-30
double oracle_sumsmoothlone(const double* x, long len, double lambda, double alpha) {
  double sum = 0;
  for (long w = 0; w < len; ++w) {
    if (std::abs(x[w]) > 1e-4)
      sum = sum + std::abs(x[w]);
    else
      sum = sum + std::pow(alpha, -1) * (std::log(1 + std::exp(-alpha * x[w])) + std::log(1 + std::exp(alpha * x[w])));
  }
  return lambda * sum;
}

// K_nu(x) as the oracle evaluates it (exposed so tests can pin it against mpmath)
double oracle_bessel_k(double nu, double x) { return bessel_k(nu, x); }

}  // extern "C"
