// Modified Bessel function K_nu(x) and the Matern correlation factor for the
// pairwise-assembly kernel.  Replaces boost::math::cyl_bessel_k at
// src/cocons_full.cpp:294,450,573 (reference) with a double-precision device
// routine: Temme's 1975 series for x <= 2, the trapezoidal rule on the integral
// representation for 2 < x < 25, the Hankel asymptotic tail beyond (Steed /
// Thompson-Barnett CF2, Boost's own middle-band method, is kept for nu > 6 and for
// nu > 3 beyond x = 18), organised for SIMT execution:
//   * one pair per lane; the band is a function of (nu, x) only, and the sites are Morton-ordered so that the 32
//     lanes of a warp see nearly the same x: a warp runs one body almost always (30.7 of 32 lanes active on
//     average);
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
#define COCONS_HD __device__ __forceinline__
#else
#define COCONS_HD inline
#endif

#include "gamma_coeffs.inc"
#include "log_table.inc"

namespace cocons {

constexpr double kPi = 3.141592653589793238462643383279502884;
constexpr double kHalfPi = 1.570796326794896619231321691639751442;
constexpr double kBesselEps = 1.0e-16;  // series / continued-fraction stopping level
constexpr double kHankelX = 25.0;       // Hankel tail used for x >= kHankelX and nu <= kHankelNuMax
constexpr double kHankelNuMax = 3.0;    // (terms shrink until k ~ 2x: 9e-16 from x = 25 on, ~23 terms there)
constexpr double kTrapXHighNu = 18.0;   // for kHankelNuMax < nu <= kTrapNuMax the trapezoidal rule stops here
constexpr int kHankelTerms = 40;

// exp(y) for -708 <= y < 709 (below: 0), the exponential of the hot loops (one call per node of the
// trapezoidal rule below: ~17 per pair of the assembly kernel).  Table-driven: N = round(128 y / ln 2) = 128 n + j,
// r = y - N ln2/128 (two-piece, |r| <= 0.0027), e^y = 2^n 2^(j/128) e^r with 2^(j/128) from a 128-entry table and
// e^r = 1 + r + r^2 (1/2 + r/6 + r^2/24 + r^3/120) (truncation 6e-19): 10 FP64 instructions, the
// library routine (and a first version here with a degree-13 polynomial and no table) needs 19 plus two UMOVs
// per 64-bit literal - ncu's source view of the assembly kernel: 11 % UMOV, 7 % LDCU.  The table is read with
// ld.global.nc (1 KB, L1-resident: lanes with different j do not serialise as they would on the constant
// cache).  Worst error against long double on 25 000 arguments: 1.5 ulp (tests/test_host.py).
#ifdef __CUDACC__
__device__
#endif
    static const double kExp2Tab[128] = {
        1, 1.0054299011128027264, 1.0108892860517004753, 1.0163783149109530957,
        1.0218971486541166271, 1.0274459491187637461, 1.0330248790212284149, 1.0386341019613787306,
        1.0442737824274137548, 1.0499440858006872102, 1.055645178360557157, 1.0613772272892620929,
        1.0671404006768236972, 1.0729348675259755552, 1.0787607977571198603, 1.0846183622133092062,
        1.0905077326652576897, 1.0964290818163768826, 1.1023825833078408909, 1.1083684117236787259,
        1.1143867425958924322, 1.1204377524096067464, 1.1265216186082418481, 1.1326385195987191956,
        1.1387886347566915646, 1.144972144431804173, 1.1511892299529826733, 1.1574400736337511209,
        1.1637248587775774755, 1.1700437696832501899, 1.1763969916502812207, 1.1827847109843410145,
        1.1892071150027210269, 1.1956643920398273284, 1.2021567314527030756, 1.2086843236265816248,
        1.2152473599804689552, 1.221846032972757623, 1.2284805361068700247, 1.2351510639369334132,
        1.241857812073484002, 1.2486009771892048192, 1.2553807570246910963, 1.2621973503942507389,
        1.2690509571917332199, 1.2759417783963920012, 1.2828700160787782636, 1.2898358734066657227,
        1.2968395546510096406, 1.3038812651919358121, 1.3109612115247644137, 1.3180796012660640493,
        1.3252366431597413232, 1.3324325470831615004, 1.3396675240533029161, 1.3469417862329458035,
        1.3542555469368926513, 1.3616090206382247541, 1.369002422974590516, 1.3764359707545301692,
        1.3839098819638320226, 1.3914243757719262362, 1.3989796725383112364, 1.4065759938190154354,
        1.4142135623730951455, 1.4218926021691655759, 1.4296133383919700233, 1.4373759974489823676,
        1.4451808069770466503, 1.4530279958490526226, 1.4609177941806470447, 1.4688504333369818422,
        1.4768261459394993462, 1.4848451658727523927, 1.492907728291264835, 1.5010140696264255844,
        1.5091644275934228414, 1.5173590411982147419, 1.5255981507445384171, 1.533881997840955913,
        1.5422108254079407441, 1.5505848776849999737, 1.5590044002378369292, 1.5674696399655529966,
        1.5759808451078864966, 1.5845382652524937495, 1.5931421513422669989, 1.6017927556826934143,
        1.6104903319492542835, 1.6192351351948637284, 1.628027421857347834, 1.6368674497669644108,
        1.6457554781539649458, 1.6546917676561943011, 1.6636765803267363761, 1.6727101796415966284,
        1.6817928305074290041, 1.6909247992693052787, 1.7001063537185234775, 1.7093377631004629258,
        1.7186192981224779341, 1.7279512309618376698, 1.7373338352737062174, 1.7467673861991690476,
        1.7562521603732994535, 1.7657884359332727264, 1.7753764925265211883, 1.7850166113189349648,
        1.7947090750031071682, 1.8044541678066239321, 1.8142521755003988559, 1.8241033854070534126,
        1.8340080864093424307, 1.8439665689586259845, 1.8539791250833854708, 1.8640460483977889794,
        1.8741676341102999626, 1.884344179032334532, 1.8945759815869656073, 1.9048633418176741383,
        1.9152065613971474001, 1.9256059436361250281, 1.9360617934922943473, 1.9465744175792332182,
        1.9571441241754001794, 1.9677712232331758813, 1.9784560263879509279, 1.9891988469672663431};

// Constants of the hot loops are chosen so that ptxas can encode them as 32-bit immediates (a double whose low
// word is zero) wherever the arithmetic allows it - every other 64-bit literal costs two UMOV issue slots per
// use, and the per-node body of the trapezoidal rule was 28 non-FP64 instructions against 19 FP64 ones:
//   * 128 log2(e) only picks N, r is formed exactly whatever N is: 21 significant bits suffice;
//   * ln2/128 = hi + lo with a 21-bit hi (fn hi is exact for |N| < 2^32) and a full-precision lo;
//   * 1/120 multiplies r^5 (|r| <= 0.0027): rounding it to 21 bits moves e^r by < 1e-22.
// ylo: a correction to the argument (|ylo| << 1) added after the reduction, where it is not lost to the rounding of y.
template <bool kCheckUnderflow, bool kHasLow = false>
COCONS_HD double exp_poly_impl(double y, double ylo = 0.0) {
  if (kCheckUnderflow && y < -708.0) return 0.0;
  const double kMagic = 6755399441055744.0;  // 2^52 + 2^51: the low word of 32 y log2(e) + kMagic is N
  const double t = fma(y, 0x1.71547p+7, kMagic);
  const double fn = t - kMagic;
  double r = fma(fn, -0x1.62e42p-8, y);          // exact
  r = fma(fn, -0x1.fdf473de6af28p-29, r);
  if (kHasLow) r += ylo;  // (not folded away for ylo = 0: r + 0.0 is not r for r = -0.0)
  double q = fma(r, 0x1.11111p-7, 1.0 / 24.0);
  q = fma(q, r, 1.0 / 6.0);
  q = fma(q, r, 0.5);
  const double p = fma(q * r, r, r);  // e^r - 1
#ifdef __CUDA_ARCH__
  const int N = __double2loint(t);
  const double T = __ldg(&kExp2Tab[N & 127]);
  const double e = fma(T, p, T);
  return __hiloint2double(__double2hiint(e) + ((N >> 7) << 20), __double2loint(e));
#else
  const long long N = (long long)fn;
  const double T = kExp2Tab[N & 127];
  return ldexp(fma(T, p, T), (int)(N >> 7));
#endif
}

COCONS_HD double exp_poly(double y) { return exp_poly_impl<true>(y); }
COCONS_HD double exp_poly2(double y, double ylo) { return exp_poly_impl<true, true>(y, ylo); }
// for arguments known to stay above -708 (the nodes of the trapezoidal rule: the loop leaves long before)
COCONS_HD double exp_poly_nocheck(double y) { return exp_poly_impl<false>(y); }


// ln(x) = H + L (unevaluated sum, |L| << |H| or both tiny) for normal positive x, absolute error < 1e-18.
// x = 2^e m, m in [1, 2); j = the top six mantissa bits pick c_j = 1 + (j + 1/2)/64 and the table gives
// 1/c_j and -ln(1/c_j) = hi_j + lo_j (gen_log_table.py); r = m/c_j - 1 (one fma, |r| < 0.0078) and
// ln(1 + r) = r + r^2 (-1/2 + r/3 - ... - r^6/8)  (next term 1e-20).  e ln2_hi + hi_j is exact (both multiples
// of 2^-40), its sum with r goes through a two-sum, everything small is collected in L.  21 FP64 instructions
// and two table loads; the library log (one double, 0.5 ulp: 4e-16 absolute at ln 350) needs ~75 instructions,
// 18 of them UMOVs.  The Matern factor multiplies this by nu and exponentiates, so what counts is the
// ABSOLUTE error, and the H + L form removes the rounding of the logarithm from the result altogether.
#ifdef __CUDACC__
__device__
#endif
    static const double kLogTabA[128]
#ifdef __CUDACC__
    __attribute__((aligned(16)))
#endif
    = COCONS_LOG_TAB_A;
#ifdef __CUDACC__
__device__
#endif
    static const double kLogTabB[64] = COCONS_LOG_TAB_B;

COCONS_HD void log_hl(double x, double& H, double& L) {
#ifdef __CUDA_ARCH__
  const int hx = __double2hiint(x);
  const int j = (hx >> 14) & 63;
  const double e = (double)((hx >> 20) - 1023);
  const double m = __hiloint2double((hx & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double2 t = __ldg(reinterpret_cast<const double2*>(kLogTabA) + j);
  const double inv = t.x, thi = t.y, tlo = __ldg(&kLogTabB[j]);
#else
  int ex;
  const double m = 2.0 * frexp(x, &ex);  // [1, 2)
  const double e = (double)(ex - 1);
  const int j = (int)((m - 1.0) * 64.0);
  const double inv = kLogTabA[2 * j], thi = kLogTabA[2 * j + 1], tlo = kLogTabB[j];
#endif
  const double r = fma(m, inv, -1.0);
  double p = fma(r, -0.125, 0x1.24925p-3);  // -1/8, 1/7 (21 bits: the term is < 3e-16)
  p = fma(p, r, -0x1.55555p-3);             // -1/6 (21 bits)
  p = fma(p, r, 0.2);
  p = fma(p, r, -0.25);
  p = fma(p, r, 1.0 / 3.0);
  p = fma(p, r, -0.5);
  const double q = (r * r) * p;
  const double h0 = fma(e, COCONS_LN2_HI, thi);  // exact
  const double s = h0 + r;
  const double bb = s - h0;
  const double err = (h0 - (s - bb)) + (r - bb);
  H = s;
  L = (fma(e, COCONS_LN2_LO, tlo) + err) + q;
}

// 1/d for a normal d away from the ends of the exponent range, ~1 ulp: the hardware seed (20 bits) and two
// Newton steps - 5 instructions where the IEEE division is ~22 with its range check and slow path
COCONS_HD double rcp_fast(double d) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(r, fma(-d, r, 1.0), r);
  return fma(r, fma(-d, r, 1.0), r);
#else
  return 1.0 / d;
#endif
}

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
  if (nl == 1) return rgp;
  // (mu + 1)(mu + 2)... in the order of the plain loop; straight-line up to nu < 4.5
  double prod = mu + 1.0;
  if (nl > 2) prod *= mu + 2.0;
  if (nl > 3) prod *= mu + 3.0;
  for (int k = 4; k < nl; ++k) prod *= (mu + (double)k);
  return rgp * rcp_fast(prod);
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

// reciprocals 1/i for the CF2 recurrence (uniform index: constant-cache broadcast on the device)
#ifdef __CUDACC__
__constant__
#else
static const
#endif
    double kInvInt[128] = {
        0.0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11, 1.0 / 12,
        1.0 / 13, 1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19, 1.0 / 20, 1.0 / 21, 1.0 / 22, 1.0 / 23,
        1.0 / 24, 1.0 / 25, 1.0 / 26, 1.0 / 27, 1.0 / 28, 1.0 / 29, 1.0 / 30, 1.0 / 31, 1.0 / 32, 1.0 / 33, 1.0 / 34,
        1.0 / 35, 1.0 / 36, 1.0 / 37, 1.0 / 38, 1.0 / 39, 1.0 / 40, 1.0 / 41, 1.0 / 42, 1.0 / 43, 1.0 / 44, 1.0 / 45,
        1.0 / 46, 1.0 / 47, 1.0 / 48, 1.0 / 49, 1.0 / 50, 1.0 / 51, 1.0 / 52, 1.0 / 53, 1.0 / 54, 1.0 / 55, 1.0 / 56,
        1.0 / 57, 1.0 / 58, 1.0 / 59, 1.0 / 60, 1.0 / 61, 1.0 / 62, 1.0 / 63, 1.0 / 64, 1.0 / 65, 1.0 / 66, 1.0 / 67,
        1.0 / 68, 1.0 / 69, 1.0 / 70, 1.0 / 71, 1.0 / 72, 1.0 / 73, 1.0 / 74, 1.0 / 75, 1.0 / 76, 1.0 / 77, 1.0 / 78,
        1.0 / 79, 1.0 / 80, 1.0 / 81, 1.0 / 82, 1.0 / 83, 1.0 / 84, 1.0 / 85, 1.0 / 86, 1.0 / 87, 1.0 / 88, 1.0 / 89,
        1.0 / 90, 1.0 / 91, 1.0 / 92, 1.0 / 93, 1.0 / 94, 1.0 / 95, 1.0 / 96, 1.0 / 97, 1.0 / 98, 1.0 / 99, 1.0 / 100,
        1.0 / 101, 1.0 / 102, 1.0 / 103, 1.0 / 104, 1.0 / 105, 1.0 / 106, 1.0 / 107, 1.0 / 108, 1.0 / 109, 1.0 / 110,
        1.0 / 111, 1.0 / 112, 1.0 / 113, 1.0 / 114, 1.0 / 115, 1.0 / 116, 1.0 / 117, 1.0 / 118, 1.0 / 119, 1.0 / 120,
        1.0 / 121, 1.0 / 122, 1.0 / 123, 1.0 / 124, 1.0 / 125, 1.0 / 126, 1.0 / 127};

// e^x K_mu(x), e^x K_{mu+1}(x) for x > 2, |mu| <= 1/2  (Steed's algorithm for CF2).
// Thompson-Barnett's q / c recurrences are folded into one division-free three-term recurrence on
// the summand p_i = c_i q_i:   p_i = (b_i p_{i-1} + a_{i-1}/(i-1) p_{i-2}) / i,  a_i = -(1/4 - mu^2) - i(i-1),
// so that the only division left per step is the continued-fraction update d = 1/(b + a d).
COCONS_HD void bessel_k_cf2_scaled(double mu, double x, double& kmu, double& kmu1) {
  const double a1 = 0.25 - mu * mu;
  double b = 2.0 * (1.0 + x);
  double d = 1.0 / b;
  double h = d, delh = d;
  double p_prev = 0.0, p_cur = a1;  // p_0, p_1
  double qsum = a1;
  double s = fma(qsum, delh, 1.0);
  double a_prev = -a1;  // a_1
  for (int i = 2; i < 128; ++i) {
    const double a = a_prev - (double)(2 * (i - 1));  // a_i
    const double p_new = fma(b, p_cur, a_prev * kInvInt[i - 1] * p_prev) * kInvInt[i];
    p_prev = p_cur;
    p_cur = p_new;
    qsum += p_new;
    b += 2.0;
    d = 1.0 / fma(a, d, b);
    delh = fma(b, d, -1.0) * delh;
    h += delh;
    const double dels = qsum * delh;
    s += dels;
    a_prev = a;
    if (fabs(dels) < fabs(s) * kBesselEps) break;
  }
  h = a1 * h;
  kmu = sqrt(kHalfPi / x) / s;
  kmu1 = kmu * (mu + x + 0.5 - h) / x;
}

// e^x K_nu(x) for the middle band (2 < x < 25 for nu <= 3, 2 < x < 18 for 3 < nu <= 6) by the trapezoidal rule on
//     e^x K_nu(x) = int_0^inf exp(-x (cosh t - 1)) cosh(nu t) dt,
// step h = 5/32, nodes t_k = k h.  The integrand is entire and decays double-exponentially, so the rule
// converges geometrically (the error falls like exp(-pi^2 / h) against a factor growing with x and nu: measured
// 5e-16 up to x = 25 for nu <= 3 and up to x = 21 for nu <= 6) and the sum is cut when a term drops below
// 1e-17 of it - 13 terms at x = 25, 26 at x = 2, every term positive (no cancellation).  Per term:
// one exp, cosh(k h) - 1 from a table, cosh(nu k h) from the difference form of its three-term recurrence
// (C_{k+1} = C_k + D_{k+1}, D_{k+1} = D_k + 4 sinh^2(nu h / 2) C_k).  Worst error against 40-digit mpmath
// on a 41 x 24 grid of (x, nu <= 6): 1.0e-15 (Steed's CF2 above: 2.9e-15) at about a third of CF2's
// instructions - CF2 needs a division per step and ~40 steps at x = 2.
// Generated by: mpmath, 50 digits, (cosh(k * 5/32) - 1) * 128 / ln 2, k = 0..31 (the unit of exp_node below).
constexpr double kTrapH = 0.15625;
constexpr double kTrapNuMax = 6.0;
constexpr int kTrapNodes = 32;
#ifdef __CUDACC__
__constant__
#else
static const
#endif
    double kTrapC[kTrapNodes] = {
        0.0, 2.25880093930198081271, 9.09046255171647670791, 20.662113059400613363,
        37.2568386991285421201, 59.2806090798623185295, 87.2722087313040666854, 121.916417806583484605,
        164.060764388805504502, 214.736258226090705961, 275.182613120505841295, 346.878575005693816417,
        431.578097652424257722, 531.353250996330622076, 648.64491178744388428, 786.322476646090392466,
        947.754058331826991611, 1136.88888249103556949, 1358.35390061854541354, 1617.56698275089144305,
        1920.86945901161542931, 2275.68125247495073464, 2690.6823984831265807, 3176.02539106475206419,
        3743.58355124895790321, 4407.24149330147192419, 5183.23479478283033573, 6090.54718004008576447,
        7151.37493373898197335, 8391.669905744046004, 9841.7743912949811834, 11537.1634180441812647};

// e^{z ln2/128}, z = a c, for the nodes of the rule: the node constants c carry the factor 128/ln2, so N = round(z)
// and the reduced argument f = a c - N is ONE fma of the exact product (no two-piece reduction, and the product is
// never rounded); e^{f ln2/128} - 1 = f (a1 + f (a2 + f (a3 + f (a4 + f a5)))), a_i = (ln2/128)^i / i!, |f| <= 1/2
// (truncation 6e-19; a4, a5 as 21-bit immediates).  No underflow test: the loop leaves long before
// (z > -60 * 185).  9 FP64 instructions.
// tab: the 2^(j/128) table; the assembly kernel passes a copy in shared memory (a 32-bit address: two integer
// instructions fewer per node than the 64-bit global one), everyone else the global table.
COCONS_HD double exp_node(double a, double c, const double* tab) {
  const double kMagic = 6755399441055744.0;
  const double t = fma(a, c, kMagic);
  const double fn = t - kMagic;
  const double f = fma(a, c, -fn);
  double q = fma(f, 0x1.5d88000000000p-45, 0x1.3b2ab00000000p-35);
  q = fma(q, f, 0x1.c6b08d704a0c0p-26);
  q = fma(q, f, 0x1.ebfbdff82c58fp-17);
  q = fma(q, f, 0x1.62e42fefa39efp-8);
  const double p = q * f;
#ifdef __CUDA_ARCH__
  const int N = __double2loint(t);
  const double T = tab ? tab[N & 127] : __ldg(&kExp2Tab[N & 127]);
  const double e = fma(T, p, T);
  return __hiloint2double(__double2hiint(e) + ((N >> 7) << 20), __double2loint(e));
#else
  const long long N = (long long)fn;
  const double T = kExp2Tab[N & 127];
  return ldexp(fma(T, p, T), (int)(N >> 7));
#endif
}

COCONS_HD double bessel_k_trap_scaled(double nu, double x, const double* tab = nullptr) {
  // sinh(a), a = nu h / 2 <= 0.47: odd series to a^15 (next term 2e-18 relative)
  const double a = 0.5 * kTrapH * nu, a2 = a * a;
  // (the three highest coefficients rounded to 21 bits - immediates, see exp_poly - move sinh by < 2e-17 relative)
  double sh = fma(a2, 0x1.ae7f4p-41, 0x1.61246p-33);
  sh = fma(sh, a2, 0x1.ae645p-26);
  sh = fma(sh, a2, 1.0 / 362880.0);
  sh = fma(sh, a2, 1.0 / 5040.0);
  sh = fma(sh, a2, 1.0 / 120.0);
  sh = fma(sh, a2, 1.0 / 6.0);
  sh = fma(sh * a2, a, a);
  const double delta = 4.0 * sh * sh;  // 2 (cosh(nu h) - 1)
  double D = 0.5 * delta;              // C_1 - C_0
  double C = 1.0 + D;                  // cosh(nu h)
  double sum = 0.5;
  const double nx = -x;
  // fully unrolled (the node constants become constant-bank operands), exit test after every second node;
  // the 31st node is never needed (26 nodes at x = 2, the lower end of the band)
#ifdef COCONS_TRAP_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
  for (int k = 1; k < kTrapNodes - 1; k += 2) {
    const double t1 = exp_node(nx, kTrapC[k], tab) * C;
    D = fma(delta, C, D);
    C += D;
    const double t2 = exp_node(nx, kTrapC[k + 1], tab) * C;
    sum += t1 + t2;
    if (t2 < 5e-18) break;  // sum >= 1/2: below 1e-17 of it
    D = fma(delta, C, D);
    C += D;
  }
  return kTrapH * sum;
}

// e^x K_nu(x) from the Hankel expansion  sum_k a_k / x^k,  a_k = a_{k-1} (4 nu^2 - (2k-1)^2) / (8k),
// x >= kHankelX, 0 <= nu <= kHankelNuMax; summed until a term drops below 1e-17 of the sum (23 terms at
// x = 25, 12 at x = 45).  In that range the terms shrink monotonically for every k <= 40 (the ratio
// |36 - (2k-1)^2| / (8 x k) stays below 0.78), so the smallest-term test of an asymptotic series is not
// needed.  The k-dependent factors come interleaved from constant memory as {1/(8k), (2k-1)^2/(8k)}:
// term_k = term_{k-1} (fma(4 nu^2, 1/(8k), -(2k-1)^2/(8k)) / x) - four FP64 instructions and one 128-bit
// constant load per term, exit test every second term (round 1's loop: 21 instructions per term).
#ifdef __CUDACC__
__constant__
#else
static const
#endif
    double kHankelW[2 * kHankelTerms] = {
        0.125, 0.125, 0.0625, 0.5625,
        0.041666666666666666667, 1.0416666666666666667, 0.03125, 1.53125,
        0.025, 2.025, 0.020833333333333333333, 2.5208333333333333333,
        0.017857142857142857143, 3.0178571428571428571, 0.015625, 3.515625,
        0.013888888888888888889, 4.0138888888888888889, 0.0125, 4.5125,
        0.011363636363636363636, 5.0113636363636363636, 0.010416666666666666667, 5.5104166666666666667,
        0.0096153846153846153846, 6.0096153846153846154, 0.0089285714285714285714, 6.5089285714285714286,
        0.0083333333333333333333, 7.0083333333333333333, 0.0078125, 7.5078125,
        0.0073529411764705882353, 8.0073529411764705882, 0.0069444444444444444444, 8.5069444444444444444,
        0.0065789473684210526316, 9.0065789473684210526, 0.00625, 9.50625,
        0.005952380952380952381, 10.005952380952380952, 0.0056818181818181818182, 10.505681818181818182,
        0.0054347826086956521739, 11.005434782608695652, 0.0052083333333333333333, 11.505208333333333333,
        0.005, 12.005, 0.0048076923076923076923, 12.504807692307692308,
        0.0046296296296296296296, 13.00462962962962963, 0.0044642857142857142857, 13.504464285714285714,
        0.0043103448275862068966, 14.004310344827586207, 0.0041666666666666666667, 14.504166666666666667,
        0.0040322580645161290323, 15.004032258064516129, 0.00390625, 15.50390625,
        0.0037878787878787878788, 16.003787878787878788, 0.0036764705882352941176, 16.503676470588235294,
        0.0035714285714285714286, 17.003571428571428571, 0.0034722222222222222222, 17.503472222222222222,
        0.0033783783783783783784, 18.003378378378378378, 0.0032894736842105263158, 18.503289473684210526,
        0.0032051282051282051282, 19.003205128205128205, 0.003125, 19.503125};

COCONS_HD double bessel_k_hankel_scaled(double nu, double x) {
  const double four_nu2 = 4.0 * nu * nu;
  const double rx = rcp_fast(x);
  double term = 1.0, sum = 1.0;
#pragma unroll
  for (int k = 0; k < kHankelTerms; k += 2) {
    const double t1 = term * (fma(four_nu2, kHankelW[2 * k], -kHankelW[2 * k + 1]) * rx);
    term = t1 * (fma(four_nu2, kHankelW[2 * k + 2], -kHankelW[2 * k + 3]) * rx);
    sum += t1 + term;
    if (fabs(term) < 1e-17 * fabs(sum) && fabs(t1) < 1e-17 * fabs(sum)) break;
  }
  return sqrt(kHalfPi * rx) * sum;
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

// which body a given (nu, x) belongs to: 0 Temme, 1 CF2, 2 Hankel, 3 trapezoidal rule
COCONS_HD int bessel_band(double nu, double x) {
  if (x <= 2.0) return 0;
#ifdef COCONS_BESSEL_NO_TRAP  // ablation: round 1's layout (CF2 in the middle band, Hankel from x = 18)
  return (x >= 18.0 && nu <= kHankelNuMax) ? 2 : 1;
#else
  if (nu <= kHankelNuMax) return (x < kHankelX) ? 3 : 2;
  return (nu <= kTrapNuMax && x < kTrapXHighNu) ? 3 : 1;
#endif
}

// K_nu(x) itself (unscaled) - used by tests and by callers that want the plain value
COCONS_HD double bessel_k(double nu, double x) {
  const int nl = (int)(nu + 0.5);
  const double mu = nu - (double)nl;
  const int band = bessel_band(nu, x);
  if (band == 2) return bessel_k_hankel_scaled(nu, x) * exp(-x);
  if (band == 3) return bessel_k_trap_scaled(nu, x) * exp(-x);
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
COCONS_HD double matern_corr(double nu, double Q, const double* exp_tab = nullptr) {
  const int nl = (int)(nu + 0.5);
  const double mu = nu - (double)nl;
  const TemmeGammas G = temme_gammas(mu);
  const double two_rgamma = 2.0 * rgamma_from(G.rgp, mu, nl);
  const int band = bessel_band(nu, Q);
  // (Q/2)^nu e^{-Q} (x > 2: the Bessel bodies return e^x K) or (Q/2)^nu (Temme band, unscaled K) from ONE
  // exponential: nu (H + L) - Q is formed as a two-term sum, so neither the rounding of the logarithm nor that
  // of the difference reaches the result
  double H, L;
  log_hl(0.5 * Q, H, L);
  const double sub = (band == 0) ? 0.0 : Q;
  const double p = nu * H;
  const double s = p - sub;
  const double bb = s - p;
  const double lo = fma(nu, L, fma(nu, H, -p) + ((p - (s - bb)) - (sub + bb)));
  const double scale = exp_poly2(s, lo);
  if (band == 0) {
    double kmu, kmu1;
    bessel_k_temme(mu, Q, G, kmu, kmu1);
    return two_rgamma * scale * bessel_k_recur(kmu, kmu1, mu, Q, nl);
  }
  double ks;
  if (band == 2) {
    ks = bessel_k_hankel_scaled(nu, Q);
  } else if (band == 3) {
    ks = bessel_k_trap_scaled(nu, Q, exp_tab);
  } else {
    double kmu, kmu1;
    bessel_k_cf2_scaled(mu, Q, kmu, kmu1);
    ks = bessel_k_recur(kmu, kmu1, mu, Q, nl);
  }
  return two_rgamma * ks * scale;
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
