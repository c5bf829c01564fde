"""Generate the polynomial coefficients used by bessel.cuh for Temme's
gamma1(mu), gamma2(mu) on |mu| <= 1/2 (both even in mu):

    gamma1(mu) = (1/Gamma(1-mu) - 1/Gamma(1+mu)) / (2 mu)
    gamma2(mu) = (1/Gamma(1-mu) + 1/Gamma(1+mu)) / 2

They are interpolated at Chebyshev nodes in t = 8 mu^2 - 1 (50-digit mpmath),
then converted to monomials in t for a Horner evaluation.  Run:
    python gen_gamma_coeffs.py > gamma_coeffs.inc
"""
import mpmath as mp

mp.mp.dps = 60
N = 14  # number of Chebyshev terms kept


def g1(mu):
    if mu == 0:
        return -mp.euler * 0 + mp.diff(lambda m: mp.rgamma(1 - m) - mp.rgamma(1 + m), 0) / 2
    return (mp.rgamma(1 - mu) - mp.rgamma(1 + mu)) / (2 * mu)


def g2(mu):
    return (mp.rgamma(1 - mu) + mp.rgamma(1 + mu)) / 2


def cheb_coeffs(f, n):
    # f as a function of t in [-1,1]
    nodes = [mp.cos(mp.pi * (k + mp.mpf(1) / 2) / n) for k in range(n)]
    vals = [f(t) for t in nodes]
    c = []
    for j in range(n):
        s = mp.fsum(vals[k] * mp.cos(mp.pi * j * (k + mp.mpf(1) / 2) / n) for k in range(n))
        c.append(2 * s / n)
    c[0] /= 2
    return c


def cheb_to_mono(c):
    n = len(c)
    T = [[mp.mpf(0)] * n for _ in range(n)]
    T[0][0] = mp.mpf(1)
    if n > 1:
        T[1][1] = mp.mpf(1)
    for k in range(2, n):
        for j in range(n):
            T[k][j] = (2 * T[k - 1][j - 1] if j > 0 else 0) - T[k - 2][j]
    return [mp.fsum(c[k] * T[k][j] for k in range(n)) for j in range(n)]


def of_t(f):
    return lambda t: f(mp.sqrt((t + 1) / 8))


HIWORD_FROM = 6  # coefficients of t^6 and above (|c| < 1e-11, |t| <= 1) are rounded to doubles whose low 32 bits
# are zero: ptxas encodes those as immediates of the DFMA instead of two UMOVs each, and the polynomial moves by
# < 3e-18 (the functions are O(1))


def hiword(v):
    import struct
    b = struct.unpack("<Q", struct.pack("<d", float(v)))[0]
    b = (b + 0x80000000) & ~0xFFFFFFFF
    return struct.unpack("<d", struct.pack("<Q", b))[0]


def as_double(k, v):
    return hiword(v) if k >= HIWORD_FROM else float(v)


def emit(name, mono):
    print("// %s(mu) = sum_k c[k] t^k, t = 8 mu^2 - 1" % name)
    print("#define COCONS_%s_COEFFS { \\" % name.upper())
    for k, v in enumerate(mono):
        if k >= HIWORD_FROM:
            print("  %s, \\" % hiword(v).hex())
        else:
            print("  %s, \\" % mp.nstr(v, 20, min_fixed=0, max_fixed=0))
    print("}")


if __name__ == "__main__":
    for name, f in (("gamma1", g1), ("gamma2", g2)):
        c = cheb_coeffs(of_t(f), N)
        mono = cheb_to_mono(c)
        # report the truncation level and a dense check in double arithmetic
        worst = mp.mpf(0)
        for i in range(0, 1001):
            mu = mp.mpf(i) / 2000
            t = float(8 * mu * mu - 1)
            acc = 0.0
            for k in range(9, -1, -1):  # bessel.cuh evaluates the terms up to t^9
                acc = acc * t + as_double(k, mono[k])
            ref = f(mu)
            worst = max(worst, abs((mp.mpf(acc) - ref) / ref))
        print("// %s: last Chebyshev coeff %s, worst relative error of the double Horner form %s" % (
            name, mp.nstr(c[-1], 3), mp.nstr(worst, 3)))
        emit(name, mono)
