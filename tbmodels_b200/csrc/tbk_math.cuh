// tbk_math.cuh -- scalar building blocks shared by the device kernels and the host-side unit tests.
//
// Everything here is __host__ __device__ so the exact code the kernels run can also be exercised on
// the CPU (tests/test_native_host.py drives it through the tbk_host_* debug entry points).
//
// Packed Hermitian layout ("hp"), N*N doubles per matrix, used for H(k) between the build and the
// eigensolver kernels:
//   real plane : element (i,j), j <= i  at  i*(i+1)/2 + j                      (N(N+1)/2 doubles)
//   imag plane : element (i,j), j <  i  at  N(N+1)/2 + i*(i-1)/2 + j            (N(N-1)/2 doubles)
// Only the lower triangle is stored -- scipy.linalg.eigvalsh reads the lower triangle only
// (reference src/tbmodels/_tb_model.py:1149 -> LAPACK zheevr, UPLO='L').
#pragma once
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define TBK_HD __host__ __device__ __forceinline__
#else
#define TBK_HD inline
#endif

namespace tbk {

TBK_HD long tri(long i) { return i * (i + 1) / 2; }  // start of row i in the real plane
TBK_HD long trs(long i) { return i * (i - 1) / 2; }  // start of row i in the imag plane (strict lower)

// Elementary reflector (same contract as LAPACK zlarfg, restated):
// given alpha = (ar, ai) and xnorm2 = sum |x_i|^2 over the remaining entries, produce a real beta, a
// complex tau and a complex scale = 1/(alpha - beta) so that with v = [1; scale * x]
//   (I - tau v v^H)^H [alpha; x] = [beta; 0].
// tau == 0 means "no reflection" (x == 0 and alpha real).
#if defined(__CUDA_ARCH__)
// 1/x and 1/sqrt(x) for normal, finite, positive-or-negative (rcp) / positive (rsqrt) x: hardware seed (MUFU.RCP64H /
// MUFU.RSQ64H, ~2^-23) + three Newton steps -> <= 1 ulp.  No subnormal / inf / nan handling: the callers' arguments
// are sums of squares and norms of O(max |H|) numbers.
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
// Two Newton steps: rcp.approx.ftz.f64 reads the upper 20 mantissa bits of its operand (relative error ~ 2^-20), so the
// error after two steps (~ 2^-80) is already far below one ulp.  Used where the reciprocal sits in a long dependent chain
// and a last-bit rounding difference is immaterial (the Sturm recurrence of the bisection: 50 N^2 of them per matrix).
__device__ __forceinline__ double fast_rcp2(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    return fma(y, e, y);
}
#endif

// sin(pi t), cos(pi t) for finite |t| < 2^51: exact reduction t = n/2 + r, |r| <= 1/4, polynomials in r^2 obtained by
// Chebyshev economisation of the Taylor series on [0, 1/16] (exact rational arithmetic; truncation < 1e-17 for the sine
// with 6 coefficients, 1.2e-19 for the cosine with 7), quadrant fix-up.  Same accuracy class (<= 2 ulp) as CUDA's sincospi
// without its special-case handling (inf / nan / huge arguments): about 40 instead of 61 instructions.
TBK_HD void sincospi_lean(double t, double& s, double& c) {
    const double n = rint(t + t);
    const double r = fma(n, -0.5, t);
    const long long q = (long long)n;
    const double r2 = r * r;
    double ps = 0.00046221129498806536;
    ps = fma(ps, r2, -0.0073701436737117);
    ps = fma(ps, r2, 0.08214587730806341);
    ps = fma(ps, r2, -0.5992645291845754);
    ps = fma(ps, r2, 2.550164039876616);
    ps = fma(ps, r2, -5.167712780049969);
    double sv = fma(r * r2, ps, r * 1.2246467991473532e-16);  // pi = 3.141592653589793 + 1.2246467991473532e-16
    sv = fma(r, 3.141592653589793, sv);
    double pc = -0.00010370086335346046;
    pc = fma(pc, r2, 0.0019294938812021883);
    pc = fma(pc, r2, -0.025806887965145374);
    pc = fma(pc, r2, 0.2353306302840066);
    pc = fma(pc, r2, -1.3352627688538097);
    pc = fma(pc, r2, 4.058712126416765);
    pc = fma(pc, r2, -4.934802200544679);
    double cv = fma(pc, r2, 1.0);
    if (q & 1) {
        const double tmp = sv;
        sv = cv;
        cv = -tmp;
    }
    if (q & 2) {
        sv = -sv;
        cv = -cv;
    }
    s = sv;
    c = cv;
}

TBK_HD void householder_gen(double ar, double ai, double xnorm2, double& beta, double& tr, double& ti,
                            double& sr, double& si) {
    if (xnorm2 == 0.0 && ai == 0.0) {
        beta = ar;
        tr = ti = 0.0;
        sr = si = 0.0;
        return;
    }
    const double h = ar * ar + ai * ai + xnorm2;
#if defined(__CUDA_ARCH__)
    const bool tame = h > 1e-280 && h < 1e280;
    const double rn = tame ? fast_rsqrt(h) : 1.0 / sqrt(h);
#else
    const double rn = 1.0 / sqrt(h);
#endif
    const double nrm = h * rn;
    beta = (ar >= 0.0) ? -nrm : nrm;
    const double binv = (ar >= 0.0) ? -rn : rn;  // 1 / beta
    tr = (beta - ar) * binv;
    ti = -ai * binv;
    const double dr = ar - beta;  // same-sign sum: no cancellation
    const double di = ai;
    const double q = dr * dr + di * di;
#if defined(__CUDA_ARCH__)
    const double den = tame ? fast_rcp(q) : 1.0 / q;
#else
    const double den = 1.0 / q;
#endif
    sr = dr * den;
    si = -di * den;
}

// Serial Hermitian -> real symmetric tridiagonal reduction on one packed matrix (unblocked Householder,
// lower storage).  Element e of the matrix lives at A[e * s]; d/e entries at d[i * sde], e[i * sde];
// wv is scratch for 4*N doubles (element q at wv[q * s]).  Used by the thread-per-matrix small-N kernel
// and by the host tests; the cooperative kernel in eig_tridiag.cu performs the same steps with a
// thread group per matrix.
TBK_HD void hetrd_serial(int N, double* A, long s, double* d, double* e, long sde, double* wv) {
    const long NRE = tri(N);
    double* Ar = A;
    double* Ai = A + NRE * s;
    double* Vr = wv;
    double* Vi = wv + (long)N * s;
    double* Pr = wv + 2L * N * s;
    double* Pi = wv + 3L * N * s;
    for (int j = 0; j < N - 1; ++j) {
        d[j * sde] = Ar[(tri(j) + j) * s];
        const int m = N - 1 - j;
        const int r0 = j + 1;
        const double ar = Ar[(tri(r0) + j) * s];
        const double ai = Ai[(trs(r0) + j) * s];
        double xn = 0.0;
        for (int a = 1; a < m; ++a) {
            const int I = r0 + a;
            const double xr = Ar[(tri(I) + j) * s], xi = Ai[(trs(I) + j) * s];
            xn += xr * xr + xi * xi;
        }
        double beta, tr, ti, sr, si;
        householder_gen(ar, ai, xn, beta, tr, ti, sr, si);
        e[j * sde] = beta;
        if (tr == 0.0 && ti == 0.0) continue;
        Vr[0] = 1.0;
        Vi[0] = 0.0;
        for (int a = 1; a < m; ++a) {
            const int I = r0 + a;
            const double xr = Ar[(tri(I) + j) * s], xi = Ai[(trs(I) + j) * s];
            Vr[a * s] = xr * sr - xi * si;
            Vi[a * s] = xr * si + xi * sr;
        }
        // p = tau * A22 * v ; dot = p^H v
        double dr = 0.0, di = 0.0;
        for (int a = 0; a < m; ++a) {
            const int I = r0 + a;
            double sumr = 0.0, sumi = 0.0;
            for (int b = 0; b < m; ++b) {
                const int J = r0 + b;
                const int lo = I < J ? I : J, hi = I < J ? J : I;
                const double arr = Ar[(tri(hi) + lo) * s];
                double aii = (hi != lo) ? Ai[(trs(hi) + lo) * s] : 0.0;
                if (J > I) aii = -aii;
                const double vr = Vr[b * s], vi = Vi[b * s];
                sumr += arr * vr - aii * vi;
                sumi += arr * vi + aii * vr;
            }
            const double pr = tr * sumr - ti * sumi;
            const double pi = tr * sumi + ti * sumr;
            Pr[a * s] = pr;
            Pi[a * s] = pi;
            dr += pr * Vr[a * s] + pi * Vi[a * s];
            di += pr * Vi[a * s] - pi * Vr[a * s];
        }
        const double alr = -0.5 * (tr * dr - ti * di);
        const double ali = -0.5 * (tr * di + ti * dr);
        for (int a = 0; a < m; ++a) {
            const double vr = Vr[a * s], vi = Vi[a * s];
            Pr[a * s] += alr * vr - ali * vi;
            Pi[a * s] += alr * vi + ali * vr;
        }
        // A22 -= v w^H + w v^H  (lower triangle only)
        for (int a = 0; a < m; ++a) {
            const int I = r0 + a;
            const double var = Vr[a * s], vai = Vi[a * s], war = Pr[a * s], wai = Pi[a * s];
            for (int b = 0; b < a; ++b) {
                const int J = r0 + b;
                const double vbr = Vr[b * s], vbi = Vi[b * s], wbr = Pr[b * s], wbi = Pi[b * s];
                Ar[(tri(I) + J) * s] -= var * wbr + vai * wbi + war * vbr + wai * vbi;
                Ai[(trs(I) + J) * s] -= vai * wbr - var * wbi + wai * vbr - war * vbi;
            }
            Ar[(tri(I) + I) * s] -= 2.0 * (var * war + vai * wai);
        }
    }
    d[(long)(N - 1) * sde] = Ar[(tri(N - 1) + (N - 1)) * s];
}

// Eigenvalues of a real symmetric tridiagonal matrix by the implicitly shifted QL iteration
// (textbook algorithm, e.g. Wilkinson/Reinsch "imtql1").  d[i*sd] diagonal (n), e[i*sd] sub-diagonal
// (n-1 used, e[(n-1)*sd] is scratch).  On return d holds the eigenvalues sorted ascending.
// Returns the number of eigenvalues that failed to converge in 64 iterations (0 = success).
TBK_HD int tridiag_ql(int n, double* d, double* e, long sd) {
    int fails = 0;
    if (n <= 0) return 0;
    e[(long)(n - 1) * sd] = 0.0;
    for (int l = 0; l < n; ++l) {
        int iter = 0;
        int m;
        do {
            for (m = l; m < n - 1; ++m) {
                const double dd = fabs(d[m * sd]) + fabs(d[(m + 1) * sd]);
                if (fabs(e[m * sd]) <= DBL_EPSILON * dd) break;
            }
            if (m != l) {
                if (iter++ == 64) {
                    ++fails;
                    break;
                }
                const double el = e[l * sd];
                double g = (d[(l + 1) * sd] - d[l * sd]) / (2.0 * el);
                double r = sqrt(g * g + 1.0);
                g = d[m * sd] - d[l * sd] + el / (g + (g >= 0.0 ? r : -r));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                bool underflow = false;
                for (i = m - 1; i >= l; --i) {
                    const double ei = e[i * sd];
                    double f = s * ei;
                    const double b = c * ei;
                    const double h = f * f + g * g;
                    if (h == 0.0) {
                        e[(i + 1) * sd] = 0.0;
                        d[(i + 1) * sd] -= p;
                        e[m * sd] = 0.0;
                        underflow = true;
                        break;
                    }
#if defined(__CUDA_ARCH__)
                    const double rinv = rsqrt(h);  // one reciprocal square root instead of sqrt + divide
#else
                    const double rinv = 1.0 / sqrt(h);
#endif
                    r = h * rinv;
                    e[(i + 1) * sd] = r;
                    s = f * rinv;
                    c = g * rinv;
                    g = d[(i + 1) * sd] - p;
                    r = (d[i * sd] - g) * s + 2.0 * c * b;
                    p = s * r;
                    d[(i + 1) * sd] = g + p;
                    g = c * r - b;
                }
                if (underflow) continue;
                d[l * sd] -= p;
                e[l * sd] = g;
                e[m * sd] = 0.0;
            }
        } while (m != l);
    }
    // insertion sort, ascending (QL leaves the spectrum almost ordered)
    for (int i = 1; i < n; ++i) {
        const double x = d[i * sd];
        int j = i - 1;
        while (j >= 0 && d[j * sd] > x) {
            d[(j + 1) * sd] = d[j * sd];
            --j;
        }
        d[(j + 1) * sd] = x;
    }
    return fails;
}

// Number of eigenvalues of the symmetric tridiagonal (d, e) that are < x (Sturm sequence of the LDL^T pivots of
// T - x I; e2[i] = e[i]^2).  |pivot| is kept >= pivmin so the recurrence never divides by zero.
TBK_HD int sturm_count(int n, const double* d, const double* e2, double x, double pivmin) {
    double q = d[0] - x;
    if (fabs(q) < pivmin) q = -pivmin;
    int cnt = q < 0.0 ? 1 : 0;
    for (int i = 1; i < n; ++i) {
#if defined(__CUDA_ARCH__)
        q = fma(-e2[i - 1], fast_rcp2(q), d[i] - x);
#else
        q = d[i] - x - e2[i - 1] / q;
#endif
        if (fabs(q) < pivmin) q = -pivmin;
        cnt += q < 0.0 ? 1 : 0;
    }
    return cnt;
}

// idx-th smallest eigenvalue (0-based) by bisection inside the Gershgorin interval [gl, gu].
TBK_HD double bisect_eig(int n, const double* d, const double* e2, int idx, double gl, double gu, double pivmin) {
    double lo = gl, hi = gu;
    const double tol = 4.0 * DBL_EPSILON * fmax(fabs(gl), fabs(gu)) + 2.0 * pivmin;
    for (int it = 0; it < 80 && hi - lo > tol; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (mid <= lo || mid >= hi) break;
        if (sturm_count(n, d, e2, mid, pivmin) <= idx) lo = mid;
        else hi = mid;
    }
    return 0.5 * (lo + hi);
}

// Closed forms for the two smallest sizes (packed input: N=1 -> {h00}; N=2 -> {h00, re h10, h11, im h10}).
TBK_HD void eig2_closed(double h00, double h11, double br, double bi, double& lo, double& hi) {
    const double mean = 0.5 * (h00 + h11);
    const double delta = 0.5 * (h00 - h11);
    const double rad = sqrt(delta * delta + br * br + bi * bi);
    lo = mean - rad;
    hi = mean + rad;
}

}  // namespace tbk
