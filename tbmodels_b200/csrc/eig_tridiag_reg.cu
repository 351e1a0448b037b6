// eig_tridiag_reg.cu -- register-resident Hermitian -> tridiagonal reduction, one warp per matrix, 21 <= N <= 48.
//
// Same role as eig_tridiag.cu (first half of the replacement for the per-k scipy.linalg.eigvalsh loop of
// Model.eigenval, reference src/tbmodels/_tb_model.py:1148-1149; LAPACK zheevr JOBZ='N', UPLO='L') for the mid-size
// matrices of the Wannier-model workloads (C3: N = 36).  The shared-memory kernel spends ten instructions per useful
// DFMA there (triangle selects, address arithmetic, two LDS per four DFMA, 20 of 32 lanes active; ncu r01l: 35.8 k
// warp instructions and 12.4 k shared-memory wavefronts per 36 x 36 matrix, 18 matrices resident per SM).  Here
//   * lane c holds the FULL row c of the Hermitian matrix (both triangles, NMAX complex numbers = 4 NMAX registers);
//     the Hermitian matrix-vector product is a plain row . v (no transposed part, no selects) and the rank-2 update
//     A -= v w^H + w v^H touches only the lane's own registers;
//   * v and w are broadcast from shared memory (one LDS.128 per column, all lanes read the same address);
//   * every register index is a compile-time constant: the column loops are fully unrolled in blocks of four with a
//     warp-uniform early exit at the size of the trailing block, and the pivot column is picked by a switch.
// The reduction runs on the index-reversed matrix B[i][j] = A[N-1-i][N-1-j] and eliminates the LAST column of the
// active leading block each step -- which is exactly the lower-storage reduction of LAPACK zhetd2 / hetrd_serial
// (tbk_math.cuh) on A (same reflectors, same tridiagonal matrix up to the summation order): the active block stays
// anchored at index 0, so its rows are lanes 0 .. p-1 and its columns registers 0 .. p-1 for every step p.
// Matrices with N > 32 keep rows 32 .. N-1 implicitly: their entries left of column 32 are the conjugates of
// columns 32 .. N-1 of the lane rows (updated there anyway); only the (N-32)^2 corner block lives in shared memory.
// Those rows are eliminated first (N - 32 "corner" steps with a few extra warp reductions), after which the loop is
// the lean 32-lane form.
// Shared memory per matrix is only the load staging (N^2 doubles) + v, w + corner: residency is bounded by registers
// (9-10 warps per SM at NMAX = 36) and the kernel is FP64-pipe / issue bound instead of shared-memory bound.
// Bits depend on N only (fixed reduction orders), never on the batch.
#include "tbk_kernels.h"
#include "tbk_math.cuh"

namespace tbk {

namespace {

__device__ __forceinline__ int itri(int i) { return (i * (i + 1)) >> 1; }
__device__ __forceinline__ int itrs(int i) { return (i * (i - 1)) >> 1; }

__device__ __forceinline__ double warp_sum(double a) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    return a;
}

template <int B, int NMAX>
__device__ __forceinline__ void pick(const double (&ar)[NMAX], const double (&ai)[NMAX], double& xr, double& xi) {
    if constexpr (B < NMAX) {
        xr = ar[B];
        xi = ai[B];
    }
}

// x = column p of the lane's row (p is warp-uniform; a switch keeps the register indices static)
template <int NMAX>
__device__ __forceinline__ void get_col(const double (&ar)[NMAX], const double (&ai)[NMAX], int p, double& xr, double& xi) {
    xr = 0.0;
    xi = 0.0;
    switch (p) {
#define TBK_PICK(B) \
    case B:         \
        pick<B, NMAX>(ar, ai, xr, xi); \
        break;
        TBK_PICK(0) TBK_PICK(1) TBK_PICK(2) TBK_PICK(3) TBK_PICK(4) TBK_PICK(5) TBK_PICK(6) TBK_PICK(7)
        TBK_PICK(8) TBK_PICK(9) TBK_PICK(10) TBK_PICK(11) TBK_PICK(12) TBK_PICK(13) TBK_PICK(14) TBK_PICK(15)
        TBK_PICK(16) TBK_PICK(17) TBK_PICK(18) TBK_PICK(19) TBK_PICK(20) TBK_PICK(21) TBK_PICK(22) TBK_PICK(23)
        TBK_PICK(24) TBK_PICK(25) TBK_PICK(26) TBK_PICK(27) TBK_PICK(28) TBK_PICK(29) TBK_PICK(30) TBK_PICK(31)
#undef TBK_PICK
        default: break;
    }
}

// q = (row of this lane) . v over the columns b < m (block granular: V is zero from m up to the next multiple of 4)
template <int NMAX, int BW>
__device__ __forceinline__ void row_dot(const double (&ar)[NMAX], const double (&ai)[NMAX], const double2* __restrict__ V,
                                        int m, double& qr, double& qi) {
    double sr[4] = {0.0, 0.0, 0.0, 0.0}, si[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int b0 = 0; b0 < NMAX; b0 += BW) {
        if (b0 >= m) break;
#pragma unroll
        for (int j = 0; j < BW; ++j) {
            if (b0 + j < NMAX) {
                const double2 v = V[b0 + j];
                sr[j & 3] = fma(ar[b0 + j], v.x, sr[j & 3]);
                si[j & 3] = fma(ar[b0 + j], v.y, si[j & 3]);
                sr[j & 3] = fma(-ai[b0 + j], v.y, sr[j & 3]);
                si[j & 3] = fma(ai[b0 + j], v.x, si[j & 3]);
            }
        }
    }
    qr = (sr[0] + sr[1]) + (sr[2] + sr[3]);
    qi = (si[0] + si[1]) + (si[2] + si[3]);
}

// row -= v_c conj(w_b) + w_c conj(v_b) over the columns b < m (block granular; V, W zero beyond m)
template <int NMAX, int BW>
__device__ __forceinline__ void row_update(double (&ar)[NMAX], double (&ai)[NMAX], const double2* __restrict__ V,
                                           const double2* __restrict__ W, int m, double vr, double vi, double wr, double wi) {
#pragma unroll
    for (int b0 = 0; b0 < NMAX; b0 += BW) {
        if (b0 >= m) break;
#pragma unroll
        for (int j = 0; j < BW; ++j) {
            if (b0 + j < NMAX) {
                const double2 vb = V[b0 + j], wb = W[b0 + j];
                ar[b0 + j] = fma(-vr, wb.x, fma(-vi, wb.y, fma(-wr, vb.x, fma(-wi, vb.y, ar[b0 + j]))));
                ai[b0 + j] = fma(-vi, wb.x, fma(vr, wb.y, fma(-wi, vb.x, fma(wr, vb.y, ai[b0 + j]))));
            }
        }
    }
}

// ---- lookahead: the reflector of step p - 1 is generated WHILE the rank-2 update of step p is in flight ----
// The serial chain of a Householder step (pivot column -> norm reduction over the warp -> 1/sqrt -> reciprocal -> v) is
// ~1500 cycles of dependent shuffles and FP64 operations during which the warp issues nothing else.  The next pivot
// column is one column of the matrix: it is brought up to date first (8 FMAs per lane, the same expression the in-place
// update uses, so the bits agree), and the chain is cut into KSTEPS short stages that are issued between the 4-column
// blocks of the update -- each stage's latency then hides behind the 32 independent DFMAs of the following block.
struct Reflector {
    double xr, xi;          // this lane's entry of the pivot column
    double alr, ali;        // alpha = entry of lane p - 1
    double aa, xn, h, hx, y, e;
    double beta, binv, tr, ti, dr, q, den, sr, si;
    bool tame;
};
constexpr int KSTEPS = 12;

// stage j of the chain for pivot column p (lanes 0 .. p - 1 are active rows); same operation order as householder_gen
template <int J>
__device__ __forceinline__ void chain_step(Reflector& r, int p, int lane) {
    if constexpr (J == 0) {
        r.alr = __shfl_sync(0xffffffffu, r.xr, p - 1);
        r.ali = __shfl_sync(0xffffffffu, r.xi, p - 1);
        r.xn = lane < p - 1 ? fma(r.xr, r.xr, r.xi * r.xi) : 0.0;
    } else if constexpr (J >= 1 && J <= 5) {
        r.xn += __shfl_xor_sync(0xffffffffu, r.xn, 32 >> J);
        if constexpr (J == 5) {
            r.h = r.alr * r.alr + r.ali * r.ali + r.xn;
            r.tame = r.h > 1e-280 && r.h < 1e280;
            r.hx = 0.5 * r.h;
        }
    } else if constexpr (J == 6) {
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r.y) : "d"(r.h));
        r.e = fma(-r.hx * r.y, r.y, 0.5);
        r.y = fma(r.y, r.e, r.y);
    } else if constexpr (J == 7) {
        r.e = fma(-r.hx * r.y, r.y, 0.5);
        r.y = fma(r.y, r.e, r.y);
    } else if constexpr (J == 8) {
        r.e = fma(-r.hx * r.y, r.y, 0.5);
        r.y = fma(r.y, r.e, r.y);  // = fast_rsqrt(h)
        const double nrm = r.h * r.y;
        r.beta = (r.alr >= 0.0) ? -nrm : nrm;
        r.binv = (r.alr >= 0.0) ? -r.y : r.y;
        r.tr = (r.beta - r.alr) * r.binv;
        r.ti = -r.ali * r.binv;
        r.dr = r.alr - r.beta;
        r.q = r.dr * r.dr + r.ali * r.ali;
    } else if constexpr (J == 9) {
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r.den) : "d"(r.q));
        r.e = fma(-r.q, r.den, 1.0);
        r.den = fma(r.den, r.e, r.den);
    } else if constexpr (J == 10) {
        r.e = fma(-r.q, r.den, 1.0);
        r.den = fma(r.den, r.e, r.den);
    } else if constexpr (J == 11) {
        r.e = fma(-r.q, r.den, 1.0);
        r.den = fma(r.den, r.e, r.den);  // = fast_rcp(q)
        r.sr = r.dr * r.den;
        r.si = -r.ali * r.den;
        if (!r.tame || (r.xn == 0.0 && r.ali == 0.0))  // warp uniform: huge / tiny norms and the trivial reflector
            householder_gen(r.alr, r.ali, r.xn, r.beta, r.tr, r.ti, r.sr, r.si);
    }
}

template <int J0>
__device__ __forceinline__ void chain_finish(Reflector& r, int p, int lane, int done) {
    if constexpr (J0 < KSTEPS) {
        if (J0 >= done) chain_step<J0>(r, p, lane);
        chain_finish<J0 + 1>(r, p, lane, done);
    }
}

__device__ __forceinline__ void chain_all(Reflector& r, int p, int lane) { chain_finish<0>(r, p, lane, 0); }

// row -= v_c conj(w_b) + w_c conj(v_b) over the columns b < m in blocks of four, one chain stage after every block;
// returns the number of chain stages issued
template <int NMAX, int B0>
__device__ __forceinline__ int row_update_la(double (&ar)[NMAX], double (&ai)[NMAX], const double2* __restrict__ V,
                                             const double2* __restrict__ W, int m, double vr, double vi, double wr,
                                             double wi, Reflector& r, int pn, int lane) {
    if constexpr (B0 < NMAX) {
        if (B0 >= m) return B0 / 4 < KSTEPS ? B0 / 4 : KSTEPS;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (B0 + j < NMAX) {
                const double2 vb = V[B0 + j], wb = W[B0 + j];
                ar[B0 + j] = fma(-vr, wb.x, fma(-vi, wb.y, fma(-wr, vb.x, fma(-wi, vb.y, ar[B0 + j]))));
                ai[B0 + j] = fma(-vi, wb.x, fma(vr, wb.y, fma(-wi, vb.x, fma(wr, vb.y, ai[B0 + j]))));
            }
        }
        if constexpr (B0 / 4 < KSTEPS) chain_step<B0 / 4>(r, pn, lane);
        return row_update_la<NMAX, B0 + 4>(ar, ai, V, W, m, vr, vi, wr, wi, r, pn, lane);
    } else {
        return NMAX / 4 < KSTEPS ? NMAX / 4 : KSTEPS;
    }
}

template <int NREG, int OCC>
constexpr int reg_min_blocks() {  // resident one-warp CTAs per SM (registers are per SM sub-partition: 16 K each)
    return OCC ? OCC : (NREG <= 20 ? 16 : 12);  // 16: <= 128, 12: <= 168, 8: <= 255 registers per thread
}

// NREG: columns (and rows) held in registers, a multiple of 4 up to 32.  XMAX: capacity for the rows / columns beyond 32
// (N <= 32 + XMAX); 0 for N <= 32.
template <int NREG, int XMAX, int BW, int OCC, bool LA>
__global__ void __launch_bounds__(32, reg_min_blocks<NREG, OCC>())
tridiag_reg_kernel(double* __restrict__ Hp, int N, long mstride, long nk, double* __restrict__ D,
                   double* __restrict__ E, int ldo, int off, int n_stop) {
    static_assert(NREG % 4 == 0 && NREG <= 32, "NREG must be a multiple of 4, at most 32");
    static_assert(XMAX == 0 || NREG == 32, "extra rows only behind a full warp of register rows");
    constexpr int NV = NREG + XMAX < 32 ? 32 : NREG + XMAX;  // entries of v / w (every lane writes its own; a multiple of 8)
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x;
    const long kk = blockIdx.x;
    if (kk >= nk) return;
    const int ntri = itri(N);
    const int nn = N * N;
    double* S = smem;                                                    // load staging: the packed matrix
    double2* V = reinterpret_cast<double2*>(S + ((nn + 1) & ~1));        // [NV]
    double2* W = V + NV;                                                 // [NV]
    double2* C = W + NV;                                                 // [XMAX][XMAX] corner block B[32+r][32+s]
    double2* XC = C + XMAX * XMAX;                                       // [XMAX][32]   XC[s][c] = B[c][32+s]
    double* DS = reinterpret_cast<double*>(XC + XMAX * 32);              // [NV] diagonal of the reversed problem
    double* ES = DS + NV;                                                // [NV] its sub-diagonal

    {   // asynchronous copy of the packed matrix (every element in flight at once, no register staging)
        const double* src = Hp + kk * mstride;
        const unsigned sS = (unsigned)__cvta_generic_to_shared(S);
        if ((nn & 1) == 0 && (reinterpret_cast<unsigned long long>(src) & 15ull) == 0) {
            for (int e = lane; e < nn / 2; e += 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sS + 16u * e), "l"(src + 2 * e) : "memory");
        } else {
            for (int e = lane; e < nn; e += 32)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sS + 8u * e), "l"(src + e) : "memory");
        }
        for (int i = lane; i < 2 * NV; i += 32) V[i] = make_double2(0.0, 0.0);  // V and W are contiguous
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncwarp();

    // ---- registers: row `lane` of the index-reversed matrix B[c][b] = A[N-1-c][N-1-b], columns b < NREG ----
    double ar[NREG], ai[NREG];
    {
        const int I = N - 1 - lane;  // row of A held by this lane (lane < N)
        const bool row_ok = lane < N;
        const double* Si = S + ntri;
        const int Ic = row_ok ? I : 0;
        const int triI = itri(Ic), trsI = itrs(Ic);
#pragma unroll
        for (int b = 0; b < NREG; ++b) {
            // b >= lane (J <= I): stored entry (I, J); else the conjugate of the stored entry (J, I).  Branch-free:
            // select the two packed offsets, load, fix the sign, zero what lies outside the matrix / on the diagonal.
            const int J = b < N ? N - 1 - b : 0;
            const bool own = b >= lane;
            const int ire = own ? triI + J : itri(J) + Ic;
            int iim = own ? trsI + J : itrs(J) + Ic;
            const bool diag = b == lane;
            iim = diag ? 0 : iim;
            const double re = S[ire];
            const double im = Si[iim];
            const bool ok = row_ok && b < N;
            ar[b] = ok ? re : 0.0;
            ai[b] = (ok && !diag) ? (own ? im : -im) : 0.0;
        }
        if constexpr (XMAX > 0) {
            const int X = N - 32;  // rows / columns 32 .. N-1 of B = rows / columns X-1 .. 0 of A
            for (int s = 0; s < X; ++s) {  // B[lane][32+s] = A[I][X-1-s], I >= X > X-1-s: stored
                const int J = X - 1 - s;
                XC[s * 32 + lane] = make_double2(S[itri(I) + J], Si[itrs(I) + J]);
            }
            for (int e = lane; e < X * X; e += 32) {
                const int r = e / X, s = e - r * X;
                const int Ir = X - 1 - r, Js = X - 1 - s;
                double re, im = 0.0;
                if (Js <= Ir) {
                    re = S[itri(Ir) + Js];
                    if (Js < Ir) im = Si[itrs(Ir) + Js];
                } else {
                    re = S[itri(Js) + Ir];
                    im = -Si[itrs(Js) + Ir];
                }
                C[r * XMAX + s] = make_double2(re, im);
            }
        }
    }
    __syncwarp();

    int p = N - 1;

    if constexpr (XMAX > 0) {
        // ---- corner steps: pivot columns N-1 .. 32.  Rows 32 .. p-1 are the "second slot" of lanes 0 .. p-33; their
        // entries left of column 32 are conj(XC), the rest is the corner block C ----
        for (; p >= 32; --p) {
            const int xc = p - 32;  // active second-slot rows = active extra columns; pivot column of XC and C
            const double2 x1 = XC[xc * 32 + lane];
            const double xr = x1.x, xi = x1.y;
            double x2r = 0.0, x2i = 0.0;
            if (lane < xc) {
                const double2 z = C[lane * XMAX + xc];
                x2r = z.x;
                x2i = z.y;
            }
            if (lane == xc) DS[p] = C[xc * XMAX + xc].x;
            double alr, ali, xn;
            if (xc >= 1) {  // alpha sits in the second slot of lane xc - 1 (row p - 1)
                alr = __shfl_sync(0xffffffffu, x2r, xc - 1);
                ali = __shfl_sync(0xffffffffu, x2i, xc - 1);
                xn = fma(xr, xr, xi * xi);
                if (lane < xc - 1) xn += fma(x2r, x2r, x2i * x2i);
            } else {        // p == 32: alpha is row 31
                alr = __shfl_sync(0xffffffffu, xr, 31);
                ali = __shfl_sync(0xffffffffu, xi, 31);
                xn = lane < 31 ? fma(xr, xr, xi * xi) : 0.0;
            }
            xn = warp_sum(xn);
            double beta, tr, ti, sr, si;
            householder_gen(alr, ali, xn, beta, tr, ti, sr, si);
            if (lane == 0) ES[p - 1] = beta;
            if (tr == 0.0 && ti == 0.0) continue;
            double vr = xr * sr - xi * si, vi = xr * si + xi * sr;
            double v2r = 0.0, v2i = 0.0;
            if (xc >= 1) {
                if (lane < xc - 1) {
                    v2r = x2r * sr - x2i * si;
                    v2i = x2r * si + x2i * sr;
                } else if (lane == xc - 1) {
                    v2r = 1.0;
                }
            } else if (lane == 31) {
                vr = 1.0;
                vi = 0.0;
            }
            V[lane] = make_double2(vr, vi);
            if (lane < XMAX) V[32 + lane] = make_double2(v2r, v2i);
            __syncwarp();
            // q = B v.  Lane rows: register columns 0 .. 31, then the extra columns 32 .. p-1 from XC
            double qr, qi;
            row_dot<NREG, BW>(ar, ai, V, 32, qr, qi);
            for (int s = 0; s < xc; ++s) {
                const double2 z = XC[s * 32 + lane], v = V[32 + s];
                qr = fma(z.x, v.x, fma(-z.y, v.y, qr));
                qi = fma(z.x, v.y, fma(z.y, v.x, qi));
            }
            // second-slot rows: sum_{b<32} conj(B[b][32+r]) v_b (one warp reduction per row) + corner part
            double q2r = 0.0, q2i = 0.0;
            for (int r = 0; r < xc; ++r) {
                const double2 z = XC[r * 32 + lane];
                double tr_ = fma(z.x, vr, z.y * vi);
                double ti_ = fma(z.x, vi, -z.y * vr);
                tr_ = warp_sum(tr_);
                ti_ = warp_sum(ti_);
                if (lane == r) {
                    q2r = tr_;
                    q2i = ti_;
                }
            }
            if (lane < xc) {
                for (int s = 0; s < xc; ++s) {
                    const double2 z = C[lane * XMAX + s], v = V[32 + s];
                    q2r = fma(z.x, v.x, fma(-z.y, v.y, q2r));
                    q2i = fma(z.x, v.y, fma(z.y, v.x, q2i));
                }
            }
            const double pr = tr * qr - ti * qi, pi = tr * qi + ti * qr;
            const double p2r = tr * q2r - ti * q2i, p2i = tr * q2i + ti * q2r;
            double dr = pr * vr + pi * vi + (p2r * v2r + p2i * v2i);
            double di = pr * vi - pi * vr + (p2r * v2i - p2i * v2r);
            dr = warp_sum(dr);
            di = warp_sum(di);
            const double cr = -0.5 * (tr * dr - ti * di), ci = -0.5 * (tr * di + ti * dr);
            const double wr = pr + cr * vr - ci * vi, wi = pi + cr * vi + ci * vr;
            const double w2r = p2r + cr * v2r - ci * v2i, w2i = p2i + cr * v2i + ci * v2r;
            W[lane] = make_double2(wr, wi);
            if (lane < XMAX) W[32 + lane] = make_double2(w2r, w2i);
            __syncwarp();
            row_update<NREG, BW>(ar, ai, V, W, 32, vr, vi, wr, wi);
            for (int s = 0; s < xc; ++s) {
                const double2 vb = V[32 + s], wb = W[32 + s];
                double2 z = XC[s * 32 + lane];
                z.x = fma(-vr, wb.x, fma(-vi, wb.y, fma(-wr, vb.x, fma(-wi, vb.y, z.x))));
                z.y = fma(-vi, wb.x, fma(vr, wb.y, fma(-wi, vb.x, fma(wr, vb.y, z.y))));
                XC[s * 32 + lane] = z;
            }
            for (int e = lane; e < xc * xc; e += 32) {
                const int r = e / xc, s = e - r * xc;
                const double2 va = V[32 + r], wa = W[32 + r], vb = V[32 + s], wb = W[32 + s];
                double2 z = C[r * XMAX + s];
                z.x = fma(-va.x, wb.x, fma(-va.y, wb.y, fma(-wa.x, vb.x, fma(-wa.y, vb.y, z.x))));
                z.y = fma(-va.y, wb.x, fma(va.x, wb.y, fma(-wa.y, vb.x, fma(wa.x, vb.y, z.y))));
                C[r * XMAX + s] = z;
            }
            __syncwarp();
        }
    }

    // ---- lean steps: pivot columns min(N-1, 31) .. 1; active rows = lanes 0 .. p-1, active columns = registers 0 .. p-1 ----
    const int p_last = n_stop > 1 ? n_stop : 1;  // staged: stop with an n_stop x n_stop block left (n_stop <= 32)
    if constexpr (LA) {
        Reflector r;
        if (p >= p_last) {  // reflector of the first lean step, the plain way
            get_col<NREG>(ar, ai, p, r.xr, r.xi);
            if (lane == p) DS[p] = r.xr;
            chain_all(r, p, lane);
        }
        for (; p >= p_last; --p) {
            // r = reflector of step p (pivot column p): beta, tau, scale and this lane's pivot entry
            if (lane == 0) ES[p - 1] = r.beta;
            const double tr = r.tr, ti = r.ti;
            const bool has_next = p - 1 >= p_last;
            if (tr == 0.0 && ti == 0.0) {  // nothing to apply; the next pivot column is already up to date
                if (has_next) {
                    get_col<NREG>(ar, ai, p - 1, r.xr, r.xi);
                    if (lane == p - 1) DS[p - 1] = r.xr;
                    chain_all(r, p - 1, lane);
                }
                continue;
            }
            double vr = 0.0, vi = 0.0;
            if (lane < p - 1) {
                vr = r.xr * r.sr - r.xi * r.si;
                vi = r.xr * r.si + r.xi * r.sr;
            } else if (lane == p - 1) {
                vr = 1.0;
            }
            V[lane] = make_double2(vr, vi);  // zero from p on: the block-granular loops read up to the next multiple of 4
            __syncwarp();
            double qr, qi;
            row_dot<NREG, 4>(ar, ai, V, p, qr, qi);
            const double pr = tr * qr - ti * qi, pi = tr * qi + ti * qr;
            double dr = pr * vr + pi * vi;  // lanes >= p: v = 0
            double di = pr * vi - pi * vr;
            dr = warp_sum(dr);
            di = warp_sum(di);
            const double cr = -0.5 * (tr * dr - ti * di), ci = -0.5 * (tr * di + ti * dr);
            double wr = 0.0, wi = 0.0;
            if (lane < p) {
                wr = pr + cr * vr - ci * vi;
                wi = pi + cr * vi + ci * vr;
            }
            W[lane] = make_double2(wr, wi);
            __syncwarp();
            if (has_next) {
                // next pivot column = column p - 1 of this lane's row after the update (same expression as row_update_la)
                double cr0, ci0;
                get_col<NREG>(ar, ai, p - 1, cr0, ci0);
                const double2 vb = V[p - 1], wb = W[p - 1];
                r.xr = fma(-vr, wb.x, fma(-vi, wb.y, fma(-wr, vb.x, fma(-wi, vb.y, cr0))));
                r.xi = fma(-vi, wb.x, fma(vr, wb.y, fma(-wi, vb.x, fma(wr, vb.y, ci0))));
                if (lane == p - 1) DS[p - 1] = r.xr;
                const int done = row_update_la<NREG, 0>(ar, ai, V, W, p, vr, vi, wr, wi, r, p - 1, lane);
                chain_finish<0>(r, p - 1, lane, done);
            } else {
                row_update<NREG, 4>(ar, ai, V, W, p, vr, vi, wr, wi);
            }
            __syncwarp();
        }
    } else {
    for (; p >= p_last; --p) {
        double xr, xi;
        get_col<NREG>(ar, ai, p, xr, xi);
        if (lane == p) DS[p] = xr;
        const double alr = __shfl_sync(0xffffffffu, xr, p - 1);
        const double ali = __shfl_sync(0xffffffffu, xi, p - 1);
        double xn = lane < p - 1 ? fma(xr, xr, xi * xi) : 0.0;
        xn = warp_sum(xn);
        double beta, tr, ti, sr, si;
        householder_gen(alr, ali, xn, beta, tr, ti, sr, si);
        if (lane == 0) ES[p - 1] = beta;
        if (tr == 0.0 && ti == 0.0) continue;
        double vr = 0.0, vi = 0.0;
        if (lane < p - 1) {
            vr = xr * sr - xi * si;
            vi = xr * si + xi * sr;
        } else if (lane == p - 1) {
            vr = 1.0;
        }
        V[lane] = make_double2(vr, vi);  // zero from p on: the block-granular loops read up to the next multiple of 4
        __syncwarp();
        double qr, qi;
        row_dot<NREG, BW>(ar, ai, V, p, qr, qi);
        const double pr = tr * qr - ti * qi, pi = tr * qi + ti * qr;
        double dr = pr * vr + pi * vi;  // lanes >= p: v = 0
        double di = pr * vi - pi * vr;
        dr = warp_sum(dr);
        di = warp_sum(di);
        const double cr = -0.5 * (tr * dr - ti * di), ci = -0.5 * (tr * di + ti * dr);
        double wr = 0.0, wi = 0.0;
        if (lane < p) {
            wr = pr + cr * vr - ci * vi;
            wi = pi + cr * vi + ci * vr;
        }
        W[lane] = make_double2(wr, wi);
        __syncwarp();
        row_update<NREG, BW>(ar, ai, V, W, p, vr, vi, wr, wi);
        __syncwarp();
    }
    }
    double* Dk = D + kk * (long)ldo + off;
    double* Ek = E + kk * (long)ldo + off;
    if (n_stop > 1) {
        // ---- staged hand-over: d / e of the eliminated rows, and the remaining leading block of B stored as the packed
        // matrix A'[I][J] = B[n'-1-I][n'-1-J] (= the trailing block of A in natural order) at the start of this matrix'
        // own region; the next launch (tridiag_reg_half_kernel) finishes it with two matrices per warp ----
        __syncwarp();
        const int np = n_stop;
        for (int i = lane; i < N - np; i += 32) {
            Dk[i] = DS[N - 1 - i];
            Ek[i] = ES[N - 2 - i];
        }
        double* out = Hp + kk * mstride;
        const int ntp = itri(np);
        if (lane < np) {
            const int I = np - 1 - lane;
#pragma unroll
            for (int b = 0; b < NREG; ++b) {
                if (b >= lane && b < np) {
                    const int J = np - 1 - b;
                    out[itri(I) + J] = ar[b];
                    if (b > lane) out[ntp + itrs(I) + J] = ai[b];
                }
            }
        }
        return;
    }
    {
        double xr, xi;
        get_col<NREG>(ar, ai, 0, xr, xi);
        if (lane == 0) DS[0] = xr;
    }
    __syncwarp();

    // ---- store, undoing the index reversal: d_A[i] = d_B[N-1-i], e_A[i] = e_B[N-2-i] ----
    for (int i = lane; i < N; i += 32) {
        Dk[i] = DS[N - 1 - i];
        if (i < N - 1) Ek[i] = ES[N - 2 - i];
    }
    if (lane == 0) Ek[N - 1] = 0.0;
}

// ---- N <= 16: two matrices per warp (lanes 0-15 / 16-31), the tail of the staged reduction ----
// Below ~16 rows a step is all latency (reductions, the reflector, a few FMAs): half the lanes of a warp-per-matrix kernel
// idle and its 160+ registers leave three warps per scheduler to hide it.  Here a half-warp owns a matrix (16 columns =
// 64 registers), so every scheduler holds >= 4 warps = 8 matrices.  Both halves run the same step sequence (N is the
// same), so control flow stays warp uniform; a step whose reflector is trivial (tau = 0) is executed with v = w = 0
// instead of being skipped.
__device__ __forceinline__ double half_sum(double a) {
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off, 16);
    return a;
}

__global__ void __launch_bounds__(32, 16)
tridiag_reg_half_kernel(const double* __restrict__ Hp, int N, long mstride, long nk, double* __restrict__ D,
                        double* __restrict__ E, int ldo, int off) {
    constexpr int NREG = 16;
    __shared__ __align__(16) double smem[2][NREG * NREG + 4 * NREG + 2 * NREG];
    const int lane = threadIdx.x & 15;
    const int sub = threadIdx.x >> 4;
    long kk = 2L * blockIdx.x + sub;
    const bool live = kk < nk;
    if (!live) kk = nk - 1;  // the idle half shadows the last matrix (no stores) so that the warp stays convergent
    const int ntri = itri(N);
    const int nn = N * N;
    double* S = smem[sub];
    double2* V = reinterpret_cast<double2*>(S + NREG * NREG);
    double2* W = V + NREG;
    double* DS = reinterpret_cast<double*>(W + NREG);
    double* ES = DS + NREG;
    {
        const double* src = Hp + kk * mstride;
        for (int e = lane; e < nn; e += 16) S[e] = src[e];
        V[lane] = make_double2(0.0, 0.0);
        W[lane] = make_double2(0.0, 0.0);
    }
    __syncwarp();
    double ar[NREG], ai[NREG];
    {
        const int I = N - 1 - lane;
        const bool row_ok = lane < N;
        const double* Si = S + ntri;
        const int Ic = row_ok ? I : 0;
        const int triI = itri(Ic), trsI = itrs(Ic);
#pragma unroll
        for (int b = 0; b < NREG; ++b) {
            const int J = b < N ? N - 1 - b : 0;
            const bool own = b >= lane;
            const int ire = own ? triI + J : itri(J) + Ic;
            int iim = own ? trsI + J : itrs(J) + Ic;
            const bool diag = b == lane;
            iim = diag ? 0 : iim;
            const double re = S[ire];
            const double im = Si[iim];
            const bool ok = row_ok && b < N;
            ar[b] = ok ? re : 0.0;
            ai[b] = (ok && !diag) ? (own ? im : -im) : 0.0;
        }
    }
    __syncwarp();
    for (int p = N - 1; p >= 1; --p) {
        double xr, xi;
        get_col<NREG>(ar, ai, p, xr, xi);
        if (lane == p) DS[p] = xr;
        const double alr = __shfl_sync(0xffffffffu, xr, p - 1, 16);
        const double ali = __shfl_sync(0xffffffffu, xi, p - 1, 16);
        double xn = lane < p - 1 ? fma(xr, xr, xi * xi) : 0.0;
        xn = half_sum(xn);
        double beta, tr, ti, sr, si;
        householder_gen(alr, ali, xn, beta, tr, ti, sr, si);
        if (lane == 0) ES[p - 1] = beta;
        const bool trivial = tr == 0.0 && ti == 0.0;  // v = w = 0: the update below is a no-op
        double vr = 0.0, vi = 0.0;
        if (lane < p - 1) {
            vr = xr * sr - xi * si;
            vi = xr * si + xi * sr;
        } else if (lane == p - 1) {
            vr = trivial ? 0.0 : 1.0;
        }
        V[lane] = make_double2(vr, vi);
        __syncwarp();
        double qr, qi;
        row_dot<NREG, 4>(ar, ai, V, p, qr, qi);
        const double pr = tr * qr - ti * qi, pi = tr * qi + ti * qr;
        double dr = pr * vr + pi * vi;
        double di = pr * vi - pi * vr;
        dr = half_sum(dr);
        di = half_sum(di);
        const double cr = -0.5 * (tr * dr - ti * di), ci = -0.5 * (tr * di + ti * dr);
        double wr = 0.0, wi = 0.0;
        if (lane < p) {
            wr = pr + cr * vr - ci * vi;
            wi = pi + cr * vi + ci * vr;
        }
        W[lane] = make_double2(wr, wi);
        __syncwarp();
        row_update<NREG, 4>(ar, ai, V, W, p, vr, vi, wr, wi);
        __syncwarp();
    }
    {
        double xr, xi;
        get_col<NREG>(ar, ai, 0, xr, xi);
        if (lane == 0) DS[0] = xr;
    }
    __syncwarp();
    if (!live) return;
    double* Dk = D + kk * (long)ldo + off;
    double* Ek = E + kk * (long)ldo + off;
    if (lane < N) {
        Dk[lane] = DS[N - 1 - lane];
        Ek[lane] = lane < N - 1 ? ES[N - 2 - lane] : 0.0;
    }
}

cudaError_t launch_reg_half(int n, const double* Hp, long nk, double* D, double* E, cudaStream_t st, long mstride, int ldo,
                            int off) {
    if (nk <= 0) return cudaSuccess;
    const long blocks = (nk + 1) / 2;
    if (blocks > 2147483647L) return cudaErrorInvalidConfiguration;
    tridiag_reg_half_kernel<<<(unsigned)blocks, 32, 0, st>>>(Hp, n, mstride, nk, D, E, ldo, off);
    return cudaGetLastError();
}

template <int NREG, int XMAX, int BW, int OCC, bool LA>
cudaError_t launch_reg_bw(int n, double* Hp, long nk, double* D, double* E, cudaStream_t st, long mstride, int ldo,
                          int off, int n_stop) {
    constexpr int NV = NREG + XMAX < 32 ? 32 : NREG + XMAX;
    const size_t smem = (size_t)(((n * n + 1) & ~1)) * 8 + (size_t)(3 * NV + XMAX * XMAX + XMAX * 32) * 16;
    cudaError_t err =
        cudaFuncSetAttribute(tridiag_reg_kernel<NREG, XMAX, BW, OCC, LA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    if (nk <= 0) return cudaSuccess;
    if (nk > 2147483647L) return cudaErrorInvalidConfiguration;
    tridiag_reg_kernel<NREG, XMAX, BW, OCC, LA><<<(unsigned)nk, 32, smem, st>>>(Hp, n, mstride, nk, D, E, ldo, off, n_stop);
    return cudaGetLastError();
}

template <int NREG, int XMAX>
cudaError_t launch_reg(int n, double* Hp, long nk, double* D, double* E, cudaStream_t st, long mstride, int ldo,
                       int off, int bw, int n_stop) {
    // Default build per size class, measured on B200 (gpurun_out/r02x_sweep.log, ms per 1000 matrices, plain / lookahead
    // at 168 registers / lookahead at 255 registers): N = 32: 0.027 / 0.033 / 0.032, 36: 0.038 / 0.045 / 0.041,
    // 40: 0.067 / 0.073 / 0.054, 48: 0.137 / 0.143 / 0.123.  The lookahead needs ~20 more live registers; at the
    // 168-register budget (12 warps per SM) they spill into the bulk loops and it loses, with 255 registers (8 warps per
    // SM) it wins where a step is long enough -- the classes with a large shared-memory corner (N > 36).
    // bw: tuning hooks.  8 = lookahead + 255 registers, 1 = plain loop at the default budget, 2 = lookahead at the
    // default budget (A/B tests)
    constexpr int OCC8 = NREG == 32 ? 8 : 0;
    if (bw == 8) return launch_reg_bw<NREG, XMAX, 4, OCC8, true>(n, Hp, nk, D, E, st, mstride, ldo, off, n_stop);
    if (bw == 1) return launch_reg_bw<NREG, XMAX, 4, 0, false>(n, Hp, nk, D, E, st, mstride, ldo, off, n_stop);
    if (bw == 2) return launch_reg_bw<NREG, XMAX, 4, 0, true>(n, Hp, nk, D, E, st, mstride, ldo, off, n_stop);
    // (9 or 10 warps per SM are no middle ground: registers are per scheduler, so any count above 8 puts three warps
    //  on one scheduler and caps the build at 168 registers again)
    if constexpr (XMAX >= 8) return launch_reg_bw<NREG, XMAX, 4, OCC8, true>(n, Hp, nk, D, E, st, mstride, ldo, off, n_stop);
    return launch_reg_bw<NREG, XMAX, 4, 0, false>(n, Hp, nk, D, E, st, mstride, ldo, off, n_stop);
}

}  // namespace

bool tridiag_reg_fits(int n) { return n >= 2 && n <= kTridiagRegMaxN; }

// Staged reduction (stop >= 2): the warp-per-matrix kernels hand a shrinking trailing block from launch to launch --
//   n > 28:  n -> mid (24) with the 168-register build (12 warps per SM),
//   n > 16:  -> stop (16) with the 24-column build (128 registers, 16 warps per SM),
//   then the two-matrices-per-warp kernel finishes the stop x stop block.
// Late steps are latency bound (two warp reductions and the reflector against a few hundred FMAs), so a stage trades
// registers for resident warps as soon as the block allows.  mid = 0 skips the middle stage (the default: measured on
// B200 the extra hand-over costs more than the fourth warp per scheduler gains -- N = 36: 0.038 ms per 1000 matrices
// without, 0.043 with mid = 24, gpurun_out/r02o_sweep.log), stop = 0 is a single launch.
// Stage sizes follow from n only, so results never depend on the batch.
cudaError_t launch_tridiag_reg(int n, double* Hp, long nk, double* D, double* E, cudaStream_t st, long mstride,
                               int ldo, int off, int bw, int stop, int mid) {
    if (mstride == 0) mstride = (long)n * n;
    if (ldo == 0) ldo = n;
    if (stop > 16) stop = 16;
    if (stop < 2) stop = 0;
    if (mid > 24 || mid <= stop) mid = 0;
    int cur = n;
    while (true) {
        if (stop && cur <= 16) return launch_reg_half(cur, Hp, nk, D, E, st, mstride, ldo, off);
        int next = 0;
        if (stop && cur >= stop + 4) next = (mid && cur >= mid + 4) ? mid : stop;
        cudaError_t err = cudaErrorInvalidValue;
        if (cur <= 12) err = launch_reg<12, 0>(cur, Hp, nk, D, E, st, mstride, ldo, off, bw, next);
        else if (cur <= 16) err = launch_reg<16, 0>(cur, Hp, nk, D, E, st, mstride, ldo, off, bw, next);
        else if (cur <= 20) err = launch_reg<20, 0>(cur, Hp, nk, D, E, st, mstride, ldo, off, bw, next);
        else if (cur <= 24) err = launch_reg<24, 0>(cur, Hp, nk, D, E, st, mstride, ldo, off, bw, next);
        else if (cur <= 28) err = launch_reg<28, 0>(cur, Hp, nk, D, E, st, mstride, ldo, off, bw, next);
        else if (cur <= 32) err = launch_reg<32, 0>(cur, Hp, nk, D, E, st, mstride, ldo, off, bw, next);
        else if (cur <= 36) err = launch_reg<32, 4>(cur, Hp, nk, D, E, st, mstride, ldo, off, bw, next);
        else if (cur <= 40) err = launch_reg<32, 8>(cur, Hp, nk, D, E, st, mstride, ldo, off, bw, next);
        else if (cur <= 48) err = launch_reg<32, 16>(cur, Hp, nk, D, E, st, mstride, ldo, off, bw, next);
        if (err != cudaSuccess || next == 0) return err;
        off += cur - next;
        cur = next;
    }
}

}  // namespace tbk
