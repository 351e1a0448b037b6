// eig_tridiag_panel.cu -- blocked (panel) Hermitian -> tridiagonal reduction for matrices that live in L2 / HBM.
//
// Same role as eig_tridiag.cu (first half of the replacement for the per-k scipy.linalg.eigvalsh loop of
// Model.eigenval, reference src/tbmodels/_tb_model.py:1148-1149; LAPACK zheevr JOBZ='N', UPLO='L'), for sizes whose
// matrix does not fit in shared memory (N > 164; C4: N = 512).  The unblocked kernel streams the trailing matrix three
// times per Householder step (product: read; rank-2 update: read + write).  Here the update is deferred over panels of
// NB = 8 columns (the textbook latrd/her2k organisation, restated):
//   for every column c of the panel
//       a      = A[c:, c] - V W[c,:]^H - W V[c,:]^H                    (column brought up to date, O(N NB))
//       v, tau = reflector(a)                                           (same contract as tbk_math.cuh householder_gen)
//       p      = A22 v - V (W^H v) - W (V^H v)                          (A22 = the STORED trailing matrix, read once)
//       w      = tau p - (tau/2)((tau p)^H v) v
//   A22 -= V W^H + W V^H   once per panel, on the FP64 tensor cores:
//       Re(A) -= [vr vi wr wi].[wr wi vr vi]^T,  Im(A) -= [vi -vr wi -wr].[wr wi vr vi]^T   -> per 8x8 block and plane
//       NB mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) accumulating in place on fragments loaded from / stored to the packed
//       planes in global memory.
// The matrix is therefore read (1 + 2/NB) times per step instead of 3.  One CTA per matrix; the Hermitian product is
// organised as coalesced row sweeps over the lower triangle exactly like tridiag_big_kernel (row part reduced per row,
// column part scattered into per-lane register accumulators and combined across warps in a fixed order: no atomics,
// bits do not depend on the batch).
#include <cstdio>
#include <cstdlib>

#include <algorithm>

#include "tbk_kernels.h"
#include "tbk_math.cuh"

namespace tbk {

namespace {

constexpr int NB = 8;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ int itri(int i) { return (i * (i + 1)) >> 1; }  // 32-bit: offsets inside one matrix
__device__ __forceinline__ int itrs(int i) { return (i * (i - 1)) >> 1; }

// Pull columns [c0, c0 + len) of row I of both planes into L2 (bulk prefetch: 16-byte granularity, no registers).
__device__ __forceinline__ void prefetch_row(const double* Ar, const double* Ai, int I, int c0, int len) {
    if (len <= 0) return;
    const unsigned long long b0 = reinterpret_cast<unsigned long long>(Ar + itri(I) + c0);
    const unsigned long long b1 = reinterpret_cast<unsigned long long>(Ai + itrs(I) + c0);
    const unsigned long long s0 = b0 & ~15ull, s1 = b1 & ~15ull;
    const unsigned n0 = (unsigned)(((b0 + 8ull * len + 15ull) & ~15ull) - s0);
    const unsigned n1 = (unsigned)(((b1 + 8ull * len + 15ull) & ~15ull) - s1);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(s0), "r"(n0) : "memory");
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(s1), "r"(n1) : "memory");
}

// 8-byte global load at a compile-time byte offset from a materialised row pointer (keeps the address arithmetic of
// the unrolled row sweep down to one pointer per row and plane).
template <int OFF>
__device__ __forceinline__ double ldg_off(unsigned long long gptr) {
    double v;
    asm volatile("ld.global.f64 %0, [%1+%2];" : "=d"(v) : "l"(gptr), "n"(OFF) : "memory");
    return v;
}

// Four LPR-column chunks of one row (both planes) starting at the row pointers; rem = columns left for this lane
// (chunk q is inside the row iff LPR q < rem).
template <int LPR>
__device__ __forceinline__ void load4(unsigned long long pr, unsigned long long pi, int rem, double (&zr)[4],
                                      double (&zi)[4]) {
    zr[0] = zr[1] = zr[2] = zr[3] = 0.0;
    zi[0] = zi[1] = zi[2] = zi[3] = 0.0;
    if (0 < rem) {
        zr[0] = ldg_off<0>(pr);
        zi[0] = ldg_off<0>(pi);
    }
    if (LPR < rem) {
        zr[1] = ldg_off<LPR * 8>(pr);
        zi[1] = ldg_off<LPR * 8>(pi);
    }
    if (2 * LPR < rem) {
        zr[2] = ldg_off<LPR * 16>(pr);
        zi[2] = ldg_off<LPR * 16>(pi);
    }
    if (3 * LPR < rem) {
        zr[3] = ldg_off<LPR * 24>(pr);
        zi[3] = ldg_off<LPR * 24>(pi);
    }
}

// Sum (a, b) over the CTA; every thread gets the result.  Fixed order; double-buffered: one barrier per call.
template <int WARPS>
__device__ __forceinline__ void cta_sum2(double& a, double& b, double* red, int tid, int& parity) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
    }
    double* buf = red + parity * 2 * WARPS;
    parity ^= 1;
    if ((tid & 31) == 0) {
        buf[2 * (tid >> 5)] = a;
        buf[2 * (tid >> 5) + 1] = b;
    }
    __syncthreads();
    a = 0.0;
    b = 0.0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
        a += buf[2 * w];
        b += buf[2 * w + 1];
    }
}

__host__ __device__ inline size_t panel_smem_doubles(int n, int warps, int maxc) {
    const int slots = warps / 4;
    const int nsw = 1;
    (void)maxc;
    return (size_t)4 * NB * n + (size_t)2 * n * (2 + nsw + slots) + 2 * (size_t)n + 4 * NB + 4 * warps + 4;
}

template <int THREADS, int MAXC, int MINB, int LPR>
__global__ void __launch_bounds__(THREADS, MINB)
tridiag_panel_kernel(double* __restrict__ Hp, int N, long nk, double* __restrict__ D, double* __restrict__ E,
                     int pfd /* L2 prefetch distance of the Hermitian product, in warp trips; 0 = off */,
                     int n_stop /* staged: hand the trailing block over once at most n_stop rows are left; 0 = reduce fully */) {
    constexpr int WARPS = THREADS / 32;
    constexpr int SLOTS = WARPS / 4;
    constexpr int NSW = 1;  // (row-part planes)
    extern __shared__ __align__(16) double smp[];
    double* X = smp;                                              // [NB][N][4] = (vr, vi, wr, wi) of panel column p, row r
    double2* V = reinterpret_cast<double2*>(X + (size_t)4 * NB * N);  // [N] current reflector, index i = row - r0
    double2* P = V + N;                                           // [N] tau * p
    double2* S = P + N;                                           // [NSW][N] row-part sums, one plane per column sweep
    double2* QW = S + (size_t)NSW * N;                            // [SLOTS][N] column-part partial sums
    double* ds = reinterpret_cast<double*>(QW + (size_t)SLOTS * N);
    double* es = ds + N;
    double2* Y = reinterpret_cast<double2*>(es + N);              // [2 NB]: Y[2p] = W_p^H v, Y[2p+1] = V_p^H v
    double* red = reinterpret_cast<double*>(Y + 2 * NB);          // [2][2 WARPS]
    double* misc = red + 4 * WARPS;                               // alpha (re, im)

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const long kk = blockIdx.x;
    if (kk >= nk) return;
    double* Ar = Hp + kk * (long)N * N;
    double* Ai = Ar + tri(N);
    int parity = 0;
#ifdef TBK_PANEL_TIMING  // debug build: cycles per phase of CTA 0 (thread 0; phase 2 additionally as the max over warps)
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64(), twarp = 0;
#define TICK(i)                                  \
    do {                                         \
        const long long now_ = clock64();        \
        tacc[i] += now_ - tlast;                 \
        tlast = now_;                            \
    } while (0)
#define TICKW(i)                                 \
    do {                                         \
        const long long now_ = clock64();        \
        twarp += now_ - tlast;                   \
        tacc[i] += now_ - tlast;                 \
        tlast = now_;                            \
    } while (0)
#else
#define TICK(i)
#define TICKW(i)
#endif

    int k_done = N;  // rows / columns eliminated when the panel loop ends
    for (int k0 = 0; k0 < N; k0 += NB) {
        if (n_stop > 0 && k0 > 0 && N - k0 <= n_stop) {  // panel boundary: the trailing block is fully up to date
            k_done = k0;
            break;
        }
        const int nb = (N - k0 < NB) ? (N - k0) : NB;
        for (int j = 0; j < nb; ++j) {
            const int c = k0 + j;
            const int m = N - 1 - c;  // size of the trailing block below / right of column c
            const int r0 = c + 1;
            // --- (1) column c of the matrix, brought up to date with the panel's previous reflectors ---
            double xn = 0.0, dummy = 0.0;
            for (int i = tid; i <= m; i += THREADS) {  // row r = c + i; i = 0 is the diagonal
                const long r = c + i;
                double ax = Ar[tri(r) + c];
                double ay = (i > 0) ? Ai[trs(r) + c] : 0.0;
                for (int p = 0; p < j; ++p) {
                    const double2 xv = *reinterpret_cast<const double2*>(X + ((size_t)p * N + r) * 4);
                    const double2 xw = *reinterpret_cast<const double2*>(X + ((size_t)p * N + r) * 4 + 2);
                    const double2 cv = *reinterpret_cast<const double2*>(X + ((size_t)p * N + c) * 4);
                    const double2 cw = *reinterpret_cast<const double2*>(X + ((size_t)p * N + c) * 4 + 2);
                    // a -= V[r,p] conj(W[c,p]) + W[r,p] conj(V[c,p])
                    ax = fma(-xv.x, cw.x, fma(-xv.y, cw.y, fma(-xw.x, cv.x, fma(-xw.y, cv.y, ax))));
                    ay = fma(-xv.y, cw.x, fma(xv.x, cw.y, fma(-xw.y, cv.x, fma(xw.x, cv.y, ay))));
                }
                if (i == 0) {
                    ds[c] = ax;
                } else {
                    V[i - 1] = make_double2(ax, ay);  // raw column, scaled below
                    if (i == 1) {
                        misc[0] = ax;
                        misc[1] = ay;
                    } else {
                        xn = fma(ax, ax, fma(ay, ay, xn));
                    }
                }
            }
            if (m == 0) {  // last column: only its diagonal entry
                if (tid == 0) es[c] = 0.0;
                __syncthreads();
                continue;
            }
            cta_sum2<WARPS>(xn, dummy, red, tid, parity);
            TICK(0);  // column gather / update + norm reduction
            double beta, tr, ti, sr, si;
            householder_gen(misc[0], misc[1], xn, beta, tr, ti, sr, si);
            // --- (2) v = [1; scale * x]; kept in V (this step) and in X[j] (panel) ---
            for (int i = tid; i < m; i += THREADS) {
                double2 v = make_double2(1.0, 0.0);  // (alpha was read from misc, nobody reads the raw V[0])
                if (i > 0) {
                    const double2 x = V[i];
                    v = make_double2(x.x * sr - x.y * si, x.x * si + x.y * sr);
                }
                V[i] = v;
                *reinterpret_cast<double2*>(X + ((size_t)j * N + r0 + i) * 4) = v;
            }
            if (tid == 0) es[c] = beta;
            __syncthreads();
            TICK(1);  // reflector, v
            // --- (2b) y1[p] = W_p^H v, y2[p] = V_p^H v for the previous panel columns: one warp per dot product ---
            for (int q = w; q < 2 * j; q += WARPS) {
                const int p = q >> 1;
                const int off = (q & 1) ? 0 : 2;  // even q: W (y1), odd q: V (y2)
                double yr = 0.0, yi = 0.0;
                for (int i = lane; i < m; i += 32) {
                    const double2 x = *reinterpret_cast<const double2*>(X + ((size_t)p * N + r0 + i) * 4 + off);
                    const double2 v = V[i];
                    yr = fma(x.x, v.x, fma(x.y, v.y, yr));
                    yi = fma(x.x, v.y, fma(-x.y, v.x, yi));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    yr += __shfl_xor_sync(0xffffffffu, yr, o);
                    yi += __shfl_xor_sync(0xffffffffu, yi, o);
                }
                if (lane == 0) Y[q] = make_double2(yr, yi);
            }
            // --- (3) Hermitian product with the stored trailing block: row sweeps over the lower triangle.
            //     LPR lanes share a row (a warp works on 32 / LPR adjacent rows per trip); lane l of a row group owns
            //     columns l, l + LPR, ... with one register accumulator per LPR-column chunk for the column part.
            //     The rows of the next trip are pulled into L2 by a bulk prefetch. ---
            constexpr int RPW = 32 / LPR;
            const int l = lane % LPR, sub = lane / LPR;
            double2 qacc[MAXC];
#pragma unroll
            for (int cc = 0; cc < MAXC; ++cc) qacc[cc] = make_double2(0.0, 0.0);
            if (l == 0) {
                for (int t = 0; t < pfd; ++t) {
                    const int a1 = RPW * (w + t * WARPS) + sub;
                    if (a1 < m) prefetch_row(Ar, Ai, r0 + a1, r0, a1);
                }
            }
            const double2* pv = V + l;
            for (int a0 = RPW * w; a0 < m; a0 += RPW * WARPS) {
                const int a = a0 + sub;
                if (l == 0 && pfd > 0) {
                    const int a2 = a + pfd * RPW * WARPS;
                    if (a2 < m) prefetch_row(Ar, Ai, r0 + a2, r0, a2);
                }
                const bool on = a < m;
                const int ac = on ? a : a0;  // row groups past the end shadow row a0 with every column masked off
                const int I = r0 + ac;
                const unsigned long long pr = (unsigned long long)__cvta_generic_to_global(Ar + (itri(I) + r0 + l));
                const unsigned long long pi = (unsigned long long)__cvta_generic_to_global(Ai + (itrs(I) + r0 + l));
                const int rem = on ? a - l : 0;    // column b = LPR q + l lies inside the row  <=>  LPR q < rem
                const double2 va = V[ac];
                const double dg = Ar[itri(I) + I];  // real diagonal (same address in every lane of the row group)
                const int alast = (a0 + RPW - 1 < m) ? a0 + RPW - 1 : m - 1;  // longest row of the trip (warp-uniform)
                double sumr = 0.0, sumi = 0.0;
#pragma unroll
                for (int c0 = 0; c0 < MAXC; c0 += 4) {
                    if (c0 * LPR < alast) {  // warp-uniform; four chunks per batch: 8 loads in flight per lane
                        double zr[4], zi[4];
                        load4<LPR>(pr + c0 * LPR * 8, pi + c0 * LPR * 8, rem - c0 * LPR, zr, zi);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (c0 + q < MAXC && (c0 + q) * LPR < rem) {
                                const double2 vb = pv[(c0 + q) * LPR];
                                sumr = fma(zr[q], vb.x, fma(-zi[q], vb.y, sumr));
                                sumi = fma(zr[q], vb.y, fma(zi[q], vb.x, sumi));
                                qacc[(c0 + q) % MAXC].x = fma(zr[q], va.x, fma(zi[q], va.y, qacc[(c0 + q) % MAXC].x));
                                qacc[(c0 + q) % MAXC].y = fma(zr[q], va.y, fma(-zi[q], va.x, qacc[(c0 + q) % MAXC].y));
                            }
                        }
                    }
                }
#pragma unroll
                for (int o = LPR / 2; o > 0; o >>= 1) {
                    sumr += __shfl_xor_sync(0xffffffffu, sumr, o);
                    sumi += __shfl_xor_sync(0xffffffffu, sumi, o);
                }
                if (l == 0 && on) S[a] = make_double2(fma(dg, va.x, sumr), fma(dg, va.y, sumi));
            }
            if (RPW > 1) {  // the row groups of a warp hold partial column sums for the same columns
#pragma unroll
                for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
                    for (int cc = 0; cc < MAXC; ++cc) {
                        qacc[cc].x += __shfl_xor_sync(0xffffffffu, qacc[cc].x, o);
                        qacc[cc].y += __shfl_xor_sync(0xffffffffu, qacc[cc].y, o);
                    }
                }
            }
            TICKW(2);  // dots + own rows of the Hermitian product (per warp, before the barrier)
            // column parts: four rounds, warp w adds into slot w / 4 in round w % 4 (fixed order)
#pragma unroll
            for (int round = 0; round < 4; ++round) {
                if ((w & 3) == round && sub == 0) {
                    double2* slot = QW + (size_t)(w >> 2) * N + l;
#pragma unroll
                    for (int cc = 0; cc < MAXC; ++cc) {
                        if (cc * LPR + l < m) {
                            if (round == 0) {
                                slot[cc * LPR] = qacc[cc];
                            } else {
                                double2 t2 = slot[cc * LPR];
                                t2.x += qacc[cc].x;
                                t2.y += qacc[cc].y;
                                slot[cc * LPR] = t2;
                            }
                        }
                    }
                }
                __syncthreads();
            }
            TICK(3);  // QW rounds (includes waiting for the slowest warp of the product)
            // --- (4) p = A22 v - V y1 - W y2;  tau p;  dot = (tau p)^H v ---
            double dr = 0.0, di = 0.0;
            for (int a = tid; a < m; a += THREADS) {
                double2 q = S[a];
#pragma unroll
                for (int sl = 0; sl < SLOTS; ++sl) {
                    const double2 t2 = QW[(size_t)sl * N + a];
                    q.x += t2.x;
                    q.y += t2.y;
                }
                for (int p = 0; p < j; ++p) {
                    const double2 xv = *reinterpret_cast<const double2*>(X + ((size_t)p * N + r0 + a) * 4);
                    const double2 xw = *reinterpret_cast<const double2*>(X + ((size_t)p * N + r0 + a) * 4 + 2);
                    const double2 y1 = Y[2 * p], y2 = Y[2 * p + 1];
                    q.x = fma(-xv.x, y1.x, fma(xv.y, y1.y, fma(-xw.x, y2.x, fma(xw.y, y2.y, q.x))));
                    q.y = fma(-xv.x, y1.y, fma(-xv.y, y1.x, fma(-xw.x, y2.y, fma(-xw.y, y2.x, q.y))));
                }
                const double pr = tr * q.x - ti * q.y;
                const double pi = tr * q.y + ti * q.x;
                P[a] = make_double2(pr, pi);
                const double2 va = V[a];
                dr += pr * va.x + pi * va.y;
                di += pr * va.y - pi * va.x;
            }
            cta_sum2<WARPS>(dr, di, red, tid, parity);
            TICK(4);  // p, tau p, dot
            const double alr = -0.5 * (tr * dr - ti * di);
            const double ali = -0.5 * (tr * di + ti * dr);
            for (int a = tid; a < m; a += THREADS) {
                const double2 v = V[a];
                const double2 pq = P[a];
                *reinterpret_cast<double2*>(X + ((size_t)j * N + r0 + a) * 4 + 2) =
                    make_double2(pq.x + alr * v.x - ali * v.y, pq.y + alr * v.y + ali * v.x);
            }
            __syncthreads();
            TICK(5);  // w -> X
        }
        // --- trailing update A22 -= V W^H + W V^H on the FP64 tensor cores (rows / columns >= k0 + NB) ---
        const int rt = k0 + nb;
        if (rt < N) {  // then nb == NB and rt is a multiple of 8
            const int I0 = rt >> 3;
            const int NBk = (N + 7) >> 3;
            const int g = lane >> 2, tq = lane & 3;
            const double sgn = (tq & 1) ? 1.0 : -1.0;
            int I = I0, s = w;
            for (;;) {
                int ns = ((I - I0) >> 2) + 1;  // 8 x 32 strips in block row I
                while (I < NBk && s >= ns) {
                    s -= ns;
                    ++I;
                    ns = ((I - I0) >> 2) + 1;
                }
                if (I >= NBk) break;
                const int row = 8 * I + g;
                const bool row_ok = row < N;
                const int rowc = row_ok ? row : N - 1;
                double* pre = Ar + tri((long)rowc);
                double* pim = Ai + trs((long)rowc);
                const int J0 = I0 + 4 * s;
                double cre[4][2], cim[4][2];
#pragma unroll
                for (int jb = 0; jb < 4; ++jb) {
                    const int col = 8 * (J0 + jb) + 2 * tq;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        cre[jb][h] = (row_ok && col + h <= row) ? pre[col + h] : 0.0;
                        cim[jb][h] = (row_ok && col + h < row) ? pim[col + h] : 0.0;
                    }
                }
                int colg[4];
#pragma unroll
                for (int jb = 0; jb < 4; ++jb) {
                    colg[jb] = 8 * (J0 + jb) + g;
                    if (colg[jb] >= N) colg[jb] = N - 1;  // (blocks right of the diagonal block are never stored)
                }
#pragma unroll
                for (int p = 0; p < NB; ++p) {
                    const double* Xp = X + (size_t)p * N * 4;
                    const double xr = -Xp[rowc * 4 + tq];             // -(vr, vi, wr, wi)[tq]
                    const double xi = sgn * Xp[rowc * 4 + (tq ^ 1)];  // (-vi, vr, -wi, wr)[tq]
#pragma unroll
                    for (int jb = 0; jb < 4; ++jb) {
                        if (J0 + jb <= I) {  // warp-uniform
                            const double y = Xp[colg[jb] * 4 + (tq ^ 2)];  // (wr, wi, vr, vi)[tq]
                            dmma884(cre[jb][0], cre[jb][1], xr, y);
                            dmma884(cim[jb][0], cim[jb][1], xi, y);
                        }
                    }
                }
#pragma unroll
                for (int jb = 0; jb < 4; ++jb) {
                    const int col = 8 * (J0 + jb) + 2 * tq;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (row_ok && col + h <= row) pre[col + h] = cre[jb][h];
                        if (row_ok && col + h < row) pim[col + h] = cim[jb][h];
                    }
                }
                s += WARPS;
            }
            __syncthreads();
            TICK(6);  // her2k
        }
    }
    __syncthreads();
#ifdef TBK_PANEL_TIMING
    if (kk == 0 && lane == 0) printf("panel timing N=%d warp %2d: own product rows %.3f Mcyc\n", N, w, twarp * 1e-6);
    if (kk == 0 && tid == 0)
        printf("panel timing N=%d phases: colupd %.3f refl %.3f product(w0) %.3f rounds+wait %.3f p/dot %.3f w %.3f her2k %.3f Mcyc\n",
               N, tacc[0] * 1e-6, tacc[1] * 1e-6, tacc[2] * 1e-6, tacc[3] * 1e-6, tacc[4] * 1e-6, tacc[5] * 1e-6,
               tacc[6] * 1e-6);
#endif
    if (k_done < N) {
        // ---- staged hand-over (launch_tridiag): d / e of the eliminated rows, and the trailing np x np block compacted
        // into packed form at the head of this matrix' slot, where the shared-memory / register kernels continue ----
        const int np = N - k_done;
        for (int i = tid; i < k_done; i += THREADS) {
            D[kk * N + i] = ds[i];
            E[kk * N + i] = es[i];
        }
        // In-place compaction, no staging buffer: every target address lies below every source address that is still
        // unread when (1) the real plane is compacted before the imaginary one (real targets < tri(np) <= tri(N) <= every
        // imaginary source; imaginary targets may overlap real SOURCES, which are consumed by then) and (2) rows go in
        // ascending order (target row I ends at tri(I) + I < tri(k + I) + k, the start of its own source row, and source
        // rows ascend).  WARPS rows per round, read into registers, barrier, write, barrier.
        constexpr int RMAX = 6;  // ceil(192 / 32): rows of the handed-over block have at most 6 elements per lane
        const int ntp = (np * (np + 1)) >> 1;
        for (int plane = 0; plane < 2; ++plane) {
            const double* src_plane = plane ? Ai : Ar;
            for (int I0 = 0; I0 < np; I0 += WARPS) {
                const int I = I0 + w;
                const int len = I < np ? (plane ? I : I + 1) : 0;  // row I: I + 1 real, I imaginary entries
                const long srow = plane ? trs(k_done + I) + k_done : tri(k_done + I) + k_done;
                double buf[RMAX];
#pragma unroll
                for (int q = 0; q < RMAX; ++q) {
                    const int J = lane + 32 * q;
                    buf[q] = J < len ? src_plane[srow + J] : 0.0;
                }
                __syncthreads();
                const int drow = plane ? ntp + ((I * (I - 1)) >> 1) : (I * (I + 1)) >> 1;
#pragma unroll
                for (int q = 0; q < RMAX; ++q) {
                    const int J = lane + 32 * q;
                    if (J < len) Ar[drow + J] = buf[q];
                }
                __syncthreads();
            }
        }
        return;
    }
    for (int i = tid; i < N; i += THREADS) {
        D[kk * N + i] = ds[i];
        E[kk * N + i] = (i < N - 1) ? es[i] : 0.0;
    }
}

template <int THREADS, int MAXC, int MINB, int LPR>
cudaError_t launch_panel_t(int n, double* Hp, long nk, double* D, double* E, cudaStream_t st, const Tuning& tune, int n_stop) {
    const size_t smem = panel_smem_doubles(n, THREADS / 32, MAXC) * 8;
    if (n_stop > 192) return cudaErrorInvalidValue;  // the hand-over copies rows of at most 192 entries
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t err = cudaFuncSetAttribute(tridiag_panel_kernel<THREADS, MAXC, MINB, LPR>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    if (nk <= 0) return cudaSuccess;
    if (nk > 2147483647L) return cudaErrorInvalidConfiguration;
    int pfd = tune.panel_pfd;  // measured (C4, ms per 1184 matrices): off 167, 1 trip 154, 2 trips 156, 4 trips 167, 8 trips 175
    if (pfd < 0 || pfd > 8) pfd = 1;
    tridiag_panel_kernel<THREADS, MAXC, MINB, LPR><<<(unsigned)nk, THREADS, smem, st>>>(Hp, n, nk, D, E, pfd, n_stop);
    return cudaGetLastError();
}

}  // namespace

bool tridiag_panel_fits(int n) {  // (the 16-warp configuration has the largest shared-memory footprint)

    return n >= 2 && n <= 640 && panel_smem_doubles(n, 16, n <= 512 ? 16 : 20) * 8 <= 227 * 1024;
}

// Rows left when the blocked kernel hands over with the limit n_stop: the first panel boundary with at most n_stop rows.
int tridiag_panel_handover(int n, int n_stop) {
    if (n_stop <= 0 || n <= n_stop) return 0;
    const int panels = (n - n_stop + NB - 1) / NB;
    return n - panels * NB;
}

cudaError_t launch_tridiag_panel(int n, double* Hp, long nk, double* D, double* E, cudaStream_t st, const Tuning& tune,
                                 int n_stop) {
    const int t = tune.panel_t, lpr = tune.panel_lpr;  // tuning hooks: threads per matrix, lanes per row of the product
    // template arguments: threads, column chunks (n <= chunks * lanes per row), min CTAs per SM, lanes per row.
    // Measured on B200, tridiagonalisation ms per 1000 matrices:
    //   N = 128: 16 lanes/row 3.11 (256 thr, 4 CTAs/SM), 32 lanes/row 3.32, 128 thr 3.69, shared-memory kernel 3.65
    //   N = 200: 16 lanes/row 9.39, 32 lanes/row 10.47 (512 thr: 12.6);  N = 256: 17.6 / 19.2 (512 thr: 20.8)
    if (n <= 128) {
        if (lpr == 32) {
            if (t == 128) return launch_panel_t<128, 4, 8, 32>(n, Hp, nk, D, E, st, tune, n_stop);
            if (t == 512) return launch_panel_t<512, 4, 2, 32>(n, Hp, nk, D, E, st, tune, n_stop);
            return launch_panel_t<256, 4, 4, 32>(n, Hp, nk, D, E, st, tune, n_stop);
        }
        if (t == 128) return launch_panel_t<128, 8, 8, 16>(n, Hp, nk, D, E, st, tune, n_stop);
        return launch_panel_t<256, 8, 4, 16>(n, Hp, nk, D, E, st, tune, n_stop);
    }
    if (n <= 256) {
        if (lpr == 32) {
            if (t == 512) return launch_panel_t<512, 8, 1, 32>(n, Hp, nk, D, E, st, tune, n_stop);
            return launch_panel_t<256, 8, 2, 32>(n, Hp, nk, D, E, st, tune, n_stop);
        }
        return launch_panel_t<256, 16, 2, 16>(n, Hp, nk, D, E, st, tune, n_stop);
    }
    if (n <= 512) return launch_panel_t<512, 16, 1, 32>(n, Hp, nk, D, E, st, tune, n_stop);
    return launch_panel_t<512, 20, 1, 32>(n, Hp, nk, D, E, st, tune, n_stop);
}

}  // namespace tbk
