// eig_band.cu -- two-stage Hermitian -> tridiagonal reduction (N >= 128).
//
// Same role as eig_tridiag.cu / eig_tridiag_panel.cu (first half of the replacement for the per-k scipy.linalg.eigvalsh loop of
// Model.eigenval, reference src/tbmodels/_tb_model.py:1148-1149; LAPACK zheevr JOBZ='N', UPLO='L'), different
// formulation.  The one-stage reduction needs one Hermitian matrix-vector product with the whole trailing matrix per
// column -- half of its flops are BLAS-2 and stream the matrix from L2 / HBM once per column (394 MB per 512 x 512
// matrix, profiles/r01l_ncu_summary.txt).  Here:
//
//   stage 1, band_reduce_kernel (one CTA per matrix): Hermitian -> band of half bandwidth 8.  Per panel of 8 columns
//       E      = A[r0:, c0:c0+8]                       (the block below the band, r0 = c0 + 8) -> shared memory
//       E      = Q [R; 0],  Q = H_0 .. H_7 = I - V T V^H   (Householder QR in shared memory, ONE block reduction per
//                                                           column: norm and the 7 - k dot products together)
//       Y      = A22 V                                 (A22 = A[r0:, r0:]; FP64 tensor cores; the stored triangle is read
//                                                        as is and as its conjugate transpose)
//       G, M   = V^H V, V^H Y                          (8 x 8, tensor cores) ;  T from G and tau (zlarft recurrence)
//       Z      = Y T - 1/2 V (T^H M T)
//       A22   -= V Z^H + Z V^H                         (tensor cores, the her2k of the blocked kernel)
//     R and the diagonal block go to the band array band[c][d] = A[c + d, c], d = 0 .. 15 (8 diagonals of room for the
//     bulges of stage 2).  The stored triangle is read three times and written once per EIGHT columns instead of once
//     per column.
//   stage 2, band_chase_pipe_kernel (one matrix per warp): band -> tridiagonal by bulge chasing with length-8 reflectors
//     (the Householder form of the Schwarz / Murata-Horikoshi algorithm): sweep j annihilates column j below the
//     sub-diagonal, the 8 x 8 bulge this opens one block further down is chased off the end of the band block by block.
//     An 8-lane group runs one sweep: lane l owns row l of the diagonal block (the full Hermitian row) and row l of the
//     block below, every product along a row is local, the few sums across lanes are 3-level shuffles.  The warp's four
//     groups run four consecutive sweeps two steps apart, so the band is streamed from HBM once per four sweeps.
//     (band_chase_kernel: the first form, four matrices per warp and one sweep at a time, kept behind TBK_BAND_CHASE=1.)
//
// Both stages use fixed-order reductions only: results do not depend on the batch.  oracle/twostage_hetrd.py is a numpy
// walk-through of exactly these steps, index conventions and the pipelined schedule (pinned against LAPACK in the CPU suite).
#include <cstdio>
#include <cstdlib>

#include <algorithm>

#include "tbk_kernels.h"
#include "tbk_math.cuh"

namespace tbk {

namespace {

constexpr int BB = 8;    // half bandwidth = panel width
constexpr int BWD = 16;  // diagonals stored per column of the band array
constexpr int GMW = 8;   // warps that take part in the G / M products
#ifndef TBK_BAND_PREFETCH
#define TBK_BAND_PREFETCH 1
#endif

// Index of element (row i, panel column c) of the V / Z arrays ([rows][8] complex, 128 bytes per row).  The column is
// swizzled with the row so that "one lane per row, same column" (panel factorisation, Z = Y T + V C2) and "8 rows x 4
// column pairs" (tensor-core operands of the trailing update) are both free of bank conflicts.
__device__ __forceinline__ int vz(int i, int c) { return i * BB + (c ^ (i & 7)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Sum NV <= 16 values (a[NV ..] must be zero) over the CTA in a fixed order; every thread gets the results.  Two levels:
// per-warp sums, then thread v adds the warps' values of entry v (warps >= nw hold no rows: exact zeros, skipped).
template <int NV, int WARPS>
__device__ __forceinline__ void cta_sum_n(double (&a)[16], double* red, int tid, int nw) {
    // reduce-scatter over the warp (16 + 8 + 4 + 2 + 1 shuffles instead of 5 per value): even lane L ends with entry L / 2
    const int lane = tid & 31;
#pragma unroll
    for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int v = 0; v < half; ++v) {
            const double mine = up ? a[half + v] : a[v];
            const double other = up ? a[v] : a[half + v];
            a[v] = mine + __shfl_xor_sync(0xffffffffu, other, off);
        }
    }
    a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
    double* tot = red + WARPS * 16;
    if ((lane & 1) == 0 && (lane >> 1) < NV) red[(tid >> 5) * 16 + (lane >> 1)] = a[0];
    __syncthreads();
    if (tid < NV) {
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += red[w * 16 + tid];
        tot[tid] = s;
    }
    __syncthreads();
#pragma unroll
    for (int v = 0; v < NV; ++v) a[v] = tot[v];
}

// Column K of the panel QR (E: [rows][8] complex).  acc holds this thread's share of s_c = sum_{i > K} conj(E[i][K])
// E[i][c], c = K .. 7 (pairs re, im; s_K = |x|^2); on return it holds the share for column K + 1 of the updated panel.
template <int K, int THREADS>
__device__ __forceinline__ void qr_step(double2* V, int m, double (&acc)[16], double* red, double* Rb, double* taus,
                                        int tid) {
    constexpr int WARPS = THREADS / 32;
    constexpr int NC = BB - K;
    if (K >= m) {  // (only in a last panel with fewer than 8 rows) no reflector: tau = 0, R row stays zero
        if (tid == 0) {
            taus[2 * K] = 0.0;
            taus[2 * K + 1] = 0.0;
        }
        return;
    }
    const int nw = (m + 31) >> 5;
    cta_sum_n<2 * NC, WARPS>(acc, red, tid, nw < WARPS ? nw : WARPS);
    const double2 al = V[vz(K, K)];
    double beta, tr, ti, sr, si;
    householder_gen(al.x, al.y, acc[0], beta, tr, ti, sr, si);
    double ur[BB], ui[BB];  // u_c = conj(tau) (E[K][c] + conj(scale) s_c): row K of R is E[K][c] - u_c
#pragma unroll
    for (int c = K + 1; c < BB; ++c) {
        const double2 e = V[vz(K, c)];
        const double s_r = acc[2 * (c - K)], s_i = acc[2 * (c - K) + 1];
        const double zr = e.x + sr * s_r + si * s_i;
        const double zi = e.y + sr * s_i - si * s_r;
        ur[c] = tr * zr + ti * zi;
        ui[c] = tr * zi - ti * zr;
        if (tid == 0) {
            Rb[(K * BB + c) * 2] = e.x - ur[c];
            Rb[(K * BB + c) * 2 + 1] = e.y - ui[c];
        }
    }
    if (tid == 0) {
        Rb[(K * BB + K) * 2] = beta;
        Rb[(K * BB + K) * 2 + 1] = 0.0;
        taus[2 * K] = tr;
        taus[2 * K + 1] = ti;
    }
#pragma unroll
    for (int v = 0; v < 16; ++v) acc[v] = 0.0;
    for (int i = tid; i < m; i += THREADS) {
        if (i <= K) continue;
        double2* e = V + i * BB;
        const int sw = i & 7;  // (swizzled columns, see vz)
        const double2 x = e[K ^ sw];
        const double vr = x.x * sr - x.y * si, vi = x.x * si + x.y * sr;
        e[K ^ sw] = make_double2(vr, vi);
        double nr[BB], ni[BB];
#pragma unroll
        for (int c = K + 1; c < BB; ++c) {
            const double2 ec = e[c ^ sw];
            nr[c] = ec.x - (vr * ur[c] - vi * ui[c]);
            ni[c] = ec.y - (vr * ui[c] + vi * ur[c]);
            e[c ^ sw] = make_double2(nr[c], ni[c]);
        }
        if (K + 1 < BB && i > K + 1) {
            constexpr int K1 = (K + 1 < BB) ? K + 1 : K;
#pragma unroll
            for (int c = K1; c < BB; ++c) {  // conj(x') e_c with x' = E[i][K + 1]
                acc[2 * (c - K1)] = fma(nr[K1], nr[c], fma(ni[K1], ni[c], acc[2 * (c - K1)]));
                acc[2 * (c - K1) + 1] = fma(nr[K1], ni[c], fma(-ni[K1], nr[c], acc[2 * (c - K1) + 1]));
            }
        }
    }
}

// A22 -= V Z^H + Z V^H on the FP64 tensor cores, rows / columns >= r0 (a multiple of 8).  V, Z: [rows][8] complex,
// row i = absolute row r0 + i.  Per 8 x 8 block and plane 8 mma.m8n8k4 with K = the panel column (lane t supplies columns
// 2t and 2t + 1 of its row):
//   Re -= Vr Zr^T + Vi Zi^T + Zr Vr^T + Zi Vi^T,   Im -= Vi Zr^T - Vr Zi^T + Zi Vr^T - Zr Vi^T,
// accumulating in place on fragments loaded from / stored to the packed planes.  A warp owns 8 x 32 strips of the lower
// triangle, round robin (same organisation as the trailing update of eig_tridiag_panel.cu).
template <int WARPS>
__device__ __forceinline__ void her2k_update(double* Ar, double* Ai, int N, int r0, const double2* V, const double2* Z,
                                             int w, int lane) {
    const int g = lane >> 2, tq = lane & 3;
    const int I0 = r0 >> 3, NBk = (N + 7) >> 3;
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
        if (TBK_BAND_PREFETCH) {  // the warp's next strip -> L2: lane = (row of the strip, 8 x 8 block), one line per plane
            int In = I, sn = s + WARPS;
            int nsn = ((In - I0) >> 2) + 1;
            while (In < NBk && sn >= nsn) {
                sn -= nsn;
                ++In;
                nsn = ((In - I0) >> 2) + 1;
            }
            const int rown = 8 * In + g, coln = 8 * (I0 + 4 * sn + tq);
            if (In < NBk && rown < N && coln < rown) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(Ar + tri((long)rown) + coln));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(Ai + trs((long)rown) + coln));
            }
        }
        const int ri = 8 * (I - I0) + g;  // (rows past the end: inside the zero padding of the arrays)
        const double2 av0 = V[vz(ri, 2 * tq)], av1 = V[vz(ri, 2 * tq + 1)];
        const double2 az0 = Z[vz(ri, 2 * tq)], az1 = Z[vz(ri, 2 * tq + 1)];
#pragma unroll
        for (int jb = 0; jb < 4; ++jb) {
            if (J0 + jb <= I) {  // warp-uniform
                const int ci = 8 * (J0 + jb - I0) + g;
                const double2 bv0 = V[vz(ci, 2 * tq)], bv1 = V[vz(ci, 2 * tq + 1)];
                const double2 bz0 = Z[vz(ci, 2 * tq)], bz1 = Z[vz(ci, 2 * tq + 1)];
                dmma884(cre[jb][0], cre[jb][1], -av0.x, bz0.x);
                dmma884(cim[jb][0], cim[jb][1], -av0.y, bz0.x);
                dmma884(cre[jb][0], cre[jb][1], -av0.y, bz0.y);
                dmma884(cim[jb][0], cim[jb][1], av0.x, bz0.y);
                dmma884(cre[jb][0], cre[jb][1], -az0.x, bv0.x);
                dmma884(cim[jb][0], cim[jb][1], -az0.y, bv0.x);
                dmma884(cre[jb][0], cre[jb][1], -az0.y, bv0.y);
                dmma884(cim[jb][0], cim[jb][1], az0.x, bv0.y);
                dmma884(cre[jb][0], cre[jb][1], -av1.x, bz1.x);
                dmma884(cim[jb][0], cim[jb][1], -av1.y, bz1.x);
                dmma884(cre[jb][0], cre[jb][1], -av1.y, bz1.y);
                dmma884(cim[jb][0], cim[jb][1], av1.x, bz1.y);
                dmma884(cre[jb][0], cre[jb][1], -az1.x, bv1.x);
                dmma884(cim[jb][0], cim[jb][1], -az1.y, bv1.x);
                dmma884(cre[jb][0], cre[jb][1], -az1.y, bv1.y);
                dmma884(cim[jb][0], cim[jb][1], az1.x, bv1.y);
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
}

__host__ __device__ inline size_t band_smem_doubles(int n, int warps) {
    const int np = (n + 7) & ~7;
    // V Z [np][8] complex, reduction buffer [warps + 1][16], G / M partials [GMW][256], G M (4 x 64),
    // T MT C2 R (64 complex each), tau
    return (size_t)4 * BB * np + (size_t)(warps + 1) * 16 + (size_t)GMW * 256 + 4 * 64 + 4 * 128 + 16;
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
band_reduce_kernel(double* __restrict__ Hp, int N, long nk, double2* __restrict__ band_all) {
    constexpr int WARPS = THREADS / 32;
    constexpr int WGM = WARPS < GMW ? WARPS : GMW;
    extern __shared__ __align__(16) double smb[];
    const int NP = (N + 7) & ~7;
    double2* V = reinterpret_cast<double2*>(smb);   // [NP][8] panel E, then V: row i = absolute row r0 + i
    double2* Z = V + (size_t)BB * NP;               // [NP][8] Y = A22 V, then Z
    double* red = reinterpret_cast<double*>(Z + (size_t)BB * NP);  // [WARPS + 1][16]
    double* part = red + (WARPS + 1) * 16;          // [GMW][256] fragment partial sums of G and M
    double* GM = part + GMW * 256;                  // [4][64]: Gr, Gi, Mr, Mi
    double* Ts = GM + 4 * 64;                       // [64] complex: T
    double* MT = Ts + 128;                          // [64] complex: M T
    double* C2 = MT + 128;                          // [64] complex: -1/2 T^H M T
    double* Rb = C2 + 128;                          // [64] complex: R of the panel QR
    double* taus = Rb + 128;                        // [8] complex

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const long kk = blockIdx.x;
    if (kk >= nk) return;
    double* Ar = Hp + kk * (long)N * N;
    double* Ai = Ar + tri(N);
    double2* band = band_all + kk * (long)N * BWD;

#ifdef TBK_BAND_TIMING  // debug build: cycles per phase of CTA 0 (thread 0)
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
#define BTICK(i)                          \
    do {                                  \
        const long long now_ = clock64(); \
        tacc[i] += now_ - tlast;          \
        tlast = now_;                     \
    } while (0)
#else
#define BTICK(i)
#endif
    int c0 = 0;
    for (; N - c0 - BB >= 2; c0 += BB) {
        const int r0 = c0 + BB;
        const int m = N - r0;
        const int mp = (m + 7) & ~7;
        // --- (a) the block below the band -> shared memory (rows m .. mp - 1: zero) ---
        for (int idx = tid; idx < mp * BB; idx += THREADS) {
            const int i = idx >> 3, c = idx & 7;
            const long r = r0 + i;
            double2 e = make_double2(0.0, 0.0);
            if (i < m) {
                e.x = Ar[tri(r) + c0 + c];
                e.y = Ai[trs(r) + c0 + c];
            }
            V[vz(i, c)] = e;
        }
        if (tid < 128) Rb[tid] = 0.0;
        __syncthreads();
        BTICK(0);  // panel load
        // --- (b) Householder QR of the panel, in place: v_k below the diagonal, R to Rb ---
        {
            double acc[16];
#pragma unroll
            for (int v = 0; v < 16; ++v) acc[v] = 0.0;
            for (int i = tid; i < m; i += THREADS) {
                if (i == 0) continue;
                const double2 x = V[vz(i, 0)];
#pragma unroll
                for (int c = 0; c < BB; ++c) {
                    const double2 e = V[vz(i, c)];
                    acc[2 * c] = fma(x.x, e.x, fma(x.y, e.y, acc[2 * c]));
                    acc[2 * c + 1] = fma(x.x, e.y, fma(-x.y, e.x, acc[2 * c + 1]));
                }
            }
            qr_step<0, THREADS>(V, m, acc, red, Rb, taus, tid);
            qr_step<1, THREADS>(V, m, acc, red, Rb, taus, tid);
            qr_step<2, THREADS>(V, m, acc, red, Rb, taus, tid);
            qr_step<3, THREADS>(V, m, acc, red, Rb, taus, tid);
            qr_step<4, THREADS>(V, m, acc, red, Rb, taus, tid);
            qr_step<5, THREADS>(V, m, acc, red, Rb, taus, tid);
            qr_step<6, THREADS>(V, m, acc, red, Rb, taus, tid);
            qr_step<7, THREADS>(V, m, acc, red, Rb, taus, tid);
        }
        __syncthreads();
        BTICK(1);  // QR
        // unit diagonal / zeros above it; columns without a reflector (m < 8) are zero vectors
        if (tid < 64) {
            const int k = tid >> 3, c = tid & 7;
            if (k < m && c >= k) V[vz(k, c)] = make_double2((c == k) ? 1.0 : 0.0, 0.0);
        }
        if (m < BB) {
            for (int idx = tid; idx < mp * BB; idx += THREADS)
                if ((idx & 7) >= m) V[vz(idx >> 3, idx & 7)] = make_double2(0.0, 0.0);
        }
        // band columns c0 .. c0 + 7: diagonal block (d <= 7 - k), R (8 - k <= d <= 8), room for the bulges (zero)
        if (tid < BB * BWD) {
            const int k = tid >> 4, d = tid & 15;
            double2 val = make_double2(0.0, 0.0);
            if (d <= 7 - k) {
                const long r = c0 + k + d;
                val.x = Ar[tri(r) + c0 + k];
                if (d > 0) val.y = Ai[trs(r) + c0 + k];
            } else if (d <= 8) {
                const int i = d - 8 + k;
                val.x = Rb[(i * BB + k) * 2];
                val.y = Rb[(i * BB + k) * 2 + 1];
            }
            band[(long)(c0 + k) * BWD + d] = val;
        }
        __syncthreads();
        BTICK(2);  // V fix-up, band write
        // --- (c) Y = A22 V on the tensor cores.  A warp owns block rows I = w, w + WARPS, ... of the FULL Hermitian
        //     matrix, two at a time (they share the V operand): element (r, c) is read as stored for c <= r and as the
        //     conjugate of (c, r) right of the diagonal.  The 8 matrix values a lane needs for the NEXT pair of 8 x 8
        //     blocks are loaded into registers before the current pair is multiplied. ---
        const int nbk = mp >> 3;
        {
            const int npair = (nbk + 1) >> 1;  // pairs of ADJACENT block rows (2p, 2p + 1): their transposed reads share lines
            const int total = ((w < npair) ? (npair - w + WARPS - 1) / WARPS : 0) * nbk;
            const auto fetch = [&](int Ia, int J, double (&o)[8]) {
                int cidx[2], tcr[2], tci[2];
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    cidx[ks] = r0 + 8 * J + 4 * ks + t;
                    tcr[ks] = (cidx[ks] * (cidx[ks] + 1)) >> 1;
                    tci[ks] = (cidx[ks] * (cidx[ks] - 1)) >> 1;
                }
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int I = Ia + rr;
                    const int r = r0 + 8 * I + g;
                    const bool rok = I < nbk && r < N;
                    const int trr = (r * (r + 1)) >> 1, tsr = (r * (r - 1)) >> 1;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const int c = cidx[ks];
                        const bool ok = rok && c < N;
                        const bool low = c <= r;
                        const int offr = low ? trr + c : tcr[ks] + r;
                        const int offi = low ? tsr + c : tci[ks] + r;
                        o[rr * 4 + 2 * ks] = ok ? Ar[offr] : 0.0;
                        const double im = (ok && r != c) ? Ai[offi] : 0.0;
                        o[rr * 4 + 2 * ks + 1] = low ? im : -im;
                    }
                }
            };
            // L2 prefetch PD pairs of blocks ahead: lane = (block row of the pair, plane, row of the 8 x 8 block), one line each
            constexpr int PD = 4;
            const auto prefetch = [&](int Ia, int J) {
                const int I = Ia + (lane >> 4), rw = lane & 7;
                const bool pl = (lane & 8) != 0;
                const bool low = J <= I;
                const int rr_ = r0 + 8 * (low ? I : J) + rw;   // stored row
                const int cc_ = r0 + 8 * (low ? J : I);        // first stored column of the piece
                if (I < nbk && rr_ < N && cc_ < rr_) {
                    const double* ptr = pl ? Ai + (((rr_ * (rr_ - 1)) >> 1) + cc_) : Ar + (((rr_ * (rr_ + 1)) >> 1) + cc_);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
                }
            };
            int Ip = 2 * w, Jp = 0, qp = 0;  // prefetch cursor
            for (; TBK_BAND_PREFETCH && qp < PD && qp < total; ++qp) {
                prefetch(Ip, Jp);
                if (++Jp == nbk) {
                    Jp = 0;
                    Ip += 2 * WARPS;
                }
            }
            int Ii = 2 * w, Ji = 0;  // fetch cursor (one pair of blocks ahead)
            double cur[8], nxt[8];
            if (total > 0) fetch(Ii, Ji, cur);
            if (++Ji == nbk) {
                Ji = 0;
                Ii += 2 * WARPS;
            }
            int Ic = 2 * w, Jc = 0;
            double ya[4] = {0.0, 0.0, 0.0, 0.0}, yb[4] = {0.0, 0.0, 0.0, 0.0};  // (yr0, yr1, yi0, yi1) of the two block rows
            for (int q = 0; q < total; ++q) {
                if (TBK_BAND_PREFETCH && qp < total) {
                    prefetch(Ip, Jp);
                    ++qp;
                    if (++Jp == nbk) {
                        Jp = 0;
                        Ip += 2 * WARPS;
                    }
                }
                if (q + 1 < total) fetch(Ii, Ji, nxt);
                if (++Ji == nbk) {
                    Ji = 0;
                    Ii += 2 * WARPS;
                }
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const int cc = 8 * Jc + 4 * ks + t;
                    const double2 bv = V[vz(cc, g)];
                    const double aar = cur[2 * ks], aai = cur[2 * ks + 1];
                    const double abr = cur[4 + 2 * ks], abi = cur[4 + 2 * ks + 1];
                    dmma884(ya[0], ya[1], aar, bv.x);
                    dmma884(yb[0], yb[1], abr, bv.x);
                    dmma884(ya[2], ya[3], aar, bv.y);
                    dmma884(yb[2], yb[3], abr, bv.y);
                    dmma884(ya[0], ya[1], -aai, bv.y);
                    dmma884(yb[0], yb[1], -abi, bv.y);
                    dmma884(ya[2], ya[3], aai, bv.x);
                    dmma884(yb[2], yb[3], abi, bv.x);
                }
                if (++Jc == nbk) {
                    const int ria = 8 * Ic + g, rib = ria + 8;
                    // (rows past the end of the matrix: zeros into the padding rows of Z)
                    Z[vz(ria, 2 * t)] = make_double2(ya[0], ya[2]);
                    Z[vz(ria, 2 * t + 1)] = make_double2(ya[1], ya[3]);
                    if (Ic + 1 < nbk) {
                        Z[vz(rib, 2 * t)] = make_double2(yb[0], yb[2]);
                        Z[vz(rib, 2 * t + 1)] = make_double2(yb[1], yb[3]);
                    }
#pragma unroll
                    for (int v = 0; v < 4; ++v) ya[v] = yb[v] = 0.0;
                    Jc = 0;
                    Ic += 2 * WARPS;
                }
#pragma unroll
                for (int v = 0; v < 8; ++v) cur[v] = nxt[v];
            }
        }
        __syncthreads();
        BTICK(3);  // Y = A22 V
        // --- (d) G = V^H V, M = V^H Y (8 x 8) on the tensor cores: warps split the rows, fragments summed in a fixed order ---
        if (w < WGM) {
            double gr0 = 0.0, gr1 = 0.0, gi0 = 0.0, gi1 = 0.0, mr0 = 0.0, mr1 = 0.0, mi0 = 0.0, mi1 = 0.0;
            for (int ks = w; ks < (mp >> 2); ks += WGM) {
                const int i = 4 * ks + t;
                const double2 a = V[vz(i, g)];
                double2 y = make_double2(0.0, 0.0);
                if (i < m) y = Z[vz(i, g)];
                dmma884(gr0, gr1, a.x, a.x);
                dmma884(gi0, gi1, a.x, a.y);
                dmma884(mr0, mr1, a.x, y.x);
                dmma884(mi0, mi1, a.x, y.y);
                dmma884(gr0, gr1, a.y, a.y);
                dmma884(gi0, gi1, -a.y, a.x);
                dmma884(mr0, mr1, a.y, y.y);
                dmma884(mi0, mi1, -a.y, y.x);
            }
            double* pw = part + w * 256;
            pw[0 * 32 + lane] = gr0;
            pw[1 * 32 + lane] = gr1;
            pw[2 * 32 + lane] = gi0;
            pw[3 * 32 + lane] = gi1;
            pw[4 * 32 + lane] = mr0;
            pw[5 * 32 + lane] = mr1;
            pw[6 * 32 + lane] = mi0;
            pw[7 * 32 + lane] = mi1;
        }
        __syncthreads();
        for (int q = tid; q < 256; q += THREADS) {
            double s = 0.0;
#pragma unroll
            for (int ww = 0; ww < WGM; ++ww) s += part[ww * 256 + q];
            const int tile = q >> 6, h = (q >> 5) & 1, ln = q & 31;
            GM[tile * 64 + (ln >> 2) * BB + 2 * (ln & 3) + h] = s;
        }
        __syncthreads();
        // --- (e) T (row i by lane i of warp 0: T[0:k, k] = -tau_k T[0:k, 0:k] G[0:k, k]), M T, C2 = -1/2 T^H M T ---
        if (tid < BB) {
            double trow_r[BB], trow_i[BB];
#pragma unroll
            for (int k = 0; k < BB; ++k) {
                const double tkr = taus[2 * k], tki = taus[2 * k + 1];
                double sr = 0.0, si = 0.0;
#pragma unroll
                for (int l = 0; l < k; ++l) {  // (trow[l] = 0 for l < i)
                    const double gr = GM[l * BB + k], gi = GM[64 + l * BB + k];
                    sr += trow_r[l] * gr - trow_i[l] * gi;
                    si += trow_r[l] * gi + trow_i[l] * gr;
                }
                double vr = 0.0, vi = 0.0;
                if (tid < k) {
                    vr = -(tkr * sr - tki * si);
                    vi = -(tkr * si + tki * sr);
                } else if (tid == k) {
                    vr = tkr;
                    vi = tki;
                }
                trow_r[k] = vr;
                trow_i[k] = vi;
                Ts[(tid * BB + k) * 2] = vr;
                Ts[(tid * BB + k) * 2 + 1] = vi;
            }
        }
        __syncthreads();
        if (tid < 64) {
            const int i = tid >> 3, j = tid & 7;
            double sr = 0.0, si = 0.0;
#pragma unroll
            for (int l = 0; l < BB; ++l) {
                const double ar = GM[128 + i * BB + l], ai = GM[192 + i * BB + l];
                const double br = Ts[(l * BB + j) * 2], bi = Ts[(l * BB + j) * 2 + 1];
                sr += ar * br - ai * bi;
                si += ar * bi + ai * br;
            }
            MT[tid * 2] = sr;
            MT[tid * 2 + 1] = si;
        }
        __syncthreads();
        if (tid < 64) {
            const int i = tid >> 3, j = tid & 7;
            double sr = 0.0, si = 0.0;
#pragma unroll
            for (int l = 0; l < BB; ++l) {  // conj(T[l][i]) MT[l][j]
                const double ar = Ts[(l * BB + i) * 2], ai = Ts[(l * BB + i) * 2 + 1];
                const double br = MT[(l * BB + j) * 2], bi = MT[(l * BB + j) * 2 + 1];
                sr += ar * br + ai * bi;
                si += ar * bi - ai * br;
            }
            C2[tid * 2] = -0.5 * sr;
            C2[tid * 2 + 1] = -0.5 * si;
        }
        __syncthreads();
        BTICK(4);  // G, M, T, C2
        // --- (f) Z = Y T + V C2, row by row (in place over Y) ---
        for (int i = tid; i < m; i += THREADS) {
            double2 y[BB], v[BB];
#pragma unroll
            for (int l = 0; l < BB; ++l) {
                y[l] = Z[vz(i, l)];
                v[l] = V[vz(i, l)];
            }
#pragma unroll
            for (int j = 0; j < BB; ++j) {
                double zr = 0.0, zi = 0.0;
#pragma unroll
                for (int l = 0; l < BB; ++l) {
                    if (l <= j) {  // T is upper triangular
                        const double br = Ts[(l * BB + j) * 2], bi = Ts[(l * BB + j) * 2 + 1];
                        zr += y[l].x * br - y[l].y * bi;
                        zi += y[l].x * bi + y[l].y * br;
                    }
                    const double cr = C2[(l * BB + j) * 2], ci = C2[(l * BB + j) * 2 + 1];
                    zr += v[l].x * cr - v[l].y * ci;
                    zi += v[l].x * ci + v[l].y * cr;
                }
                Z[vz(i, j)] = make_double2(zr, zi);
            }
        }
        __syncthreads();
        BTICK(5);  // Z
        // --- (g) A22 -= V Z^H + Z V^H ---
        her2k_update<WARPS>(Ar, Ai, N, r0, V, Z, w, lane);
        __syncthreads();
        BTICK(6);  // trailing update
    }
#ifdef TBK_BAND_TIMING
    if (kk == 0 && tid == 0)
        printf("band timing N=%d Mcyc: load %.3f qr %.3f fixup %.3f AV %.3f GMT %.3f Z %.3f her2k %.3f\n", N, tacc[0] * 1e-6,
               tacc[1] * 1e-6, tacc[2] * 1e-6, tacc[3] * 1e-6, tacc[4] * 1e-6, tacc[5] * 1e-6, tacc[6] * 1e-6);
#endif
    // the remaining block (at most 9 rows) lies inside the band already
    for (int idx = tid; idx < (N - c0) * BWD; idx += THREADS) {
        const int c = c0 + (idx >> 4), d = idx & 15;
        double2 val = make_double2(0.0, 0.0);
        if (d <= BB && c + d < N) {
            const long r = c + d;
            val.x = Ar[tri(r) + c];
            if (d > 0) val.y = Ai[trs(r) + c];
        }
        band[(long)c * BWD + d] = val;
    }
}

// Stage 2: band[c][d] = A[c + d, c] (d < 16) -> d, e.  Eight lanes per matrix, four matrices per warp; lane gl holds row
// R0 + gl of the diagonal block (the full Hermitian row: left of the diagonal from the band columns R0 .. R0 + gl, right
// of it the conjugates of its own column) and row R0 + 8 + gl of the block below, columns R0 .. R0 + 7.  Products along a
// row are local; the only sums across lanes are (tau p)^H v, the reflector norm and v2^H B.
template <int WPB>
__global__ void __launch_bounds__(32 * WPB, 16 / WPB)
band_chase_kernel(double2* band_all, int N, long nk, double* __restrict__ D, double* __restrict__ E) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, gl = lane & 7;
    const long mat = ((long)blockIdx.x * WPB + (threadIdx.x >> 5)) * 4 + (lane >> 3);
    const bool act = mat < nk;
    double2* band = band_all + (act ? mat : 0) * (long)N * BWD;
    for (int j = 0; j < N - 1; ++j) {
        // reflector that annihilates column j below the sub-diagonal (rows j + 1 .. j + 8)
        double xr = 0.0, xi = 0.0;
        if (act && j + 1 + gl < N) {
            const double2 x = band[(long)j * BWD + 1 + gl];
            xr = x.x;
            xi = x.y;
        }
        double xn = (gl >= 1) ? xr * xr + xi * xi : 0.0;
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) xn += __shfl_xor_sync(FULL, xn, off);
        const double alr = __shfl_sync(FULL, xr, 0, 8), ali = __shfl_sync(FULL, xi, 0, 8);
        double beta, tr, ti, sr, si;
        householder_gen(alr, ali, xn, beta, tr, ti, sr, si);
        if (act && gl == 0) {
            D[mat * N + j] = band[(long)j * BWD].x;
            E[mat * N + j] = beta;
        }
        double ovr = (gl == 0) ? 1.0 : xr * sr - xi * si;  // own element of v, then the whole vector
        double ovi = (gl == 0) ? 0.0 : xr * si + xi * sr;
        double vr[BB], vi[BB];
#pragma unroll
        for (int c = 0; c < BB; ++c) {
            vr[c] = __shfl_sync(FULL, ovr, c, 8);
            vi[c] = __shfl_sync(FULL, ovi, c, 8);
        }
        for (int R0 = j + 1; R0 < N; R0 += BB) {
            const bool topok = act && (R0 + gl < N);
            const bool botok = act && (R0 + BB + gl < N);
            const double2* pcol = band + (long)R0 * BWD + gl;       // + 15 c: element (row R0 + gl, column R0 + c), c <= gl
            const double2* pown = band + (long)(R0 + gl) * BWD;     // + d: element (row R0 + gl + d, column R0 + gl)
            double dr[BB], di[BB], br[BB], bi[BB];
#pragma unroll
            for (int c = 0; c < BB; ++c) {
                double2 e = make_double2(0.0, 0.0);
                if (topok && c <= gl) e = pcol[15 * c];
                if (topok && c > gl && R0 + c < N) {  // right of the diagonal: conjugate of the own column
                    e = pown[c - gl];
                    e.y = -e.y;
                }
                dr[c] = e.x;
                di[c] = (c == gl) ? 0.0 : e.y;
                double2 f = make_double2(0.0, 0.0);
                if (botok) f = pcol[15 * c + BB];
                br[c] = f.x;
                bi[c] = f.y;
            }
            // p = D v (local), tau p, dot = (tau p)^H v, w = tau p - 1/2 tau dot v
            double p_r = 0.0, p_i = 0.0, y_r = 0.0, y_i = 0.0;
#pragma unroll
            for (int c = 0; c < BB; ++c) {
                p_r = fma(dr[c], vr[c], fma(-di[c], vi[c], p_r));
                p_i = fma(dr[c], vi[c], fma(di[c], vr[c], p_i));
                y_r = fma(br[c], vr[c], fma(-bi[c], vi[c], y_r));
                y_i = fma(br[c], vi[c], fma(bi[c], vr[c], y_i));
            }
            const double tpr = tr * p_r - ti * p_i, tpi = tr * p_i + ti * p_r;
            double d_r = tpr * ovr + tpi * ovi, d_i = tpr * ovi - tpi * ovr;
#pragma unroll
            for (int off = 4; off > 0; off >>= 1) {
                d_r += __shfl_xor_sync(FULL, d_r, off);
                d_i += __shfl_xor_sync(FULL, d_i, off);
            }
            const double al_r = -0.5 * (tr * d_r - ti * d_i), al_i = -0.5 * (tr * d_i + ti * d_r);
            const double owr = tpr + al_r * ovr - al_i * ovi, owi = tpi + al_r * ovi + al_i * ovr;
            const double tyr = tr * y_r - ti * y_i, tyi = tr * y_i + ti * y_r;  // tau (B v)
#pragma unroll
            for (int c = 0; c < BB; ++c) {
                const double wr = __shfl_sync(FULL, owr, c, 8), wi = __shfl_sync(FULL, owi, c, 8);
                // D -= v w^H + w v^H (only the stored part, c <= gl, is needed);  B -= tau (B v) v^H
                dr[c] -= ovr * wr + ovi * wi + owr * vr[c] + owi * vi[c];
                di[c] -= ovi * wr - ovr * wi + owi * vr[c] - owr * vi[c];
                br[c] -= tyr * vr[c] + tyi * vi[c];
                bi[c] -= tyi * vr[c] - tyr * vi[c];
            }
            // next reflector: annihilates column 0 of the block below under its first row
            double xn2 = (gl >= 1) ? br[0] * br[0] + bi[0] * bi[0] : 0.0;
#pragma unroll
            for (int off = 4; off > 0; off >>= 1) xn2 += __shfl_xor_sync(FULL, xn2, off);
            const double a2r = __shfl_sync(FULL, br[0], 0, 8), a2i = __shfl_sync(FULL, bi[0], 0, 8);
            double beta2, t2r, t2i, s2r, s2i;
            householder_gen(a2r, a2i, xn2, beta2, t2r, t2i, s2r, s2i);
            const double mv2r = (gl == 0) ? 1.0 : br[0] * s2r - bi[0] * s2i;
            const double mv2i = (gl == 0) ? 0.0 : br[0] * s2i + bi[0] * s2r;
            const double cvr = t2r * mv2r + t2i * mv2i, cvi = t2r * mv2i - t2i * mv2r;  // conj(tau2) v2_own
#pragma unroll
            for (int c = 0; c < BB; ++c) {  // z_c = v2^H B[:, c];  B -= conj(tau2) v2 z
                double z_r = mv2r * br[c] + mv2i * bi[c];
                double z_i = mv2r * bi[c] - mv2i * br[c];
#pragma unroll
                for (int off = 4; off > 0; off >>= 1) {
                    z_r += __shfl_xor_sync(FULL, z_r, off);
                    z_i += __shfl_xor_sync(FULL, z_i, off);
                }
                br[c] -= cvr * z_r - cvi * z_i;
                bi[c] -= cvr * z_i + cvi * z_r;
            }
            br[0] = (gl == 0) ? beta2 : 0.0;
            bi[0] = 0.0;
            double2* qcol = band + (long)R0 * BWD + gl;
#pragma unroll
            for (int c = 0; c < BB; ++c) {
                if (topok && c <= gl) qcol[15 * c] = make_double2(dr[c], (c == gl) ? 0.0 : di[c]);
                if (botok) qcol[15 * c + BB] = make_double2(br[c], bi[c]);
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < BB; ++c) {
                vr[c] = __shfl_sync(FULL, mv2r, c, 8);
                vi[c] = __shfl_sync(FULL, mv2i, c, 8);
            }
            ovr = mv2r;
            ovi = mv2i;
            tr = t2r;
            ti = t2i;
        }
    }
    if (act && gl == 0) {
        D[mat * N + N - 1] = band[(long)(N - 1) * BWD].x;
        E[mat * N + N - 1] = 0.0;
    }
}

// Stage 2, pipelined form: ONE matrix per warp, its four 8-lane groups run four consecutive sweeps at once.  Sweep s may
// perform its step k once sweep s - 1 has finished step k + 1 (their column ranges are disjoint from then on), so in
// warp lockstep group g trails group g - 1 by two steps; a group starts its next sweep (s + 4) when it is free and that
// sweep's predecessor is two steps ahead.  The start times follow a data-independent recurrence every lane evaluates for
// itself.  Per-sweep arithmetic is exactly that of band_chase_kernel (bit-identical results); what changes is that the
// four sweeps touch the same part of the band within a few steps -- the band is streamed from HBM once per FOUR sweeps
// -- and that a small batch exposes four times as many warps.
template <int WPB, int MINB>
__global__ void __launch_bounds__(32 * WPB, MINB)
band_chase_pipe_kernel(double2* band_all, int N, long nk, double* __restrict__ D, double* __restrict__ E) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, gl = lane & 7, grp = lane >> 3;
    const long mat = (long)blockIdx.x * WPB + (threadIdx.x >> 5);
    if (mat >= nk) return;  // (warp-uniform)
    double2* band = band_all + mat * (long)N * BWD;
    const int nsweeps = N - 1;
    // start times of the four most recent sweeps of the schedule, indexed by sweep & 3
    int st0 = 0, st1 = 2, st2 = 4, st3 = 6, ns = 4;
    int s = grp, k = 0, tstart = 2 * grp;
    bool alive = s < nsweeps;
    double vr[BB], vi[BB], ovr = 0.0, ovi = 0.0, tr = 0.0, ti = 0.0;
#pragma unroll
    for (int c = 0; c < BB; ++c) vr[c] = vi[c] = 0.0;
    for (int t = 0; __any_sync(FULL, alive); ++t) {
        const bool run = alive && t >= tstart;
        if (__any_sync(FULL, run && k == 0)) {
            // groups that start a sweep: the reflector that annihilates column s below the sub-diagonal
            const bool first = run && k == 0;
            double xr = 0.0, xi = 0.0;
            if (first && s + 1 + gl < N) {
                const double2 x = band[(long)s * BWD + 1 + gl];
                xr = x.x;
                xi = x.y;
            }
            double xn = (gl >= 1) ? xr * xr + xi * xi : 0.0;
#pragma unroll
            for (int off = 4; off > 0; off >>= 1) xn += __shfl_xor_sync(FULL, xn, off);
            const double alr = __shfl_sync(FULL, xr, 0, 8), ali = __shfl_sync(FULL, xi, 0, 8);
            double beta, ntr, nti, sr, si;
            householder_gen(alr, ali, xn, beta, ntr, nti, sr, si);
            if (first && gl == 0) {
                D[mat * N + s] = band[(long)s * BWD].x;
                E[mat * N + s] = beta;
            }
            const double novr = (gl == 0) ? 1.0 : xr * sr - xi * si;
            const double novi = (gl == 0) ? 0.0 : xr * si + xi * sr;
#pragma unroll
            for (int c = 0; c < BB; ++c) {
                const double a = __shfl_sync(FULL, novr, c, 8), b = __shfl_sync(FULL, novi, c, 8);
                if (first) {
                    vr[c] = a;
                    vi[c] = b;
                }
            }
            if (first) {
                ovr = novr;
                ovi = novi;
                tr = ntr;
                ti = nti;
            }
        }
        const int R0 = s + 1 + BB * k;
        const bool topok = run && (R0 + gl < N);
        const bool botok = run && (R0 + BB + gl < N);
        const double2* pcol = band + (long)R0 * BWD + gl;    // + 15 c: element (row R0 + gl, column R0 + c), c <= gl
        const double2* pown = band + (long)(R0 + gl) * BWD;  // + d: element (row R0 + gl + d, column R0 + gl)
        double dr[BB], di[BB], br[BB], bi[BB];
#pragma unroll
        for (int c = 0; c < BB; ++c) {
            double2 e = make_double2(0.0, 0.0);
            if (topok && c <= gl) e = pcol[15 * c];
            if (topok && c > gl && R0 + c < N) {  // right of the diagonal: conjugate of the own column
                e = pown[c - gl];
                e.y = -e.y;
            }
            dr[c] = e.x;
            di[c] = (c == gl) ? 0.0 : e.y;
            double2 f = make_double2(0.0, 0.0);
            if (botok) f = pcol[15 * c + BB];
            br[c] = f.x;
            bi[c] = f.y;
        }
        double p_r = 0.0, p_i = 0.0, y_r = 0.0, y_i = 0.0;
#pragma unroll
        for (int c = 0; c < BB; ++c) {
            p_r = fma(dr[c], vr[c], fma(-di[c], vi[c], p_r));
            p_i = fma(dr[c], vi[c], fma(di[c], vr[c], p_i));
            y_r = fma(br[c], vr[c], fma(-bi[c], vi[c], y_r));
            y_i = fma(br[c], vi[c], fma(bi[c], vr[c], y_i));
        }
        const double tpr = tr * p_r - ti * p_i, tpi = tr * p_i + ti * p_r;
        double d_r = tpr * ovr + tpi * ovi, d_i = tpr * ovi - tpi * ovr;
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) {
            d_r += __shfl_xor_sync(FULL, d_r, off);
            d_i += __shfl_xor_sync(FULL, d_i, off);
        }
        const double al_r = -0.5 * (tr * d_r - ti * d_i), al_i = -0.5 * (tr * d_i + ti * d_r);
        const double owr = tpr + al_r * ovr - al_i * ovi, owi = tpi + al_r * ovi + al_i * ovr;
        const double tyr = tr * y_r - ti * y_i, tyi = tr * y_i + ti * y_r;
#pragma unroll
        for (int c = 0; c < BB; ++c) {
            const double wr = __shfl_sync(FULL, owr, c, 8), wi = __shfl_sync(FULL, owi, c, 8);
            dr[c] -= ovr * wr + ovi * wi + owr * vr[c] + owi * vi[c];
            di[c] -= ovi * wr - ovr * wi + owi * vr[c] - owr * vi[c];
            br[c] -= tyr * vr[c] + tyi * vi[c];
            bi[c] -= tyi * vr[c] - tyr * vi[c];
        }
        double xn2 = (gl >= 1) ? br[0] * br[0] + bi[0] * bi[0] : 0.0;
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) xn2 += __shfl_xor_sync(FULL, xn2, off);
        const double a2r = __shfl_sync(FULL, br[0], 0, 8), a2i = __shfl_sync(FULL, bi[0], 0, 8);
        double beta2, t2r, t2i, s2r, s2i;
        householder_gen(a2r, a2i, xn2, beta2, t2r, t2i, s2r, s2i);
        const double mv2r = (gl == 0) ? 1.0 : br[0] * s2r - bi[0] * s2i;
        const double mv2i = (gl == 0) ? 0.0 : br[0] * s2i + bi[0] * s2r;
        const double cvr = t2r * mv2r + t2i * mv2i, cvi = t2r * mv2i - t2i * mv2r;
#pragma unroll
        for (int c = 0; c < BB; ++c) {
            double z_r = mv2r * br[c] + mv2i * bi[c];
            double z_i = mv2r * bi[c] - mv2i * br[c];
#pragma unroll
            for (int off = 4; off > 0; off >>= 1) {
                z_r += __shfl_xor_sync(FULL, z_r, off);
                z_i += __shfl_xor_sync(FULL, z_i, off);
            }
            br[c] -= cvr * z_r - cvi * z_i;
            bi[c] -= cvr * z_i + cvi * z_r;
        }
        br[0] = (gl == 0) ? beta2 : 0.0;
        bi[0] = 0.0;
        double2* qcol = band + (long)R0 * BWD + gl;
#pragma unroll
        for (int c = 0; c < BB; ++c) {
            if (topok && c <= gl) qcol[15 * c] = make_double2(dr[c], (c == gl) ? 0.0 : di[c]);
            if (botok) qcol[15 * c + BB] = make_double2(br[c], bi[c]);
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < BB; ++c) {
            const double a = __shfl_sync(FULL, mv2r, c, 8), b = __shfl_sync(FULL, mv2i, c, 8);
            if (run) {
                vr[c] = a;
                vi[c] = b;
            }
        }
        if (run) {
            ovr = mv2r;
            ovi = mv2i;
            tr = t2r;
            ti = t2i;
            const int L = (N + 6 - s) >> 3;  // steps of sweep s
            if (++k >= L) {
                s += 4;
                k = 0;
                alive = s < nsweeps;
                if (alive) {
                    while (ns <= s) {  // start(ns) = max(start(ns - 1) + 2, start(ns - 4) + steps(ns - 4))
                        const int i = ns & 3, ip = (ns - 1) & 3;
                        const int prev = ip == 0 ? st0 : ip == 1 ? st1 : ip == 2 ? st2 : st3;
                        const int own = i == 0 ? st0 : i == 1 ? st1 : i == 2 ? st2 : st3;
                        const int a = prev + 2, b = own + ((N + 6 - (ns - 4)) >> 3);
                        const int v = a > b ? a : b;
                        if (i == 0) st0 = v; else if (i == 1) st1 = v; else if (i == 2) st2 = v; else st3 = v;
                        ++ns;
                    }
                    const int i = s & 3;
                    tstart = i == 0 ? st0 : i == 1 ? st1 : i == 2 ? st2 : st3;
                }
            }
        }
    }
    __syncwarp();
    if (lane == 0) {
        D[mat * N + N - 1] = band[(long)(N - 1) * BWD].x;
        E[mat * N + N - 1] = 0.0;
    }
}

template <int THREADS, int MINB>
cudaError_t launch_band_reduce_t(int n, double* Hp, long nk, double2* band, cudaStream_t st) {
    const size_t smem = band_smem_doubles(n, THREADS / 32) * 8;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t err = cudaFuncSetAttribute(band_reduce_kernel<THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    band_reduce_kernel<THREADS, MINB><<<(unsigned)nk, THREADS, smem, st>>>(Hp, n, nk, band);
    return cudaGetLastError();
}

}  // namespace

bool tridiag_twostage_fits(int n) { return n >= 12 && band_smem_doubles(n, 16) * 8 <= 227 * 1024; }

size_t tridiag_twostage_scratch_bytes(int n, long nk) { return (size_t)nk * n * BWD * sizeof(double2); }

// Stage 1 of nk matrices: Hp -> band (nk * n * 16 complex at band_ws).
cudaError_t launch_band_reduce(int n, double* Hp, long nk, double* band_ws, cudaStream_t st, const Tuning& tune) {
    if (nk <= 0) return cudaSuccess;
    if (nk > 2147483647L || band_ws == nullptr || !tridiag_twostage_fits(n)) return cudaErrorInvalidConfiguration;
    double2* band = reinterpret_cast<double2*>(band_ws);
    // threads per matrix (tuning hook TBK_BAND_T).  Measured on B200, both stages, ms per 1000 matrices (r04k_sweep2.log):
    // N = 200: 256 threads 9.30, 2 x 256 per SM 7.90, 512 threads 9.54;  N = 256: 16.0 / 14.0 / 15.9 -- two independent
    // CTAs per SM overlap one matrix' barrier-bound panel factorisation with the other's tensor-core passes while both
    // fit in shared memory; the final kernels cross over near N = 340 (two 256-thread CTAs / one 512-thread CTA: N = 288:
    // 13.0 / 14.6, 304: 16.6 / 17.1, 320: 19.4 / 19.8, 352: 24.4 / 24.1; r05h_sweep2.log).  Four 128-thread CTAs per SM
    // help only the smallest sizes, where the barriers of the panel factorisation dominate (N = 128: 1.88 against 2.05,
    // 144: 2.53 / 2.57, 160: 3.22 / 3.20, 200: 5.65 / 5.34; r05e_sweep.log).
    const int t = tune.band_t > 0 ? tune.band_t : (n <= 136 ? 128 : n <= 336 ? 257 : 512);
    if (t == 128) return launch_band_reduce_t<128, 4>(n, Hp, nk, band, st);  // (four CTAs per SM)
    if (t == 256) return launch_band_reduce_t<256, 1>(n, Hp, nk, band, st);
    if (t == 257) return launch_band_reduce_t<256, 2>(n, Hp, nk, band, st);  // (two CTAs per SM: 128 registers)
    return launch_band_reduce_t<512, 1>(n, Hp, nk, band, st);
}

long band_chase_wave_matrices(const Tuning& tune) {
    if (tune.band_wave > 0) return tune.band_wave;
    static const long wave = [] {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        return (long)sms * 12;  // 12 warps per SM (168 registers), one matrix per warp
    }();
    return wave;
}

// Stage 2 of nk matrices: band -> D, E [nk][n].
cudaError_t launch_band_chase(int n, double* band_ws, long nk, double* D, double* E, cudaStream_t st, const Tuning& tune) {
    if (nk <= 0) return cudaSuccess;
    if (band_ws == nullptr) return cudaErrorInvalidConfiguration;
    constexpr int WPB = 4;
    if (tune.band_chase == 1) {  // (first form: four matrices per warp, one sweep at a time)
        const long ctas = (nk + 4 * WPB - 1) / (4 * WPB);
        if (ctas > 2147483647L) return cudaErrorInvalidConfiguration;
        band_chase_kernel<WPB><<<(unsigned)ctas, 32 * WPB, 0, st>>>(reinterpret_cast<double2*>(band_ws), n, nk, D, E);
        return cudaGetLastError();
    }
    const long ctas = (nk + WPB - 1) / WPB;
    if (ctas > 2147483647L) return cudaErrorInvalidConfiguration;
    // 12 warps per SM (168 registers, no spills).  Measured against 16 warps at 128 registers with 136 bytes of spills
    // (TBK_BAND_CHASE=4), both stages, ms per 1000 matrices: N = 256: 9.54 / 10.74 (18 944 matrices), 14.6 / 20.1 (296);
    // N = 512: 62.3 / 65.7 (9 472), 81.9 / 104.0 (296) -- gpurun_out/r04y_sweep.log
    if (tune.band_chase == 4)
        band_chase_pipe_kernel<WPB, 4><<<(unsigned)ctas, 32 * WPB, 0, st>>>(reinterpret_cast<double2*>(band_ws), n, nk, D, E);
    else
        band_chase_pipe_kernel<WPB, 3><<<(unsigned)ctas, 32 * WPB, 0, st>>>(reinterpret_cast<double2*>(band_ws), n, nk, D, E);
    return cudaGetLastError();
}

}  // namespace tbk
