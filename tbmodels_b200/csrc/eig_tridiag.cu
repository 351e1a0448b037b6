// eig_tridiag.cu -- batched Hermitian -> real symmetric tridiagonal reduction (eigenvalues only).
//
// First half of the replacement for the per-k scipy.linalg.eigvalsh loop of Model.eigenval
// (reference src/tbmodels/_tb_model.py:1148-1149; LAPACK zheevr with JOBZ='N', UPLO='L'): unblocked Householder
// reflections on the packed lower triangle,
//     x = A[j+1:, j]  ->  (beta, tau, v);   p = tau A22 v;   w = p - (tau/2)(p^H v) v;   A22 -= v w^H + w v^H.
// Four kernels, chosen from N only (never from the batch, so a k-point's bits do not depend on the batch):
//   tridiag_smem_kernel<G, CS>  N <= 164: matrix in shared memory as interleaved complex packed rows, G threads per
//                               matrix (8 .. 512) = row threads x CS column slices
//   tridiag_big_kernel<MAXC>    165 <= N <= 640: 16 warps per matrix, in place in global memory, coalesced row sweeps
//   tridiag_global_kernel       larger: thread-per-row fallback
//   tridiag_mma_kernel          experimental tensor-core rank-2 update (opt-in, see below)
// Every group executes exactly the same instruction sequence (no data-dependent early exits), so sub-warp groups can
// share a warp and use __syncwarp().  Outputs: D[k][0..N) diagonal, E[k][0..N-1) sub-diagonal, consumed by eig_ql.cu.
#include <cstdlib>
#include <cstring>

#include "tbk_kernels.h"
#include "tbk_math.cuh"

namespace tbk {

namespace {

constexpr int TPB = 256;
constexpr size_t kSmemLimit = 220 * 1024;  // shared memory one CTA of the shared-memory kernels may use

template <int G>
__device__ __forceinline__ void group_sync(int group) {
    if (G <= 32) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(G) : "memory");
    }
}

// Sum (a, b) over the G threads of a group; every thread gets the result.  Fixed butterfly order.
template <int G>
__device__ __forceinline__ void group_sum2(double& a, double& b, double* red, int group, int t, int& parity) {
    constexpr int W = G < 32 ? G : 32;
#pragma unroll
    for (int off = W / 2; off > 0; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off, W);
        b += __shfl_xor_sync(0xffffffffu, b, off, W);
    }
    if (G > 32) {
        constexpr int NW = G / 32;
        double* buf = red + parity * 2 * NW;  // double-buffered: one barrier per reduction
        parity ^= 1;
        if ((t & 31) == 0) {
            buf[2 * (t >> 5)] = a;
            buf[2 * (t >> 5) + 1] = b;
        }
        group_sync<G>(group);
        a = 0.0;
        b = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            a += buf[2 * w];
            b += buf[2 * w + 1];
        }
    }
}

// Last-resort variant for matrices beyond the row-sweep kernel's limit (N > 640): one CTA of 256 threads per matrix,
// in place on the packed planes in global memory, one thread per row.  Simple and slow; kept so that no size fails.
__global__ void __launch_bounds__(TPB)
tridiag_global_kernel(double* __restrict__ Hp, int N, long nk, double* __restrict__ D, double* __restrict__ E) {
    constexpr int G = TPB;
    extern __shared__ __align__(16) double sm[];
    const int t = threadIdx.x;
    const long NN = (long)N * N;
    const long nre = tri(N);
    const long kk = blockIdx.x;
    if (kk >= nk) return;
    double* A = Hp + kk * NN;
    double* Vr = sm;
    double* Vi = Vr + N;
    double* Pr = Vi + N;
    double* Pi = Pr + N;
    double* ds = Pi + N;
    double* es = ds + N;
    double* red = es + N;
    double* Ar = A;
    double* Ai = A + nre;
    int parity = 0;

    for (int j = 0; j < N - 1; ++j) {
        const int m = N - 1 - j;
        const int r0 = j + 1;
        const double ar = Ar[tri(r0) + j];
        const double ai = Ai[trs(r0) + j];
        double xn = 0.0, dummy = 0.0;
        for (int a = 1 + t; a < m; a += G) {
            const long I = r0 + a;
            const double xr = Ar[tri(I) + j], xi = Ai[trs(I) + j];
            xn += xr * xr + xi * xi;
        }
        group_sum2<G>(xn, dummy, red, 0, t, parity);
        double beta, tr, ti, sr, si;
        householder_gen(ar, ai, xn, beta, tr, ti, sr, si);
        if (t == 0) {
            ds[j] = Ar[tri(j) + j];
            es[j] = beta;
        }
        for (int a = t; a < m; a += G) {
            if (a == 0) {
                Vr[0] = 1.0;
                Vi[0] = 0.0;
            } else {
                const long I = r0 + a;
                const double xr = Ar[tri(I) + j], xi = Ai[trs(I) + j];
                Vr[a] = xr * sr - xi * si;
                Vi[a] = xr * si + xi * sr;
            }
        }
        group_sync<G>(0);
        double dr = 0.0, di = 0.0;
        for (int a = t; a < m; a += G) {
            const long I = r0 + a;
            double sumr = 0.0, sumi = 0.0;
            const long rowI = tri(I), rowIs = trs(I);
            long triJ = tri(r0), trsJ = trs(r0);
            for (int b = 0; b < m; ++b) {
                const long J = r0 + b;
                const bool left = b < a;
                const long ir = left ? rowI + J : triJ + I;
                const long ii = (b == a) ? 0 : (left ? rowIs + J : trsJ + I);
                const double arr = Ar[ir];
                double aii = Ai[ii];
                aii = (b == a) ? 0.0 : (left ? aii : -aii);
                const double vr = Vr[b], vi = Vi[b];
                sumr += arr * vr - aii * vi;
                sumi += arr * vi + aii * vr;
                triJ += J + 1;
                trsJ += J;
            }
            const double pr = tr * sumr - ti * sumi;
            const double pi = tr * sumi + ti * sumr;
            Pr[a] = pr;
            Pi[a] = pi;
            dr += pr * Vr[a] + pi * Vi[a];
            di += pr * Vi[a] - pi * Vr[a];
        }
        group_sum2<G>(dr, di, red, 0, t, parity);
        const double alr = -0.5 * (tr * dr - ti * di);
        const double ali = -0.5 * (tr * di + ti * dr);
        for (int a = t; a < m; a += G) {
            const double vr = Vr[a], vi = Vi[a];
            Pr[a] += alr * vr - ali * vi;
            Pi[a] += alr * vi + ali * vr;
        }
        group_sync<G>(0);
        for (int a = t; a < m; a += G) {
            const long I = r0 + a;
            const double var = Vr[a], vai = Vi[a], war = Pr[a], wai = Pi[a];
            double* rowr = Ar + tri(I) + r0;
            double* rowi = Ai + trs(I) + r0;
            for (int b = 0; b < a; ++b) {
                const double vbr = Vr[b], vbi = Vi[b], wbr = Pr[b], wbi = Pi[b];
                rowr[b] -= var * wbr + vai * wbi + war * vbr + wai * vbi;
                rowi[b] -= vai * wbr - var * wbi + wai * vbr - war * vbi;
            }
            rowr[a] -= 2.0 * (var * war + vai * wai);
        }
        group_sync<G>(0);
    }
    if (t == 0) {
        ds[N - 1] = Ar[tri(N - 1) + (N - 1)];
        es[N - 1] = 0.0;
    }
    group_sync<G>(0);
    for (int i = t; i < N; i += G) {
        D[kk * N + i] = ds[i];
        E[kk * N + i] = es[i];
    }
}

// ---------------------------------------------------------------------------------------------------------
// Shared-memory resident variant (N <= ~164): the packed matrix is re-laid out on load as interleaved
// complex (double2) lower-packed rows, so every element access is one LDS.128 / STS.128, all indices are
// 32-bit, and the address space is known at compile time.  Row starts tri(i) are distinct mod 8 over aligned
// groups of 8 rows, consecutive columns are consecutive 16-byte words: both access patterns of the
// Hermitian matrix-vector product (own row / conjugated column) are (nearly) bank-conflict free.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int itri(int i) { return (i * (i + 1)) >> 1; }

// G threads per matrix = RT row-threads x CS column slices.  CS > 1 splits every row of the matrix-vector product and
// of the rank-2 update over CS threads (partial sums combined through shared memory): more warps per resident
// matrix, shorter serial loops -- the kernel is latency bound, shared memory caps the number of resident matrices.
template <int G, int CS>
__global__ void __launch_bounds__((G > 512 ? G : 512))
tridiag_smem_kernel(double* __restrict__ Hp, int N, long mstride, long nk, double* __restrict__ D, double* __restrict__ E,
                    int ldo, int off, int nsteps) {
    // Staged reduction (launch_tridiag): the matrix of this launch is the trailing N x N block of a larger one that sits
    // at Hp + k * mstride; the launch performs nsteps Householder steps (all N - 1 if nsteps >= N - 1), writes the
    // diagonal / sub-diagonal entries it produces to D / E [k * ldo + off + i] and, if steps remain, stores the
    // remaining (N - nsteps)^2 trailing block back in packed form at the head of the same slot for the next launch.
    constexpr int NW = G > 32 ? G / 32 : 1;
    constexpr int RT = G / CS;
    static_assert(CS == 1 || (RT % 32 == 0), "column slices must be whole warps");
    extern __shared__ __align__(16) double2 sm2[];
    const int MPB = blockDim.x / G;
    const int group = threadIdx.x / G;
    const int t = threadIdx.x % G;
    const int rr = t % RT, cs = t / RT;
    const int ntri = itri(N);
    const long NN = mstride;
    const int nst = nsteps < N - 1 ? nsteps : N - 1;

    const long kidx = (long)blockIdx.x * MPB + group;
    const bool valid = kidx < nk;
    const long kk = valid ? kidx : nk - 1;  // idle groups shadow the last matrix and store nothing

    const int per_group = ntri + 3 * N + (CS > 1 ? CS * N : 0) + 2 * NW + 1;  // in double2 units
    double2* A = sm2 + (size_t)group * per_group;
    double2* V = A + ntri;
    double2* P = V + N;
    double* ds = reinterpret_cast<double*>(P + N);
    double* es = ds + N;
    double2* PP = reinterpret_cast<double2*>(es + N);  // [CS][N] partial row sums (CS > 1)
    double* red = reinterpret_cast<double*>(PP + (CS > 1 ? CS * N : 0));

    {   // load + interleave.  The real plane has the same packed order as the complex rows (flat copy); an imag
        // plane entry f = trs(i) + j lands at f + i.  Flat loops keep many independent loads in flight.
        // asynchronous 8-byte copies (cp.async, SASS LDGSTS): every element of the matrix is in flight at once, no
        // register staging -- the CTAs of a wave all start in this phase, so it is otherwise pure load latency
        const double* src = Hp + kk * NN;
        const double* srci = src + ntri;
        double* Ad = reinterpret_cast<double*>(A);
        const unsigned sA = (unsigned)__cvta_generic_to_shared(Ad);
#pragma unroll 4
        for (int e = t; e < ntri; e += G)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sA + 16u * e), "l"(src + e) : "memory");
        const int nim = ntri - N;
#pragma unroll 4
        for (int f = t; f < nim; f += G) {
            int i = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)f)) * 0.5f);  // row of strict-lower entry f (approx.)
            while ((i * (i - 1)) / 2 > f) --i;
            while ((i * (i + 1)) / 2 <= f) ++i;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sA + 16u * (f + i) + 8u), "l"(srci + f)
                         : "memory");
        }
        for (int i = t; i < N; i += G) Ad[2 * (itri(i) + i) + 1] = 0.0;
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    group_sync<G>(group);

    int parity = 0;
    for (int j = 0; j < nst; ++j) {
        const int m = N - 1 - j;
        const int r0 = j + 1;
        // --- reflector from column j ---
        const double2 alpha = A[itri(r0) + j];
        double xn = 0.0, dummy = 0.0;
        const double2 x0 = (t < m) ? A[itri(r0 + t) + j] : make_double2(0.0, 0.0);  // first row of this thread
        if (t >= 1 && t < m) xn = x0.x * x0.x + x0.y * x0.y;
        for (int a = t + G; a < m; a += G) {
            const double2 x = A[itri(r0 + a) + j];
            xn += x.x * x.x + x.y * x.y;
        }
        group_sum2<G>(xn, dummy, red, group, t, parity);
        double beta, tr, ti, sr, si;
        householder_gen(alpha.x, alpha.y, xn, beta, tr, ti, sr, si);
        if (t == 0) {
            ds[j] = A[itri(j) + j].x;
            es[j] = beta;
            V[0] = make_double2(1.0, 0.0);
        } else if (t < m) {
            V[t] = make_double2(x0.x * sr - x0.y * si, x0.x * si + x0.y * sr);
        }
        for (int a = t + G; a < m; a += G) {
            const double2 x = A[itri(r0 + a) + j];
            V[a] = make_double2(x.x * sr - x.y * si, x.x * si + x.y * sr);
        }
        group_sync<G>(group);
        // --- p = tau * A22 v (row part from the own packed row, column part conjugated), dot = p^H v ---
        // two independent accumulator pairs halve the DFMA dependency chains
        double dr = 0.0, di = 0.0;
        for (int a = rr; a < m; a += RT) {
            const int I = r0 + a;
            const int rowI = itri(I) + r0;
            double sr0 = 0.0, si0 = 0.0, sr1 = 0.0, si1 = 0.0;
            int b = cs;
            int c0 = itri(r0 + cs) + I;  // tri(J) + I for J = r0 + b, advanced incrementally
            for (; b + CS < m; b += 2 * CS) {
                const int b1 = b + CS;
                const int c1 = c0 + CS * (r0 + b) + (CS * (CS + 1)) / 2;  // tri(J + CS) - tri(J) = CS J + CS (CS + 1) / 2
                const bool l0 = b < a, l1 = b1 < a;
                const double2 z0 = A[l0 ? rowI + b : c0];
                const double2 z1 = A[l1 ? rowI + b1 : c1];
                const double2 v0 = V[b], v1 = V[b1];
                const double y0 = l0 ? z0.y : -z0.y;
                const double y1 = l1 ? z1.y : -z1.y;
                sr0 = fma(z0.x, v0.x, fma(-y0, v0.y, sr0));
                si0 = fma(z0.x, v0.y, fma(y0, v0.x, si0));
                sr1 = fma(z1.x, v1.x, fma(-y1, v1.y, sr1));
                si1 = fma(z1.x, v1.y, fma(y1, v1.x, si1));
                c0 = c1 + CS * (r0 + b1) + (CS * (CS + 1)) / 2;
            }
            if (b < m) {
                const bool l0 = b < a;
                const double2 z0 = A[l0 ? rowI + b : c0];
                const double2 v0 = V[b];
                const double y0 = l0 ? z0.y : -z0.y;
                sr0 = fma(z0.x, v0.x, fma(-y0, v0.y, sr0));
                si0 = fma(z0.x, v0.y, fma(y0, v0.x, si0));
            }
            const double sumr = sr0 + sr1, sumi = si0 + si1;
            if (CS == 1) {
                const double pr = tr * sumr - ti * sumi;
                const double pi = tr * sumi + ti * sumr;
                P[a] = make_double2(pr, pi);
                const double2 va = V[a];
                dr += pr * va.x + pi * va.y;
                di += pr * va.y - pi * va.x;
            } else {
                PP[cs * N + a] = make_double2(sumr, sumi);
            }
        }
        if (CS > 1) {
            group_sync<G>(group);
            for (int a = t; a < m; a += G) {
                double sumr = 0.0, sumi = 0.0;
#pragma unroll
                for (int c = 0; c < CS; ++c) {
                    const double2 q = PP[c * N + a];
                    sumr += q.x;
                    sumi += q.y;
                }
                const double pr = tr * sumr - ti * sumi;
                const double pi = tr * sumi + ti * sumr;
                P[a] = make_double2(pr, pi);
                const double2 va = V[a];
                dr += pr * va.x + pi * va.y;
                di += pr * va.y - pi * va.x;
            }
        }
        group_sum2<G>(dr, di, red, group, t, parity);
        const double alr = -0.5 * (tr * dr - ti * di);
        const double ali = -0.5 * (tr * di + ti * dr);
        for (int a = t; a < m; a += G) {
            const double2 v = V[a];
            double2 p = P[a];
            p.x += alr * v.x - ali * v.y;
            p.y += alr * v.y + ali * v.x;
            P[a] = p;
        }
        group_sync<G>(group);
        // --- A22 -= v w^H + w v^H (lower triangle) ---
        for (int a = rr; a < m; a += RT) {
            const int I = r0 + a;
            const double2 va = V[a], wa = P[a];
            double2* row = A + itri(I) + r0;
#pragma unroll 4
            for (int b = cs; b < a; b += CS) {
                const double2 vb = V[b], wb = P[b];
                double2 z = row[b];
                z.x = fma(-va.x, wb.x, fma(-va.y, wb.y, fma(-wa.x, vb.x, fma(-wa.y, vb.y, z.x))));
                z.y = fma(-va.y, wb.x, fma(va.x, wb.y, fma(-wa.y, vb.x, fma(wa.x, vb.y, z.y))));
                row[b] = z;
            }
            if (cs == 0) row[a].x -= 2.0 * (va.x * wa.x + va.y * wa.y);
        }
        group_sync<G>(group);
    }
    if (nst < N - 1) {
        // stage boundary: emit the entries produced so far and hand the trailing block to the next launch
        if (valid) {
            for (int i = t; i < nst; i += G) {
                D[kidx * ldo + off + i] = ds[i];
                E[kidx * ldo + off + i] = es[i];
            }
            const int M = N - nst;
            double* dst = Hp + kk * NN;
            const int mtri = itri(M);
            for (int e = t; e < mtri; e += G) {  // packed element e of the M x M block = (i, c), c <= i
                int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
                while (itri(i) > e) --i;
                while (itri(i + 1) <= e) ++i;
                const int c = e - itri(i);
                const double2 z = A[itri(nst + i) + nst + c];
                dst[e] = z.x;
                if (c < i) dst[mtri + ((i * (i - 1)) >> 1) + c] = z.y;
            }
        }
        return;
    }
    if (t == 0) {
        ds[N - 1] = A[itri(N - 1) + (N - 1)].x;
        es[N - 1] = 0.0;
    }
    group_sync<G>(group);
    if (valid) {
        for (int i = t; i < N; i += G) {
            D[kidx * ldo + off + i] = ds[i];
            E[kidx * ldo + off + i] = es[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Tensor-core variant for 21 <= N <= 64: one warp per matrix, the FULL Hermitian matrix (both triangles,
// interleaved complex, odd row stride) in shared memory.
//   * the Hermitian matrix-vector product reads plain rows (no triangle selects): LDS.128 + 4 DFMA per element;
//   * the rank-2 update  A -= v w^H + w v^H  runs on the FP64 tensor cores.  Per 8x8 block it is a real product
//     with K = 4:   Re(A) -= [vr vi wr wi] . [wr wi vr vi]^T,   Im(A) -= [vi -vr wi -wr] . [wr wi vr vi]^T,
//     i.e. exactly one mma.sync.m8n8k4.f64 (DMMA.8x8x4) per plane per block, accumulating in place on the
//     C fragments loaded from / stored to shared memory (lower blocks only; each off-diagonal block is mirrored
//     with a conjugate store so the full matrix stays Hermitian for the next matrix-vector product).
//   v and w live in one table XY[row] = (vr, vi, wr, wi) that is zero outside the trailing sub-matrix, so
//   blocks straddling its border need no masking.
// Same arithmetic as the scalar kernels up to the summation order inside the 4-term block products.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_acc(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct MmaLayout {
    int N, NB, S;          // size, 8x8 blocks per dimension, row stride in double2 (odd)
    int as_words;          // N * S
    int per_matrix;        // double2 units per matrix
    __host__ __device__ explicit MmaLayout(int n) {
        N = n;
        NB = (n + 7) / 8;
        S = n | 1;
        as_words = n * S;
        per_matrix = as_words + 2 * NB * 8 /*XY*/ + n /*ds, es*/ + 2 /*pad*/;
    }
};

__global__ void __launch_bounds__(TPB)
tridiag_mma_kernel(const double* __restrict__ Hp, int N, long nk, double* __restrict__ D, double* __restrict__ E) {
    extern __shared__ __align__(16) double2 sm2[];
    const MmaLayout L(N);
    const int MPB = blockDim.x >> 5;
    const int group = threadIdx.x >> 5;
    const int t = threadIdx.x & 31;
    const int g = t >> 2, tq = t & 3;
    const int S = L.S, NB = L.NB;
    const int ntri = itri(N);
    const long NN = (long)N * N;

    const long kidx = (long)blockIdx.x * MPB + group;
    const bool valid = kidx < nk;
    const long kk = valid ? kidx : nk - 1;  // idle warps shadow the last matrix and store nothing

    double2* As = sm2 + (size_t)group * L.per_matrix;
    double2* XY2 = As + L.as_words;                          // row i: XY2[2i] = v_i, XY2[2i+1] = w_i
    double* XY = reinterpret_cast<double*>(XY2);             // row i: (vr, vi, wr, wi)
    double* ds = reinterpret_cast<double*>(XY2 + 2 * NB * 8);
    double* es = ds + N;

    {   // load the packed planes and mirror them into the full matrix
        const double* src = Hp + kk * NN;
        const double* srci = src + ntri;
        double* Ad = reinterpret_cast<double*>(As);
#pragma unroll 4
        for (int e = t; e < ntri; e += 32) {
            int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
            while (itri(i) > e) --i;
            while (itri(i + 1) <= e) ++i;
            const int j = e - itri(i);
            const double v = src[e];
            Ad[2 * (i * S + j)] = v;
            Ad[2 * (j * S + i)] = v;
        }
        const int nim = ntri - N;
#pragma unroll 4
        for (int f = t; f < nim; f += 32) {
            int i = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)f)) * 0.5f);
            while ((i * (i - 1)) / 2 > f) --i;
            while ((i * (i + 1)) / 2 <= f) ++i;
            const int j = f - (i * (i - 1)) / 2;
            const double v = srci[f];
            Ad[2 * (i * S + j) + 1] = v;
            Ad[2 * (j * S + i) + 1] = -v;
        }
        for (int i = t; i < N; i += 32) Ad[2 * (i * S + i) + 1] = 0.0;
        for (int i = t; i < 2 * NB * 8; i += 32) XY2[i] = make_double2(0.0, 0.0);
    }
    __syncwarp();

    for (int j = 0; j < N - 1; ++j) {
        const int r0 = j + 1;
        // --- reflector from column j (rows r0 .. N-1); this lane owns rows r0 + t, r0 + t + 32 ---
        const double2 alpha = As[r0 * S + j];
        const int i0 = r0 + t, i1 = i0 + 32;
        const double2 x0 = (i0 < N) ? As[i0 * S + j] : make_double2(0.0, 0.0);
        const double2 x1 = (i1 < N) ? As[i1 * S + j] : make_double2(0.0, 0.0);
        double xn = x1.x * x1.x + x1.y * x1.y;
        if (t >= 1) xn += x0.x * x0.x + x0.y * x0.y;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) xn += __shfl_xor_sync(0xffffffffu, xn, off);
        double beta, tr, ti, sr, si;
        householder_gen(alpha.x, alpha.y, xn, beta, tr, ti, sr, si);
        if (t == 0) {
            ds[j] = As[j * S + j].x;
            es[j] = beta;
            XY2[2 * j] = make_double2(0.0, 0.0);      // row j leaves the trailing block: v_j = w_j = 0
            XY2[2 * j + 1] = make_double2(0.0, 0.0);
            XY2[2 * r0] = make_double2(1.0, 0.0);
        } else if (i0 < N) {
            XY2[2 * i0] = make_double2(x0.x * sr - x0.y * si, x0.x * si + x0.y * sr);
        }
        if (i1 < N) XY2[2 * i1] = make_double2(x1.x * sr - x1.y * si, x1.x * si + x1.y * sr);
        __syncwarp();
        // --- p = tau * A22 v (plain rows of the full matrix), dot = p^H v ---
        double dr = 0.0, di = 0.0;
        double2 pv[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = i0 + 32 * q;
            pv[q] = make_double2(0.0, 0.0);
            if (i < N) {
                const double2* row = As + i * S;
                double sr0 = 0.0, si0 = 0.0, sr1 = 0.0, si1 = 0.0;
                int b = r0;
#pragma unroll 2
                for (; b + 1 < N; b += 2) {
                    const double2 z0 = row[b], z1 = row[b + 1];
                    const double2 v0 = XY2[2 * b], v1 = XY2[2 * b + 2];
                    sr0 = fma(z0.x, v0.x, fma(-z0.y, v0.y, sr0));
                    si0 = fma(z0.x, v0.y, fma(z0.y, v0.x, si0));
                    sr1 = fma(z1.x, v1.x, fma(-z1.y, v1.y, sr1));
                    si1 = fma(z1.x, v1.y, fma(z1.y, v1.x, si1));
                }
                if (b < N) {
                    const double2 z0 = row[b];
                    const double2 v0 = XY2[2 * b];
                    sr0 = fma(z0.x, v0.x, fma(-z0.y, v0.y, sr0));
                    si0 = fma(z0.x, v0.y, fma(z0.y, v0.x, si0));
                }
                const double sumr = sr0 + sr1, sumi = si0 + si1;
                const double pr = tr * sumr - ti * sumi;
                const double pi = tr * sumi + ti * sumr;
                pv[q] = make_double2(pr, pi);
                const double2 va = XY2[2 * i];
                dr += pr * va.x + pi * va.y;
                di += pr * va.y - pi * va.x;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            dr += __shfl_xor_sync(0xffffffffu, dr, off);
            di += __shfl_xor_sync(0xffffffffu, di, off);
        }
        const double alr = -0.5 * (tr * dr - ti * di);
        const double ali = -0.5 * (tr * di + ti * dr);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = i0 + 32 * q;
            if (i < N) {
                const double2 v = XY2[2 * i];
                XY2[2 * i + 1] = make_double2(pv[q].x + alr * v.x - ali * v.y, pv[q].y + alr * v.y + ali * v.x);
            }
        }
        __syncwarp();
        // --- A -= v w^H + w v^H on the tensor cores, lower 8x8 blocks that touch the trailing sub-matrix ---
        const int qb = r0 >> 3;
        const double sgn = (tq & 1) ? 1.0 : -1.0;
        const bool ragged = (N & 7) != 0;
        for (int I = qb; I < NB; ++I) {
            const int row = I * 8 + g;
            const double xr = -XY[row * 4 + tq];                 // -(vr, vi, wr, wi)[tq]
            const double xi = sgn * XY[row * 4 + (tq ^ 1)];      // (-vi, vr, -wi, wr)[tq]
            double2* rp = As + row * S + qb * 8 + 2 * tq;        // C fragment (row, J*8 + 2 tq), J = qb
            double2* mp = As + (qb * 8 + 2 * tq) * S + row;      // its mirror (J*8 + 2 tq, row)
            const double* yp = XY + (qb * 8 + g) * 4 + (tq ^ 2);  // (wr, wi, vr, vi)[tq] of row J*8 + g
            if (!(ragged && I == NB - 1)) {
                // every element of these blocks exists: no predicates
                for (int J = qb; J < I; ++J) {
                    const double y = *yp;
                    const double2 z0 = rp[0], z1 = rp[1];
                    double cre0 = z0.x, cre1 = z1.x, cim0 = z0.y, cim1 = z1.y;
                    dmma_acc(cre0, cre1, xr, y);
                    dmma_acc(cim0, cim1, xi, y);
                    rp[0] = make_double2(cre0, cim0);
                    rp[1] = make_double2(cre1, cim1);
                    mp[0] = make_double2(cre0, -cim0);
                    mp[S] = make_double2(cre1, -cim1);
                    rp += 8;
                    mp += 8 * S;
                    yp += 32;
                }
                {   // diagonal block: the product also yields its upper half; keep the diagonal exactly real
                    const double y = *yp;
                    const double2 z0 = rp[0], z1 = rp[1];
                    double cre0 = z0.x, cre1 = z1.x, cim0 = z0.y, cim1 = z1.y;
                    dmma_acc(cre0, cre1, xr, y);
                    dmma_acc(cim0, cim1, xi, y);
                    if (g == 2 * tq) cim0 = 0.0;
                    if (g == 2 * tq + 1) cim1 = 0.0;
                    rp[0] = make_double2(cre0, cim0);
                    rp[1] = make_double2(cre1, cim1);
                }
            } else {
                // last block row of a matrix whose size is not a multiple of 8: rows / columns >= N do not exist
                const bool row_ok = row < N;
                for (int J = qb; J <= I; ++J) {
                    const double y = *yp;
                    const int c0 = J * 8 + 2 * tq;
                    const bool ok0 = row_ok && c0 < N, ok1 = row_ok && c0 + 1 < N;
                    const double2 z0 = ok0 ? rp[0] : make_double2(0.0, 0.0);
                    const double2 z1 = ok1 ? rp[1] : make_double2(0.0, 0.0);
                    double cre0 = z0.x, cre1 = z1.x, cim0 = z0.y, cim1 = z1.y;
                    dmma_acc(cre0, cre1, xr, y);
                    dmma_acc(cim0, cim1, xi, y);
                    if (J == I) {
                        if (g == 2 * tq) cim0 = 0.0;
                        if (g == 2 * tq + 1) cim1 = 0.0;
                    }
                    if (ok0) rp[0] = make_double2(cre0, cim0);
                    if (ok1) rp[1] = make_double2(cre1, cim1);
                    if (J != I) {
                        if (ok0) mp[0] = make_double2(cre0, -cim0);
                        if (ok1) mp[S] = make_double2(cre1, -cim1);
                    }
                    rp += 8;
                    mp += 8 * S;
                    yp += 32;
                }
            }
        }
        __syncwarp();
    }
    if (t == 0) {
        ds[N - 1] = As[(N - 1) * S + (N - 1)].x;
        es[N - 1] = 0.0;
    }
    __syncwarp();
    if (valid) {
        for (int i = t; i < N; i += 32) {
            D[kidx * N + i] = ds[i];
            E[kidx * N + i] = es[i];
        }
    }
}

cudaError_t launch_mma(int n, double* Hp, long nk, double* D, double* E, cudaStream_t st) {
    const MmaLayout L(n);
    const size_t per_mat = (size_t)L.per_matrix * 16;
    // matrices per CTA that maximise the number resident per SM (228 KB, 1 KB reserved per CTA); ties -> smaller CTA
    int best_mpb = 1, best_res = 0;
    for (int mpb = 1; mpb <= TPB / 32; ++mpb) {
        const size_t cta = per_mat * mpb + 1024;
        if (per_mat * mpb > 227 * 1024) break;
        const int res = (int)((228 * 1024) / cta) * mpb;
        if (res > best_res) {
            best_res = res;
            best_mpb = mpb;
        }
    }
    const size_t smem = per_mat * best_mpb;
    cudaError_t err = cudaFuncSetAttribute(tridiag_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    const long blocks = (nk + best_mpb - 1) / best_mpb;
    if (blocks <= 0) return cudaSuccess;
    if (blocks > 2147483647L) return cudaErrorInvalidConfiguration;
    tridiag_mma_kernel<<<(unsigned)blocks, 32 * best_mpb, smem, st>>>(Hp, n, nk, D, E);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Large matrices (N > 164, up to 32 * MAXC): one CTA of 16 warps per matrix, in place on the packed planes in
// global memory (L2 / HBM).  Everything is organised as coalesced ROW sweeps: a warp owns rows a = w, w + 16, ...
// of the trailing block and strides its lanes over the columns.  The Hermitian matrix-vector product needs the
// column part  sum_{J > I} conj(A[J][I]) v_J  as well; instead of gathering columns, each row sweep also scatters
// conj(A[I][J]) v_I into per-lane register accumulators (one per 32-column chunk), which are combined across the
// 16 warps through shared memory in a fixed order (deterministic, no atomics).  The matrix is therefore read once
// per product and read + written once per rank-2 update.
// ---------------------------------------------------------------------------------------------------------
constexpr int BIG_THREADS = 512;
constexpr int BIG_WARPS = BIG_THREADS / 32;

__device__ __forceinline__ void block_sum2(double& a, double& b, double* red, int tid, int& parity) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
    }
    double* buf = red + parity * 2 * BIG_WARPS;
    parity ^= 1;
    if ((tid & 31) == 0) {
        buf[2 * (tid >> 5)] = a;
        buf[2 * (tid >> 5) + 1] = b;
    }
    __syncthreads();
    a = 0.0;
    b = 0.0;
#pragma unroll
    for (int w = 0; w < BIG_WARPS; ++w) {
        a += buf[2 * w];
        b += buf[2 * w + 1];
    }
}

template <int MAXC>
__global__ void __launch_bounds__(BIG_THREADS, 1)
tridiag_big_kernel(double* __restrict__ Hp, int N, long nk, double* __restrict__ D, double* __restrict__ E) {
    extern __shared__ __align__(16) double2 smb[];
    double2* V = smb;                 // [N]
    double2* P = V + N;               // [N] p, then w
    double2* S = P + N;               // [N] row-part sums
    double2* QW = S + N;              // [BIG_WARPS][N] column-part partials
    double* ds = reinterpret_cast<double*>(QW + (size_t)BIG_WARPS * N);
    double* es = ds + N;
    double* red = es + N;             // [2][2 * BIG_WARPS]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const long kk = blockIdx.x;
    const long NN = (long)N * N;
    double* Ar = Hp + kk * NN;
    double* Ai = Ar + tri(N);
    int parity = 0;

    for (int j = 0; j < N - 1; ++j) {
        const int m = N - 1 - j;
        const int r0 = j + 1;
        // --- reflector from column j (strided gather: one element per row) ---
        const double ar = Ar[tri(r0) + j], ai = Ai[trs(r0) + j];
        double xn = 0.0, dummy = 0.0;
        for (int a = 1 + tid; a < m; a += BIG_THREADS) {
            const long I = r0 + a;
            const double xr = Ar[tri(I) + j], xi = Ai[trs(I) + j];
            xn += xr * xr + xi * xi;
            V[a] = make_double2(xr, xi);  // raw column, scaled below
        }
        block_sum2(xn, dummy, red, tid, parity);
        double beta, tr, ti, sr, si;
        householder_gen(ar, ai, xn, beta, tr, ti, sr, si);
        for (int a = 1 + tid; a < m; a += BIG_THREADS) {
            const double2 x = V[a];
            V[a] = make_double2(x.x * sr - x.y * si, x.x * si + x.y * sr);
        }
        if (tid == 0) {
            ds[j] = Ar[tri(j) + j];
            es[j] = beta;
            V[0] = make_double2(1.0, 0.0);
        }
        __syncthreads();
        // --- row sweep: row parts into S, column parts into per-lane accumulators ---
        double2 qacc[MAXC];
#pragma unroll
        for (int c = 0; c < MAXC; ++c) qacc[c] = make_double2(0.0, 0.0);
        for (int a = w; a < m; a += BIG_WARPS) {
            const long I = r0 + a;
            const double* rre = Ar + tri(I) + r0;
            const double* rim = Ai + trs(I) + r0;
            const double2 va = V[a];
            double sumr = 0.0, sumi = 0.0;
#pragma unroll
            for (int c0 = 0; c0 < MAXC; c0 += 4) {
                if (c0 * 32 < a) {  // warp-uniform; four 32-column chunks per batch keep 8 loads in flight per lane
                    double zr[4], zi[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int b = (c0 + q) * 32 + lane;
                        const bool ok = b < a;
                        zr[q] = ok ? rre[b] : 0.0;
                        zi[q] = ok ? rim[b] : 0.0;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int b = (c0 + q) * 32 + lane;
                        if (b < a) {
                            const double2 vb = V[b];
                            sumr = fma(zr[q], vb.x, fma(-zi[q], vb.y, sumr));
                            sumi = fma(zr[q], vb.y, fma(zi[q], vb.x, sumi));
                            qacc[c0 + q].x = fma(zr[q], va.x, fma(zi[q], va.y, qacc[c0 + q].x));   // conj(z) * v_a
                            qacc[c0 + q].y = fma(zr[q], va.y, fma(-zi[q], va.x, qacc[c0 + q].y));
                        }
                    }
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                sumr += __shfl_xor_sync(0xffffffffu, sumr, off);
                sumi += __shfl_xor_sync(0xffffffffu, sumi, off);
            }
            if (lane == 0) {
                const double dg = rre[a];  // real diagonal
                S[a] = make_double2(fma(dg, va.x, sumr), fma(dg, va.y, sumi));
            }
        }
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            const int b = c * 32 + lane;
            if (b < m) QW[(size_t)w * N + b] = qacc[c];
        }
        __syncthreads();
        double dr = 0.0, di = 0.0;
        for (int a = tid; a < m; a += BIG_THREADS) {
            double2 q = S[a];
#pragma unroll
            for (int ww = 0; ww < BIG_WARPS; ++ww) {
                const double2 t2 = QW[(size_t)ww * N + a];
                q.x += t2.x;
                q.y += t2.y;
            }
            const double pr = tr * q.x - ti * q.y;
            const double pi = tr * q.y + ti * q.x;
            P[a] = make_double2(pr, pi);
            const double2 va = V[a];
            dr += pr * va.x + pi * va.y;
            di += pr * va.y - pi * va.x;
        }
        block_sum2(dr, di, red, tid, parity);
        const double alr = -0.5 * (tr * dr - ti * di);
        const double ali = -0.5 * (tr * di + ti * dr);
        for (int a = tid; a < m; a += BIG_THREADS) {
            const double2 v = V[a];
            double2 pq = P[a];
            pq.x += alr * v.x - ali * v.y;
            pq.y += alr * v.y + ali * v.x;
            P[a] = pq;
        }
        __syncthreads();
        // --- rank-2 update, row sweeps ---
        for (int a = w; a < m; a += BIG_WARPS) {
            const long I = r0 + a;
            double* rre = Ar + tri(I) + r0;
            double* rim = Ai + trs(I) + r0;
            const double2 va = V[a], wa = P[a];
            for (int b0 = 0; b0 < a; b0 += 128) {  // batches of four chunks: loads first, then arithmetic and stores
                double zr[4], zi[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int b = b0 + q * 32 + lane;
                    const bool ok = b < a;
                    zr[q] = ok ? rre[b] : 0.0;
                    zi[q] = ok ? rim[b] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int b = b0 + q * 32 + lane;
                    if (b < a) {
                        const double2 vb = V[b], wb = P[b];
                        rre[b] = fma(-va.x, wb.x, fma(-va.y, wb.y, fma(-wa.x, vb.x, fma(-wa.y, vb.y, zr[q]))));
                        rim[b] = fma(-va.y, wb.x, fma(va.x, wb.y, fma(-wa.y, vb.x, fma(wa.x, vb.y, zi[q]))));
                    }
                }
            }
            if (lane == 0) rre[a] -= 2.0 * (va.x * wa.x + va.y * wa.y);
        }
        __syncthreads();
        __threadfence_block();
    }
    if (tid == 0) {
        ds[N - 1] = Ar[tri(N - 1) + (N - 1)];
        es[N - 1] = 0.0;
    }
    __syncthreads();
    for (int i = tid; i < N; i += BIG_THREADS) {
        D[kk * N + i] = ds[i];
        E[kk * N + i] = es[i];
    }
}

template <int MAXC>
cudaError_t launch_big(int n, double* Hp, long nk, double* D, double* E, cudaStream_t st) {
    const size_t smem = ((size_t)(3 + BIG_WARPS) * n * 16) + (size_t)2 * n * 8 + 4 * BIG_WARPS * 8 + 64;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t err =
        cudaFuncSetAttribute(tridiag_big_kernel<MAXC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    if (nk <= 0) return cudaSuccess;
    if (nk > 2147483647L) return cudaErrorInvalidConfiguration;
    tridiag_big_kernel<MAXC><<<(unsigned)nk, BIG_THREADS, smem, st>>>(Hp, n, nk, D, E);
    return cudaGetLastError();
}

template <int G, int CS>
cudaError_t launch_g(int n, double* Hp, long nk, double* D, double* E, cudaStream_t st, const Tuning& tune, long mstride = 0,
                     int ldo = 0, int off = 0, int nsteps = 1 << 30) {
    if (mstride == 0) mstride = (long)n * n;
    if (ldo == 0) ldo = n;
    constexpr int NW = G > 32 ? G / 32 : 1;
    constexpr int CTA = G > TPB ? G : TPB;
    const long ntri = (long)n * (n + 1) / 2;
    const size_t per_mat = (size_t)(ntri + 3L * n + (CS > 1 ? (long)CS * n : 0) + 2 * NW + 1) * 16;  // see the kernel
    if (per_mat <= kSmemLimit) {
        // matrices per CTA that maximise the number resident per SM (228 KB, 1 KB reserved per CTA, 2048 threads,
        // 15 named barriers per CTA); ties -> smaller CTA
        int best_mpb = 1, best_res = 0;
        for (int mpb = 1; mpb <= CTA / G && mpb <= 15; ++mpb) {
            if (per_mat * mpb > kSmemLimit) break;
            int ctas = (int)((228 * 1024) / (per_mat * mpb + 1024));
            const int by_threads = 2048 / (G * mpb);
            if (ctas > by_threads) ctas = by_threads;
            if (ctas * mpb > best_res) {
                best_res = ctas * mpb;
                best_mpb = mpb;
            }
        }
        {   // tuning hook: matrices per CTA
            const int v = tune.tridiag_mpb;
            if (v >= 1 && v <= 15 && (size_t)v * G <= 1024 && per_mat * v <= kSmemLimit) best_mpb = v;
        }
        const size_t smem = per_mat * best_mpb;
        cudaError_t err = cudaFuncSetAttribute(tridiag_smem_kernel<G, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smem);
        if (err != cudaSuccess) return err;
        const long blocks = (nk + best_mpb - 1) / best_mpb;
        if (blocks <= 0) return cudaSuccess;
        if (blocks > 2147483647L) return cudaErrorInvalidConfiguration;
        tridiag_smem_kernel<G, CS><<<(unsigned)blocks, G * best_mpb, smem, st>>>(Hp, n, mstride, nk, D, E, ldo, off, nsteps);
        return cudaGetLastError();
    }
    // matrix does not fit in shared memory: in place on the packed scratch (L2 / HBM).  The kernels below reduce a whole
    // matrix stored at its natural stride; a stage of the staged reduction (trailing block, partial step count) must never
    // land here -- fail loudly instead of computing on the wrong layout (a tuning-hook combination can ask for it)
    if (mstride != (long)n * n || ldo != n || off != 0 || nsteps < n - 1) return cudaErrorInvalidConfiguration;
    if (!tune.tridiag_nopanel && tridiag_panel_fits(n)) return launch_tridiag_panel(n, Hp, nk, D, E, st, tune);
    if (!tune.tridiag_oldbig) {
        if (n <= 32 * 16) return launch_big<16>(n, Hp, nk, D, E, st);
        if (n <= 32 * 20) return launch_big<20>(n, Hp, nk, D, E, st);  // shared memory: (3 + 16) * 16 N bytes <= 227 KB
    }
    const size_t smem = (size_t)(6L * n + 4 * 8 + 2) * 8;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;  // N > ~4800
    cudaError_t err = cudaFuncSetAttribute(tridiag_global_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    if (nk <= 0) return cudaSuccess;
    if (nk > 2147483647L) return cudaErrorInvalidConfiguration;
    tridiag_global_kernel<<<(unsigned)nk, TPB, smem, st>>>(Hp, n, nk, D, E);
    return cudaGetLastError();
}

}  // namespace

// Thread group (threads per matrix, column slices) of the shared-memory kernel for a matrix / trailing block of size n,
// and how many such matrices one SM then holds (the arithmetic of launch_g).
void smem_default_group(int n, int& gg, int& cc) {
    if (n <= 10) gg = 8;
    else if (n <= 20) gg = 16;
    else if (n <= 48) gg = 32;
    else if (n <= 96) gg = 128;
    else if (n <= 112) gg = 256;
    else gg = 512;  // one matrix per SM: 16 warps on it (measured: N = 128 2.66 -> 2.50, 144 4.12 -> 3.87, 160 5.25 -> 5.02 ms
                    // per 1000 matrices, gpurun_out/r03d_sweep.log)
    cc = n <= 48 ? 1 : (n <= 96 ? 2 : 4);  // measured on B200: N = 36 best at (32, 1), N = 128 at (512, 4)
}

int smem_residency(int n) {
    int G, CS;
    smem_default_group(n, G, CS);
    const int NW = G > 32 ? G / 32 : 1;
    const int CTA = G > TPB ? G : TPB;
    const size_t per_mat = (size_t)((long)n * (n + 1) / 2 + 3L * n + (CS > 1 ? (long)CS * n : 0) + 2 * NW + 1) * 16;
    if (per_mat > kSmemLimit) return 0;
    int best = 0;
    for (int mpb = 1; mpb <= CTA / G && mpb <= 15; ++mpb) {
        if (per_mat * mpb > kSmemLimit) break;
        int ctas = (int)((228 * 1024) / (per_mat * mpb + 1024));
        const int by_threads = 2048 / (G * mpb);
        if (ctas > by_threads) ctas = by_threads;
        if (ctas * mpb > best) best = ctas * mpb;
    }
    return best;
}

constexpr int kTwoStageMinN = 128;  // two-stage reduction (eig_band.cu) from this size on, see tridiag_twostage_default
constexpr int kPanelMinN = 161;   // blocked kernel from this size on (re-measured in round 2, see below)
constexpr int kStagedMaxN = 160;  // staged shared-memory reduction while the packed matrix + work vectors fit 227 KB

// Sizes the two-stage reduction (eig_band.cu) serves: TBK_TRIDIAG_TWOSTAGE = first such N (0 = never).  The hooks that force
// one of the one-stage kernels win.  Measured on B200, ms per 1000 matrices, one-stage (staged shared-memory kernels up to
// 160, blocked kernel + staged tail above) / two-stage, batches of 9 472 - 37 888 matrices (profiles/r05a_twostage_small_n_sweep.log,
// r04z_twostage_threshold_sweep.log, r04w_onestage_vs_twostage_sweep.log): N = 96: 1.05 / 1.25, 112: 1.54 / 1.63,
// 128: 2.38 / 2.05, 144: 3.59 / 2.57, 160: 4.99 / 3.20, 200: 7.97 / 5.35, 224: 11.1 / 6.82, 256: 16.4 / 9.54, 320: 34.9 / 21.2,
// 512: 117.2 / 62.9, 640: 561 / 154, 700: 2328 / 202 (the blocked kernel ends at N = 640).  Batches of a few hundred matrices
// (296): N = 128: 3.03 / 3.73, 160: 5.52 / 5.58, 224: 11.8 / 10.9, 256: 16.9 / 14.6, 512: 117 / 81.9.  The choice may depend
// on N only (results must not depend on the batch): the two-stage reduction serves every size where it wins on large batches.
bool tridiag_twostage_default(int n, const Tuning& tune) {
    const int from = tune.tridiag_twostage >= 0 ? tune.tridiag_twostage : kTwoStageMinN;
    return from > 0 && n >= from && tune.tridiag_g == 0 && tune.tridiag_panel_min == 0 && !tune.tridiag_nopanel &&
           tridiag_twostage_fits(n);
}

cudaError_t launch_tridiag(int n, double* Hp, long nk, double* D, double* E, cudaStream_t st, const Tuning& tune) {
    const int g = tune.tridiag_g, cs = tune.tridiag_cs;  // tuning hooks: threads per matrix / column slices
    if (g == 1) return launch_mma(n, Hp, nk, D, E, st);
    const int panel_from = tune.tridiag_panel_min > 0 ? tune.tridiag_panel_min : kPanelMinN;  // tuning hook
    if (tune.tridiag_panel_min > 0 && n >= panel_from && tridiag_panel_fits(n))
        return launch_tridiag_panel(n, Hp, nk, D, E, st, tune);
    // (the tensor-core variant launch_mma is correct for n <= 64 but, with only ~9 single-warp CTAs resident per
    //  SM, it is latency bound and measured 35 % slower than the packed kernel on B200: opt-in via TBK_TRIDIAG_G=1)
    // blocked kernel (eig_tridiag_panel.cu) where the packed matrix no longer fits in shared memory (N > 160).  Measured
    // on B200, ms per 1000 matrices, single-launch smem / blocked (round 1): N = 100: 1.47 / 1.95, 128: 3.65 / 3.31,
    // 164: 9.9 / 6.7, 256: 30.7 / 19.0, 384: 110 / 56, 512: 275 / 125; round 2, STAGED smem reduction ending in the register
    // kernels / blocked: N = 120: 2.46 / 2.74, 128: 2.84 / 3.11, 136: 3.38 / 4.22, 144: 4.16 / 4.71, 160: 5.30 / 5.81
    // (gpurun_out/r02q_sweep.log) -- so the staged reduction now serves every size that fits.
    int cur = n, done = 0;
    if (g == 0 && n >= panel_from && !tune.tridiag_nopanel && tridiag_panel_fits(n)) {
        // The blocked kernel is built for matrices that stream from L2 / HBM; once the trailing block fits in shared
        // memory its steps are all latency (one 512-thread CTA per small block).  It therefore stops at the first panel
        // boundary with at most `limit` rows left and the staged kernels below take over.
        // Where: as soon as the staged kernels hold at least as many matrices per SM as the blocked kernel does -- two for
        // N <= 256 (staged blocks <= 112), one above (<= 160).  Measured ms per 1000 matrices without / with the hand-over:
        // N = 164: 6.13 / 4.78, 200: 9.37 / 7.99, 256: 17.5 / 16.2, 384: 58.2 / 54.5, 512: 131.4 / 127.5 (r03o_sweep.log)
        int limit = tune.tridiag_panel_stop >= 0 ? tune.tridiag_panel_stop : (n <= 256 ? 112 : 160);
        if (limit > kStagedMaxN) limit = kStagedMaxN;
        const int np = tune.tridiag_stages != 0 ? tridiag_panel_handover(n, limit) : 0;
        if (np < 25) return launch_tridiag_panel(n, Hp, nk, D, E, st, tune);
        const cudaError_t err = launch_tridiag_panel(n, Hp, nk, D, E, st, tune, limit);
        if (err != cudaSuccess) return err;
        cur = np;
        done = n - np;
    }
    const int n0 = cur;  // size the staged plan starts from
    // register-resident warp-per-matrix kernel (eig_tridiag_reg.cu): the whole reduction in one launch
    const auto use_reg = [&](int m) {
        return g == 0 && tune.tridiag_reg_max > 0 && m >= tune.tridiag_reg_min && m <= tune.tridiag_reg_max &&
               tridiag_reg_fits(m);
    };
    if (use_reg(n)) return launch_tridiag_reg(n, Hp, nk, D, E, st, 0, 0, 0, tune.tridiag_reg_bw, tune.tridiag_reg_stop, tune.tridiag_reg_mid);
    // Staged reduction (shared-memory kernels, 25 <= N < 120): the trailing block shrinks, so after every stage the
    // remaining (smaller) problem is relaunched with several times more matrices resident per SM -- shared memory per
    // matrix ~ N^2 caps residency and the kernel is latency bound.  Stage sizes follow from N only (results never depend
    // on the batch): N -> ratio * N -> ... until <= 16.  TBK_TRIDIAG_STAGES="0" disables, "p" sets the ratio in percent.
    const int ratio = tune.tridiag_stages >= 0 ? tune.tridiag_stages : ((n0 >= 88 && n0 <= 140) ? 80 : 67);
    const bool staged = g == 0 && n0 >= 25 && n0 <= kStagedMaxN && ratio > 0 && ratio < 100;  // (above: not in shared memory)
    const long ms = (long)n * n;
    for (;;) {
        int next = 0;
        if (staged && cur > 16) {
            next = (cur * ratio + 50) / 100;
            if (tune.tridiag_stages < 0 && cur > 88) {  // (below: the ratio rule measured 1 - 3 % better)
                // default plan for the large sizes: the kernel is latency bound and its speed is the number of matrices an
                // SM holds, so a stage ends exactly where one more matrix fits (at least 8 steps per stage, and never
                // later than the ratio rule would end it: a stage is also one more pass over the data)
                const int res = smem_residency(cur);
                for (int cand = cur - 8; cand >= next && cand > 40; --cand)
                    if (smem_residency(cand) > res) {
                        next = cand;
                        break;
                    }
            }
            if (next < 12) next = 12;
            if (next >= cur) next = 0;
            // hand over to the register kernel at its largest size instead of staging past it
            const int rm = tune.tridiag_reg_max;
            if (g == 0 && rm > 0 && cur > rm && next > 0 && next <= rm + rm / 5 && use_reg(rm)) next = rm;
        }
        // a trailing block the register kernel serves is finished there (same lower-storage reduction, one launch)
        if (cur != n && use_reg(cur))
            return launch_tridiag_reg(cur, Hp, nk, D, E, st, ms, n, done, tune.tridiag_reg_bw, tune.tridiag_reg_stop, tune.tridiag_reg_mid);
        const int nsteps = next ? cur - next : (1 << 30);
        int gg = g, cc = cs;
        if (gg == 0) {  // defaults, from the size only
            smem_default_group(cur, gg, cc);
            if (cur > 112 && tune.tridiag_g1 > 0) {  // tuning hook: thread group of the one-matrix-per-SM stages
                gg = tune.tridiag_g1;
                cc = tune.tridiag_cs1 > 0 ? tune.tridiag_cs1 : 4;
            }
        }
        cudaError_t err = cudaErrorInvalidValue;
#define TBK_CASE(G_, CS_) \
    if (gg == G_ && cc == CS_) err = launch_g<G_, CS_>(cur, Hp, nk, D, E, st, tune, ms, n, done, nsteps)
        TBK_CASE(8, 1);
        TBK_CASE(16, 1);
        TBK_CASE(32, 1);
        TBK_CASE(64, 1);
        TBK_CASE(64, 2);
        TBK_CASE(128, 1);
        TBK_CASE(128, 2);
        TBK_CASE(128, 4);
        TBK_CASE(256, 1);
        TBK_CASE(256, 2);
        TBK_CASE(256, 4);
        TBK_CASE(256, 8);
        TBK_CASE(512, 4);
        TBK_CASE(512, 8);
#undef TBK_CASE
        if (err != cudaSuccess || !next) return err;
        done += cur - next;
        cur = next;
    }
}

}  // namespace tbk
