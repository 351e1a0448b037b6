// eig_ql.cu -- batched eigenvalues of real symmetric tridiagonal matrices, one thread per matrix.
//
// Second half of the replacement for scipy.linalg.eigvalsh in Model.eigenval (reference
// src/tbmodels/_tb_model.py:1148-1149).  The QL sweep is a serial recurrence, so parallelism comes from
// the batch: each lane runs tbk::tridiag_ql on its own (d, e).  A CTA of T threads (128 for small N down to 8
// for N = 512: shared memory per thread is 2 N doubles) stages its T matrices through shared memory with
// coalesced global loads/stores and a thread-strided (conflict-free) layout; only when even 8 matrices do not
// fit does each thread work in place on its row of D/E in global memory.
// N > 128 switches to bisection (bisect_kernel below).  Eigenvalues come out ascending, like LAPACK's.
#include <algorithm>

#include "tbk_kernels.h"
#include "tbk_math.cuh"

namespace tbk {

namespace {

constexpr int TPB_GLOBAL = 128;

// T matrices (threads) per CTA.  Shared memory per thread is 2 N doubles, so large N means small CTAs; the odd row
// stride T + 1 keeps the transposing loads/stores conflict free.
template <int T>
__global__ void __launch_bounds__(T)
ql_smem_kernel(double* __restrict__ D, double* __restrict__ E, int N, long nk, int* __restrict__ fail_count) {
    constexpr int LDS = T + 1;
    extern __shared__ __align__(16) double sm[];
    double* ds = sm;                  // [N][LDS]
    double* es = sm + (size_t)N * LDS;
    const long k0 = (long)blockIdx.x * T;
    const int nmat = (int)((nk - k0) < T ? (nk - k0) : T);
    const int tid = threadIdx.x;
    const long total = (long)nmat * N;
    const double* Dg = D + k0 * N;
    const double* Eg = E + k0 * N;
    for (long idx = tid; idx < total; idx += T) {
        const int mat = (int)(idx / N), i = (int)(idx - (long)mat * N);
        ds[i * LDS + mat] = Dg[idx];
        es[i * LDS + mat] = Eg[idx];
    }
    __syncthreads();
    if (tid < nmat) {
        const int fails = tridiag_ql(N, ds + tid, es + tid, LDS);
        if (fails && fail_count) atomicAdd(fail_count, fails);
    }
    __syncthreads();
    double* Do = D + k0 * N;
    for (long idx = tid; idx < total; idx += T) {
        const int mat = (int)(idx / N), i = (int)(idx - (long)mat * N);
        Do[idx] = ds[i * LDS + mat];
    }
}

// Background variant: a FEW persistent CTAs (sized to fit beside the H(k) GEMM's shared-memory ring) pull groups of T
// matrices from an atomic counter.  tbk_api.cu runs it on a side stream for chunk i while the main stream builds and
// tridiagonalises chunk i + 1: the QL recurrence is latency bound (one dependent FP64 chain per thread), so it costs the
// co-resident kernels almost nothing, and its 12 % of the step disappears behind them.  Same per-matrix arithmetic as
// ql_smem_kernel (bits do not depend on which CTA picks a matrix up).
template <int T>
__global__ void __launch_bounds__(T)
ql_background_kernel(double* __restrict__ D, double* __restrict__ E, int N, long nk, int* __restrict__ fail_count,
                     unsigned long long* __restrict__ counter) {
    constexpr int LDS = T + 1;
    extern __shared__ __align__(16) double sm[];
    double* ds = sm;
    double* es = sm + (size_t)N * LDS;
    __shared__ long s_base;
    const int tid = threadIdx.x;
    for (;;) {
        if (tid == 0) s_base = (long)atomicAdd(counter, (unsigned long long)T);
        __syncthreads();
        const long k0 = s_base;
        if (k0 >= nk) break;
        const int nmat = (int)((nk - k0) < T ? (nk - k0) : T);
        const long total = (long)nmat * N;
        const double* Dg = D + k0 * N;
        const double* Eg = E + k0 * N;
        for (long idx = tid; idx < total; idx += T) {
            const int mat = (int)(idx / N), i = (int)(idx - (long)mat * N);
            ds[i * LDS + mat] = Dg[idx];
            es[i * LDS + mat] = Eg[idx];
        }
        __syncthreads();
        if (tid < nmat) {
            const int fails = tridiag_ql(N, ds + tid, es + tid, LDS);
            if (fails && fail_count) atomicAdd(fail_count, fails);
        }
        __syncthreads();
        double* Do = D + k0 * N;
        for (long idx = tid; idx < total; idx += T) {
            const int mat = (int)(idx / N), i = (int)(idx - (long)mat * N);
            Do[idx] = ds[i * LDS + mat];
        }
        __syncthreads();  // s_base and the staging buffers are reused by the next group
    }
}

__global__ void __launch_bounds__(TPB_GLOBAL)
ql_global_kernel(double* __restrict__ D, double* __restrict__ E, int N, long nk, int* __restrict__ fail_count) {
    const long kk = (long)blockIdx.x * TPB_GLOBAL + threadIdx.x;
    if (kk >= nk) return;
    const int fails = tridiag_ql(N, D + kk * N, E + kk * N, 1);
    if (fails && fail_count) atomicAdd(fail_count, fails);
}

// Large N (> 128): the serial QL recurrence leaves the GPU idle (shared memory allows only a few threads per SM), so
// the eigenvalues are found by bisection instead -- one CTA per matrix, one thread per eigenvalue, every thread runs
// the same Sturm recurrence on the broadcast (d, e^2) in shared memory.  ~50 N^2 divisions per matrix, all parallel.
template <int TPB_BISECT>
__global__ void __launch_bounds__(TPB_BISECT)
bisect_kernel(double* __restrict__ D, const double* __restrict__ E, int N, long nk) {
    extern __shared__ __align__(16) double sm[];
    double* ds = sm;
    double* e2 = sm + N;
    __shared__ double red[4][TPB_BISECT / 32];
    const long kk = blockIdx.x;
    const int tid = threadIdx.x;
    double lo = 1e300, hi = -1e300, emax = 0.0;
    for (int i = tid; i < N; i += TPB_BISECT) {
        const double di = D[kk * N + i];
        const double el = i > 0 ? fabs(E[kk * N + i - 1]) : 0.0;
        const double er = i + 1 < N ? fabs(E[kk * N + i]) : 0.0;
        ds[i] = di;
        e2[i] = er * er;
        lo = fmin(lo, di - el - er);  // Gershgorin
        hi = fmax(hi, di + el + er);
        emax = fmax(emax, er * er);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
        emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, off));
    }
    if ((tid & 31) == 0) {
        red[0][tid >> 5] = lo;
        red[1][tid >> 5] = hi;
        red[2][tid >> 5] = emax;
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < TPB_BISECT / 32; ++w) {
        lo = fmin(lo, red[0][w]);
        hi = fmax(hi, red[1][w]);
        emax = fmax(emax, red[2][w]);
    }
    if (emax == 0.0) {
        // already diagonal (e.g. the empty model: the reference returns exact zeros): the eigenvalues are the diagonal
        // entries, sorted by rank -- exact, where bisection would return values within pivmin of them
        for (int i = tid; i < N; i += TPB_BISECT) {
            const double di = ds[i];
            int rank = 0;
            for (int q = 0; q < N; ++q) {
                const double dq = ds[q];
                rank += (dq < di || (dq == di && q < i)) ? 1 : 0;
            }
            D[kk * N + rank] = di;
        }
        return;
    }
    const double pivmin = DBL_MIN * fmax(1.0, emax);
    const double span = fmax(fabs(lo), fabs(hi));
    const double gl = lo - 2.0 * DBL_EPSILON * span * N - 2.0 * pivmin;
    const double gu = hi + 2.0 * DBL_EPSILON * span * N + 2.0 * pivmin;
    for (int i = tid; i < N; i += TPB_BISECT) D[kk * N + i] = bisect_eig(N, ds, e2, i, gl, gu, pivmin);
}

size_t ql_smem_bytes(int n, int t) { return (size_t)2 * n * (t + 1) * 8; }

// Largest CTA that still lets three (else two, else one) CTAs share an SM; 0 -> global-memory fallback.
int ql_pick_threads(int n) {
    const int cand[5] = {128, 64, 32, 16, 8};
    const size_t limits[3] = {72 * 1024, 110 * 1024, 220 * 1024};
    for (size_t lim : limits)
        for (int t : cand)
            if (ql_smem_bytes(n, t) <= lim) return t;
    return 0;
}

template <int T>
cudaError_t launch_t(int n, double* D, double* E, long nk, int* fail_count, cudaStream_t st, long* wave) {
    const size_t smem = ql_smem_bytes(n, T);
    cudaError_t err = cudaFuncSetAttribute(ql_smem_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    if (wave) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ql_smem_kernel<T>, T, smem);
        *wave = (err == cudaSuccess) ? (long)sms * per_sm * T : 0;
        return err;
    }
    const long blocks = (nk + T - 1) / T;
    if (blocks > 2147483647L) return cudaErrorInvalidConfiguration;
    ql_smem_kernel<T><<<(unsigned)blocks, T, smem, st>>>(D, E, n, nk, fail_count);
    return cudaGetLastError();
}

// Crossover measured on B200 (tools/tridiag_sweep.py default,bisect; ms per 1000 matrices, QL / bisection with 256 threads):
// N = 96: 0.204 / 0.387, 112: 0.542 / 0.603, 128: 0.701 / 0.691 -- the QL kernel falls off an occupancy cliff above N = 96
// (batch of 8192 = a partial wave); 128-thread CTAs for N <= 128 measured the same as 256 (0.688 at N = 128), so QL stays
// up to N = 128.
constexpr int kBisectMinN = 129;

template <int TPB>
cudaError_t launch_bisect(int n, double* D, const double* E, long nk, cudaStream_t st) {
    const size_t smem = (size_t)2 * n * 8;
    cudaError_t err = cudaFuncSetAttribute(bisect_kernel<TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    if (nk > 2147483647L) return cudaErrorInvalidConfiguration;
    bisect_kernel<TPB><<<(unsigned)nk, TPB, smem, st>>>(D, E, n, nk);
    return cudaGetLastError();
}

cudaError_t dispatch(int n, double* D, double* E, long nk, int* fail_count, cudaStream_t st, long* wave, const Tuning& tune) {
    const int bisect_min = tune.ql_bisect_min > 0 ? tune.ql_bisect_min : kBisectMinN;
    if (tune.ql_global_min > 0 && n >= tune.ql_global_min) {  // experiment hook: thread-per-matrix QL in place in global memory
        if (wave) {
            *wave = 0;
            return cudaSuccess;
        }
        const long blocks = (nk + TPB_GLOBAL - 1) / TPB_GLOBAL;
        if (blocks > 2147483647L) return cudaErrorInvalidConfiguration;
        ql_global_kernel<<<(unsigned)blocks, TPB_GLOBAL, 0, st>>>(D, E, n, nk, fail_count);
        return cudaGetLastError();
    }
    if (n >= bisect_min && (size_t)2 * n * 8 <= 200 * 1024) {
        if (wave) {
            *wave = 0;  // one CTA per matrix: no wave quantisation to respect
            return cudaSuccess;
        }
        if (n <= 64) return launch_bisect<64>(n, D, E, nk, st);
        if (n <= 128) return launch_bisect<128>(n, D, E, nk, st);
        return launch_bisect<256>(n, D, E, nk, st);
    }
    switch (ql_pick_threads(n)) {
        case 128: return launch_t<128>(n, D, E, nk, fail_count, st, wave);
        case 64: return launch_t<64>(n, D, E, nk, fail_count, st, wave);
        case 32: return launch_t<32>(n, D, E, nk, fail_count, st, wave);
        case 16: return launch_t<16>(n, D, E, nk, fail_count, st, wave);
        case 8: return launch_t<8>(n, D, E, nk, fail_count, st, wave);
        default: break;
    }
    if (wave) {
        *wave = 0;
        return cudaSuccess;
    }
    const long blocks = (nk + TPB_GLOBAL - 1) / TPB_GLOBAL;
    if (blocks > 2147483647L) return cudaErrorInvalidConfiguration;
    ql_global_kernel<<<(unsigned)blocks, TPB_GLOBAL, 0, st>>>(D, E, n, nk, fail_count);
    return cudaGetLastError();
}

template <int T>
cudaError_t launch_bg_t(int n, double* D, double* E, long nk, int* fail_count, unsigned long long* counter, int ctas,
                        cudaStream_t st) {
    const size_t smem = ql_smem_bytes(n, T);
    cudaError_t err = cudaFuncSetAttribute(ql_background_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    err = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st);
    if (err != cudaSuccess) return err;
    const long groups = (nk + T - 1) / T;
    const unsigned grid = (unsigned)std::min<long>(groups, ctas);
    ql_background_kernel<T><<<grid, T, smem, st>>>(D, E, n, nk, fail_count, counter);
    return cudaGetLastError();
}

}  // namespace

// Background QL (see ql_background_kernel): CTAs that together use at most smem_budget bytes of shared memory per SM.
// Returns cudaErrorNotSupported when the size is served by bisection or no CTA shape fits the budget -- the caller then
// runs launch_ql in stream order as before.
cudaError_t launch_ql_background(int n, double* D, double* E, long nk, int* fail_count, unsigned long long* counter,
                                 size_t smem_budget, cudaStream_t st, const Tuning& tune) {
    if (nk <= 0 || n <= 0) return cudaSuccess;
    const int bisect_min = tune.ql_bisect_min > 0 ? tune.ql_bisect_min : kBisectMinN;
    if (n >= bisect_min || (tune.ql_global_min > 0 && n >= tune.ql_global_min)) return cudaErrorNotSupported;
    const int full_t = ql_pick_threads(n);
    if (full_t == 0) return cudaErrorNotSupported;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int cand[4] = {64, 32, 16, 8};
    for (int t : cand) {
        if (t > full_t) continue;
        const size_t per = ql_smem_bytes(n, t) + 1024;
        const int per_sm = (int)(smem_budget / per);
        if (per_sm < 1) continue;
        const int ctas = sms * per_sm;
        switch (t) {
            case 64: return launch_bg_t<64>(n, D, E, nk, fail_count, counter, ctas, st);
            case 32: return launch_bg_t<32>(n, D, E, nk, fail_count, counter, ctas, st);
            case 16: return launch_bg_t<16>(n, D, E, nk, fail_count, counter, ctas, st);
            default: return launch_bg_t<8>(n, D, E, nk, fail_count, counter, ctas, st);
        }
    }
    return cudaErrorNotSupported;
}

long ql_wave_matrices(int n, const Tuning& tune) {
    // matrices one full wave of the shared-memory QL kernel processes (0: not applicable)
    long wave = 0;
    if (n <= 0 || dispatch(n, nullptr, nullptr, 0, nullptr, nullptr, &wave, tune) != cudaSuccess) return 0;
    return wave;
}

cudaError_t launch_ql(int n, double* D, double* E, long nk, int* fail_count, cudaStream_t st, const Tuning& tune) {
    if (nk <= 0 || n <= 0) return cudaSuccess;
    return dispatch(n, D, E, nk, fail_count, st, nullptr, tune);
}

}  // namespace tbk
