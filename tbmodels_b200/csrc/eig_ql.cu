// eig_ql.cu -- batched eigenvalues of real symmetric tridiagonal matrices, one thread per matrix.
//
// Second half of the replacement for scipy.linalg.eigvalsh in Model.eigenval (reference
// src/tbmodels/_tb_model.py:1148-1149).  The QL sweep is a serial recurrence, so parallelism comes from
// the batch: each lane runs tbk::tridiag_ql on its own (d, e).  For N <= 100 a CTA of 128 threads stages
// its 128 matrices through shared memory with coalesced global loads/stores and a thread-strided
// (conflict-free) layout; above that each thread works in place on its row of D/E in global memory.
// Eigenvalues come out ascending, like LAPACK's.
#include "tbk_kernels.h"
#include "tbk_math.cuh"

namespace tbk {

namespace {

constexpr int TPB = 128;
constexpr int LDS = TPB + 1;  // odd row stride: the transposing loads/stores are conflict-free too

__global__ void __launch_bounds__(TPB)
ql_smem_kernel(double* __restrict__ D, double* __restrict__ E, int N, long nk, int* __restrict__ fail_count) {
    extern __shared__ __align__(16) double sm[];
    double* ds = sm;                  // [N][LDS]
    double* es = sm + (size_t)N * LDS;
    const long k0 = (long)blockIdx.x * TPB;
    const int nmat = (int)((nk - k0) < TPB ? (nk - k0) : TPB);
    const int tid = threadIdx.x;
    const long total = (long)nmat * N;
    const double* Dg = D + k0 * N;
    const double* Eg = E + k0 * N;
    for (long idx = tid; idx < total; idx += TPB) {
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
    for (long idx = tid; idx < total; idx += TPB) {
        const int mat = (int)(idx / N), i = (int)(idx - (long)mat * N);
        Do[idx] = ds[i * LDS + mat];
    }
}

__global__ void __launch_bounds__(TPB)
ql_global_kernel(double* __restrict__ D, double* __restrict__ E, int N, long nk, int* __restrict__ fail_count) {
    const long kk = (long)blockIdx.x * TPB + threadIdx.x;
    if (kk >= nk) return;
    const int fails = tridiag_ql(N, D + kk * N, E + kk * N, 1);
    if (fails && fail_count) atomicAdd(fail_count, fails);
}

}  // namespace

long ql_wave_matrices(int n) {
    // matrices one full wave of the shared-memory QL kernel processes (0: not applicable)
    const size_t smem = (size_t)2 * n * LDS * 8;
    if (n <= 0 || smem > 210 * 1024) return 0;
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(ql_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ql_smem_kernel, TPB, smem) != cudaSuccess) return 0;
    return (long)sms * per_sm * TPB;
}

cudaError_t launch_ql(int n, double* D, double* E, long nk, int* fail_count, cudaStream_t st) {
    if (nk <= 0 || n <= 0) return cudaSuccess;
    const long blocks = (nk + TPB - 1) / TPB;
    if (blocks > 2147483647L) return cudaErrorInvalidConfiguration;
    const size_t smem = (size_t)2 * n * LDS * 8;
    if (smem <= 210 * 1024) {
        cudaError_t err = cudaFuncSetAttribute(ql_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        ql_smem_kernel<<<(unsigned)blocks, TPB, smem, st>>>(D, E, n, nk, fail_count);
    } else {
        ql_global_kernel<<<(unsigned)blocks, TPB, 0, st>>>(D, E, n, nk, fail_count);
    }
    return cudaGetLastError();
}

}  // namespace tbk
