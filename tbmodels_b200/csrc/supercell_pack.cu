// supercell_pack.cu -- device-side packing of a supercell model.
//
// Model.supercell (reference src/tbmodels/_tb_model.py:1645-1724) copies every hopping matrix of the base model into
// the (cell a, cell b) block of a new N' x N' matrix, N' = N * prod(size), one new lattice vector R' per distinct
// floor((cell + R) / size); the constructor then folds R' onto the half set (contains_cc=False, :281-298).  The result
// is block sparse (BASELINE C4: 5.5 % of the entries), yet the reference -- and a host-side pack -- materialise
// n_R' dense N' x N' matrices (58.7 MB for C4, growing with N'^2).  Here the host only enumerates the block list
// (which base matrix lands in which block of which R', transposed-conjugated or not, scaled by 1/2 for the R' = 0
// symmetrisation); this kernel gathers the Hermitian-split weights
//     W[2q] = hp(T_q + T_q^H),   W[2q+1] = hp(i (T_q - T_q^H))            (tbk_api.cu pack_weights_host)
// straight into the stage tiles the DMMA GEMM consumes.  Gather, not scatter: every output element sums its (few)
// contributions in a fixed order, so the weights are bit-reproducible.  stage_flags_kernel then marks the (column tile,
// K-chunk) stages that hold anything non-zero; the GEMM skips the others (hk_gemm.cu, SP = true).
#include "tbk_kernels.h"
#include "tbk_math.cuh"

namespace tbk {

namespace {

constexpr int THREADS = 256;

__device__ __forceinline__ size_t wt_index(int bn, int kchunks, int q, long e) {
    const int sb = bn + 4;
    const long nt = e / bn;
    const int col = (int)(e - nt * bn);
    const int c = q / kGemmKC, kk = q - c * kGemmKC;
    return (((size_t)nt * kchunks + c) * kGemmKC + kk) * sb + col;
}

// block (a, b) element (mu, nu) of the folded hopping matrix of new lattice vector q
__device__ __forceinline__ double2 gather(const double* __restrict__ base_hop, const int2* __restrict__ table,
                                          const SupEntry* __restrict__ entries, int n, int vol, int q, int a, int b, int mu,
                                          int nu) {
    const int2 t = table[((size_t)q * vol + a) * vol + b];
    double re = 0.0, im = 0.0;
    for (int e = t.x; e < t.x + t.y; ++e) {
        const SupEntry ent = entries[e];
        const double* mat = base_hop + (size_t)ent.r * n * n * 2;
        if (ent.herm) {
            const double* z = mat + ((size_t)nu * n + mu) * 2;
            re = fma(ent.scale, z[0], re);
            im = fma(-ent.scale, z[1], im);
        } else {
            const double* z = mat + ((size_t)mu * n + nu) * 2;
            re = fma(ent.scale, z[0], re);
            im = fma(ent.scale, z[1], im);
        }
    }
    return make_double2(re, im);
}

__global__ void __launch_bounds__(THREADS)
supercell_pack_kernel(int n, int vol, int bn, int kchunks, const double* __restrict__ base_hop,
                      const int2* __restrict__ table, const SupEntry* __restrict__ entries, double* __restrict__ Wt) {
    const int N = n * vol;
    const long ntri = (long)N * (N + 1) / 2;
    const long idx = (long)blockIdx.x * THREADS + threadIdx.x;
    if (idx >= ntri) return;
    const int q = blockIdx.y;
    int I = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
    while ((long)(I + 1) * (I + 2) / 2 <= idx) ++I;
    while ((long)I * (I + 1) / 2 > idx) --I;
    const int J = (int)(idx - (long)I * (I + 1) / 2);
    const int a = I / n, mu = I - a * n, b = J / n, nu = J - b * n;
    const double2 tij = gather(base_hop, table, entries, n, vol, q, a, b, mu, nu);
    const double2 tji = gather(base_hop, table, entries, n, vol, q, b, a, nu, mu);
    // same arithmetic as pack_weights_host
    Wt[wt_index(bn, kchunks, 2 * q, idx)] = tij.x + tji.x;
    Wt[wt_index(bn, kchunks, 2 * q + 1, idx)] = -(tij.y + tji.y);
    if (J < I) {
        const long eim = ntri + (long)I * (I - 1) / 2 + J;
        Wt[wt_index(bn, kchunks, 2 * q, eim)] = tij.y - tji.y;
        Wt[wt_index(bn, kchunks, 2 * q + 1, eim)] = tij.x - tji.x;
    }
}

// flags[stage] = 1 if the stage block (kGemmKC rows x (bn + 4) doubles) holds a non-zero
__global__ void __launch_bounds__(THREADS)
stage_flags_kernel(const double* __restrict__ Wt, int stage_doubles, unsigned char* __restrict__ flags) {
    const double* blk = Wt + (size_t)blockIdx.x * stage_doubles;
    int any = 0;
    for (int i = threadIdx.x; i < stage_doubles; i += THREADS) any |= (blk[i] != 0.0) ? 1 : 0;
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) flags[blockIdx.x] = any ? 1 : 0;
}

}  // namespace

cudaError_t launch_supercell_pack(int n_base, int vol, int nq, int bn, int kchunks, const double* base_hop,
                                  const int2* table, const SupEntry* entries, double* Wt, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    const long N = (long)n_base * vol;
    const long ntri = N * (N + 1) / 2;
    const long bx = (ntri + THREADS - 1) / THREADS;
    if (bx > 2147483647L || nq > 65535) return cudaErrorInvalidConfiguration;
    supercell_pack_kernel<<<dim3((unsigned)bx, (unsigned)nq), THREADS, 0, st>>>(n_base, vol, bn, kchunks, base_hop, table,
                                                                              entries, Wt);
    return cudaGetLastError();
}

cudaError_t launch_stage_flags(const double* Wt, long n_stages, int stage_doubles, unsigned char* flags, cudaStream_t st) {
    if (n_stages <= 0) return cudaSuccess;
    if (n_stages > 2147483647L) return cudaErrorInvalidConfiguration;
    stage_flags_kernel<<<(unsigned)n_stages, THREADS, 0, st>>>(Wt, stage_doubles, flags);
    return cudaGetLastError();
}

}  // namespace tbk
