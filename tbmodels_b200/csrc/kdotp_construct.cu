// kdotp_construct.cu -- Taylor coefficients of Model.construct_kdotp on the device.
//
// Reference src/tbmodels/_tb_model.py:942-982: for every power tuple p with |p| <= order
//     C_p = (2 pi i)^|p| / prod_d p_d!  *  sum_{R in half set} [ R^p e^{2 pi i k.R} T_R  +  (-R)^p e^{-2 pi i k.R} T_R^H ]
// with R^p = prod_d R_d^{p_d} (0^0 = 1).  Since (-R)^p = (-1)^|p| R^p the bracket is S_p + (-1)^|p| S_p^H,
// S_p = sum_R R^p e^{2 pi i k.R} T_R -- "the same Fourier sum with R^p weights" -- and with the Hermitian-split weights
// W[2r] = hp(T_r + T_r^H), W[2r+1] = hp(i (T_r - T_r^H)) of the H(k) build (c_r, s_r = cos, sin 2 pi k.R_r):
//     |p| even:  S + S^H = sum_r R^p ( c_r W[2r] + s_r W[2r+1])            C_p = (-1)^(|p|/2)     (2 pi)^|p| / p! * (S + S^H)
//     |p| odd :  S - S^H = i sum_r R^p ( s_r W[2r] - c_r W[2r+1])          C_p = (-1)^((|p|+1)/2) (2 pi)^|p| / p! * (...)
// so every coefficient is a REAL multiple of a Hermitian matrix, accumulated in the packed lower triangle and mirrored
// on output (exactly Hermitian, which the KdotpModel constructor checks, kdotp.py:40-44).
// One CTA per (block of 256 lower-triangle elements, Taylor term, expansion point); the 2 n_R real coefficients of the
// (point, term) pair are computed once per CTA into shared memory.  The weights are read from whichever layout the handle
// holds: the plain rows of the fused small-N path or the stage tiles of the GEMM path.
#include <algorithm>

#include "tbk_kernels.h"
#include "tbk_math.cuh"

namespace tbk {

namespace {

constexpr int THREADS = 256;

__device__ __forceinline__ double w_at(const ModelDev& md, long NN, int q, long e) {
    if (md.W) return md.W[(size_t)q * NN + e];
    const int bn = 16 * md.na, sb = bn + 4;
    const long nt = e / bn;
    const int col = (int)(e - nt * bn);
    const int c = q / kGemmKC, kk = q - c * kGemmKC;
    return md.Wt[((size_t)nt * md.kchunks + c) * ((size_t)kGemmKC * sb) + (size_t)kk * sb + col];
}

__global__ void __launch_bounds__(THREADS)
kdotp_coeff_kernel(ModelDev md, const double* __restrict__ kpts, const int* __restrict__ powers,
                   const double* __restrict__ fac, int n_terms, double* __restrict__ out) {
    extern __shared__ __align__(16) double coef[];  // [2 nR]: (a_r, b_r)
    const int n = md.n, dim = md.dim, nR = md.nR;
    const long NN = (long)n * n;
    const int term = blockIdx.y;
    const long kp = blockIdx.z;
    const int* pw = powers + (size_t)term * dim;
    int order = 0;
    for (int d = 0; d < dim; ++d) order += pw[d];
    const double f = fac[term];
    for (int r = threadIdx.x; r < nR; r += THREADS) {
        double x = 0.0, rp = 1.0;
        for (int d = 0; d < dim; ++d) {
            const double Rv = md.Rd[(size_t)r * dim + d];
            x = fma(kpts[kp * dim + d], Rv, x);
            for (int q = 0; q < pw[d]; ++q) rp *= Rv;  // integer powers, exact in double; 0^0 = 1 like numpy
        }
        double sn, cs;
        sincospi_lean(2.0 * x, sn, cs);
        const double w = f * rp;
        coef[2 * r] = (order & 1) ? w * sn : w * cs;
        coef[2 * r + 1] = (order & 1) ? -w * cs : w * sn;
    }
    __syncthreads();
    const long ntri = (long)n * (n + 1) / 2;
    const long idx = (long)blockIdx.x * THREADS + threadIdx.x;
    if (idx >= ntri) return;
    // idx -> (i, j <= i) of the packed lower triangle
    int i = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
    while ((long)(i + 1) * (i + 2) / 2 <= idx) ++i;
    while ((long)i * (i + 1) / 2 > idx) --i;
    const int j = (int)(idx - (long)i * (i + 1) / 2);
    const long ire = idx;
    const long iim = ntri + (long)i * (i - 1) / 2 + j;  // only used for j < i
    double re = 0.0, im = 0.0;
    for (int q = 0; q < 2 * nR; ++q) {
        const double c = coef[q];
        re = fma(c, w_at(md, NN, q, ire), re);
        if (j < i) im = fma(c, w_at(md, NN, q, iim), im);
    }
    double* o = out + ((size_t)kp * n_terms + term) * (size_t)NN * 2;
    o[((size_t)i * n + j) * 2] = re;
    o[((size_t)i * n + j) * 2 + 1] = im;
    if (j < i) {
        o[((size_t)j * n + i) * 2] = re;
        o[((size_t)j * n + i) * 2 + 1] = -im;
    }
}

}  // namespace

cudaError_t launch_kdotp_coeff(const ModelDev& md, const double* k, long nk, const int* powers, const double* fac,
                               int n_terms, double* out, cudaStream_t st) {
    if (nk <= 0 || n_terms <= 0) return cudaSuccess;
    if (nk > 65535 || n_terms > 65535) return cudaErrorInvalidConfiguration;
    const long ntri = (long)md.n * (md.n + 1) / 2;
    const size_t smem = (size_t)std::max(2 * md.nR, 2) * 8;
    cudaError_t err = cudaFuncSetAttribute(kdotp_coeff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    dim3 grid((unsigned)((ntri + THREADS - 1) / THREADS), (unsigned)n_terms, (unsigned)nk);
    kdotp_coeff_kernel<<<grid, THREADS, smem, st>>>(md, k, powers, fac, n_terms, out);
    return cudaGetLastError();
}

}  // namespace tbk
