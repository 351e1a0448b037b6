// hk_gemm.cu -- H(k) build as a real GEMM on the FP64 tensor cores (DMMA, sm_100a).
//
// Replaces the Fourier-sum loop of Model.hamilton (reference src/tbmodels/_tb_model.py:1111-1123):
//     H(k) = sum_{R in half set} e^{2 pi i k.R} T_R  +  h.c.
// Because H is Hermitian the sum is rewritten with REAL coefficients,
//     H(k) = sum_r  cos(2 pi k.R_r) * (T_r + T_r^H)  +  sin(2 pi k.R_r) * i (T_r - T_r^H),
// so the lower triangle of H (N*N real numbers in the packed layout of tbk_math.cuh) is
//     Hp[k, e] = sum_q Q[k, q] * W[q, e],   Q = [cos | sin] interleaved, q = 2r / 2r+1,  W = packed weights,
// a real [n_k x 2 n_R] . [2 n_R x N^2] GEMM with half the flops of the complex product.
//
// One CTA computes a 128 (k-points) x BN (packed columns) tile:
//   * Q is never materialised in HBM: each pipeline stage's 128 x 16 slice is generated on chip with
//     sincospi(2 k.R) straight into shared memory (A operand);
//   * W was tiled at model-create time so that each stage's 16 x (BN+4) slice is one contiguous block,
//     fetched with a single TMA bulk copy (cp.async.bulk + mbarrier) into a 4-deep shared-memory ring;
//   * 8 warps (4 x 2), warp tile 32 x 8*NA, mma.sync.m8n8k4.f64 (SASS: DMMA.8x8x4), accumulators in
//     registers; padded strides (20 / BN+4 doubles) make every fragment load bank-conflict free.
// The reduction order over R is fixed by the tiling, so a k-point's result does not depend on its
// position in the batch (the reference tests compare batched and per-k calls at rtol 1e-7, atol 0).
#include "tbk_kernels.h"

namespace tbk {

namespace {

constexpr int THREADS = 256;
constexpr int BM = kGemmBM;
constexpr int KC = kGemmKC;
constexpr int SA = KC + 4;  // A row stride in doubles: (20 g + t) mod 16 distinct over a half warp
constexpr int STAGES = kGemmStages;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    uint32_t spins = 0;
    do {
        if (++spins > (1u << 26)) __trap();  // a lost TMA completion becomes an error, not a hung GPU
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}

// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int NA>
struct Cfg {
    static constexpr int BN = 16 * NA;           // 2 warps across N, NA atoms of 8 columns each
    static constexpr int SB = BN + 4;            // B row stride in doubles: (4 t + g) mod 16 distinct
    static constexpr int A_STAGE = BM * SA;      // doubles
    static constexpr int B_STAGE = KC * SB;      // doubles
    static constexpr int STAGE = A_STAGE + B_STAGE;
    static constexpr size_t smem_bytes(int dim) {
        return (size_t)STAGES * STAGE * 8 + (size_t)BM * dim * 8 + STAGES * 8;
    }
};

template <int NA>
__global__ void __launch_bounds__(THREADS, 1)
hk_gemm_kernel(const double* __restrict__ kpts, long nk, const double* __restrict__ Rd, const double* __restrict__ Wt,
               int dim, int kchunks, int n_tiles, int NN, double* __restrict__ Hp) {
    using C = Cfg<NA>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stages = reinterpret_cast<double*>(smem_raw);
    double* ks = stages + (size_t)STAGES * C::STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ks + BM * dim);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int wm = warp >> 1;  // 0..3
    const int wn = warp & 1;   // 0..1
    const int g = lane >> 2;   // 0..7
    const int t = lane & 3;    // 0..3

    const long tile = blockIdx.x;
    const long m_tile = tile / n_tiles;
    const int n_tile = (int)(tile - m_tile * n_tiles);
    const long m0 = m_tile * BM;

    // k tile -> shared (rows past the end of the batch evaluate k = 0 and are never stored)
    for (int i = tid; i < BM * dim; i += THREADS) {
        const long row = m0 + i / dim;
        ks[i] = (row < nk) ? kpts[row * dim + (i % dim)] : 0.0;
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const double* wsrc = Wt + ((size_t)n_tile * kchunks) * C::B_STAGE;

    // A-operand generator: this thread owns R vector (c*8 + rq) of the stage and rows mq, mq+32, mq+64, mq+96
    const int rq = tid & 7;
    const int mq = tid >> 3;
    auto produce = [&](int c) {
        const int s = c % STAGES;
        double* As = stages + (size_t)s * C::STAGE;
        if (tid == 0) {
            const uint32_t bar = smem_u32(&bars[s]);
            mbar_expect_tx(bar, C::B_STAGE * 8);
            tma_bulk_g2s(smem_u32(As + C::A_STAGE), wsrc + (size_t)c * C::B_STAGE, C::B_STAGE * 8, bar);
        }
        const double* rv = Rd + ((size_t)c * 8 + rq) * dim;
#pragma unroll
        for (int it = 0; it < BM / 32; ++it) {
            const int m = mq + it * 32;
            double x = 0.0;
            for (int d = 0; d < dim; ++d) x = fma(ks[m * dim + d], __ldg(rv + d), x);
            double sn, cs;
            sincospi(2.0 * x, &sn, &cs);
            *reinterpret_cast<double2*>(As + m * SA + 2 * rq) = make_double2(cs, sn);
        }
    };

    double acc[4][NA][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NA; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int c = 0; c < STAGES - 1 && c < kchunks; ++c) produce(c);

    for (int c = 0; c < kchunks; ++c) {
        const int s = c % STAGES;
        mbar_wait(smem_u32(&bars[s]), (uint32_t)((c / STAGES) & 1));
        __syncthreads();  // A(c) visible; every warp is done with the stage refilled below
        if (c + STAGES - 1 < kchunks) produce(c + STAGES - 1);

        const double* As = stages + (size_t)s * C::STAGE;
        const double* Bs = As + C::A_STAGE;
        const double* ap = As + (wm * 32 + g) * SA + t;
        const double* bp = Bs + t * C::SB + wn * (8 * NA) + g;
#pragma unroll
        for (int k4 = 0; k4 < KC / 4; ++k4) {
            double a[4], b[NA];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = ap[i * 8 * SA + k4 * 4];
#pragma unroll
            for (int j = 0; j < NA; ++j) b[j] = bp[k4 * 4 * C::SB + j * 8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < NA; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }

    // epilogue: registers -> packed H in global memory
    const int n0 = n_tile * C::BN + wn * (8 * NA) + 2 * t;
    const bool vec_ok = (NN & 1) == 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long row = m0 + wm * 32 + i * 8 + g;
        if (row >= nk) continue;
        double* out = Hp + row * (long)NN;
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            const int col = n0 + j * 8;
            if (vec_ok) {
                if (col + 1 < NN) *reinterpret_cast<double2*>(out + col) = make_double2(acc[i][j][0], acc[i][j][1]);
                else if (col < NN) out[col] = acc[i][j][0];
            } else {
                if (col < NN) out[col] = acc[i][j][0];
                if (col + 1 < NN) out[col + 1] = acc[i][j][1];
            }
        }
    }
}

template <int NA>
cudaError_t launch_na(const ModelDev& md, const double* k, long nk, double* Hp, cudaStream_t st) {
    using C = Cfg<NA>;
    const size_t smem = C::smem_bytes(md.dim);
    cudaError_t err = cudaFuncSetAttribute(hk_gemm_kernel<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    const long m_tiles = (nk + BM - 1) / BM;
    const long grid = m_tiles * md.n_tiles;
    if (grid <= 0) return cudaSuccess;
    if (grid > 2147483647L) return cudaErrorInvalidConfiguration;
    hk_gemm_kernel<NA><<<(unsigned)grid, THREADS, smem, st>>>(k, nk, md.Rd, md.Wt, md.dim, md.kchunks, md.n_tiles,
                                                              md.n * md.n, Hp);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_hk_gemm(const ModelDev& md, const double* k, long nk, double* Hp, cudaStream_t st) {
    switch (md.na) {
        case 4: return launch_na<4>(md, k, nk, Hp, st);
        case 8: return launch_na<8>(md, k, nk, Hp, st);
        case 9: return launch_na<9>(md, k, nk, Hp, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace tbk
