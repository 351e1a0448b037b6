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
// Two kernels per chunk of k-points:
//   hk_phase_kernel  writes Q = [cos | sin](2 pi k.R) with sincospi, once per (k, R) -- not once per column
//                    tile -- already cut into the 128 x 16 stage tiles the GEMM consumes (XOR-swizzled rows,
//                    so the tiles need no padding); the chunk's Q (4 KB per k-point for 251 R) stays in L2.
//   hk_gemm_kernel   one CTA computes a 128 (k-points) x BN (packed columns) tile.  Both operands of a stage
//                    (A: 16 KB of Q, B: the 16 x (BN+4) slice of W that was tiled at model-create time) are
//                    contiguous blocks fetched by TMA bulk copies (cp.async.bulk + mbarrier expect_tx) into a
//                    5-deep shared-memory ring; 8 warps (4 x 2), warp tile 32 x 8*NA, mma.sync.m8n8k4.f64
//                    (SASS: DMMA.8x8x4), accumulators in registers; swizzle / padded stride make every
//                    fragment load bank-conflict free.  The warps do nothing but LDS + DMMA.
// (Round-1 history: generating Q inside the GEMM CTA cost 9x redundant sincospi work and left the tensor pipe
//  idle during generation: 62 % DMMA utilisation, profiles/r01a_ncu_summary.txt.)
// The reduction order over R is fixed by the tiling, so a k-point's result does not depend on its
// position in the batch (the reference tests compare batched and per-k calls at rtol 1e-7, atol 0).
#include "tbk_kernels.h"
#include "tbk_math.cuh"

namespace tbk {

namespace {

constexpr int THREADS = 256;
constexpr int BM = kGemmBM;
constexpr int KC = kGemmKC;
constexpr int SA = KC;  // A row stride in doubles; column index is XOR-swizzled with (row & 3) << 2
constexpr int STAGES = kGemmStages;
constexpr int A_TILE = BM * SA;  // doubles per stage of Q

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
    static constexpr size_t smem_bytes() { return (size_t)STAGES * STAGE * 8 + STAGES * 8; }
};

// Q tiles: Qt[(m_tile * kchunks + c) * A_TILE + row * 16 + (col ^ ((row & 3) << 2))], col = 2*rq (cos), 2*rq+1 (sin)
__global__ void __launch_bounds__(THREADS)
hk_phase_kernel(const double* __restrict__ kpts, long nk, const double* __restrict__ Rd, const int* __restrict__ Pw,
                int kind, int dim, int kchunks, double* __restrict__ Qt) {
    extern __shared__ __align__(16) double ks[];  // [BM][dim]
    const int tid = threadIdx.x;
    const long m_tile = blockIdx.x;
    const long m0 = m_tile * BM;
    for (int i = tid; i < BM * dim; i += THREADS) {
        const long row = m0 + i / dim;
        ks[i] = (row < nk) ? kpts[row * dim + (i % dim)] : 0.0;  // rows past the batch: k = 0, never stored by the GEMM
    }
    __syncthreads();
    const int rq = tid & 7;
    const int mq = tid >> 3;
    if (kind == 1) {
        // k.p model: column q of Q is the monomial prod_d k_d^{p_d} of Taylor term q (kdotp.py:74); this thread
        // writes terms 2 rq and 2 rq + 1 of each 16-term chunk
        for (int c = blockIdx.y; c < kchunks; c += gridDim.y) {
            const int* pw = Pw + ((size_t)c * 16 + 2 * rq) * dim;
            double* tile = Qt + ((size_t)m_tile * kchunks + c) * A_TILE;
#pragma unroll
            for (int it = 0; it < BM / 32; ++it) {
                const int m = mq + it * 32;
                double mono[2];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    double v = 1.0;
                    for (int d = 0; d < dim; ++d) {
                        const double x = ks[m * dim + d];
                        const int e = __ldg(pw + s * dim + d);
                        for (int i = 0; i < e; ++i) v *= x;
                    }
                    mono[s] = v;
                }
                *reinterpret_cast<double2*>(tile + m * SA + ((2 * rq) ^ ((m & 3) << 2))) = make_double2(mono[0], mono[1]);
            }
        }
        return;
    }
    for (int c = blockIdx.y; c < kchunks; c += gridDim.y) {
        const double* rv = Rd + ((size_t)c * 8 + rq) * dim;
        double* tile = Qt + ((size_t)m_tile * kchunks + c) * A_TILE;
#pragma unroll
        for (int it = 0; it < BM / 32; ++it) {
            const int m = mq + it * 32;
            double x = 0.0;
            for (int d = 0; d < dim; ++d) x = fma(ks[m * dim + d], __ldg(rv + d), x);
            double sn, cs;
            sincospi_lean(2.0 * x, sn, cs);
            *reinterpret_cast<double2*>(tile + m * SA + ((2 * rq) ^ ((m & 3) << 2))) = make_double2(cs, sn);
        }
    }
}

// SP = true: block-sparse weights (supercells): kc_cnt[n_tile] stages of this column tile hold a
// non-zero block of W; their K-chunk indices are kc_idx[n_tile * kchunks + 0 .. cnt).  All other stages are skipped --
// they would add exact zeros -- so the K loop, the TMA traffic and the DMMA count shrink with the fill of the tile.
template <int NA, bool SP>
__global__ void __launch_bounds__(THREADS, 1)
hk_gemm_kernel(const double* __restrict__ Qt, long nk, const double* __restrict__ Wt, int kchunks, int n_tiles, int NN,
               double* __restrict__ Hp, const int* __restrict__ kc_cnt, const int* __restrict__ kc_idx) {
    using C = Cfg<NA>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stages = reinterpret_cast<double*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)STAGES * C::STAGE);

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

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const double* asrc = Qt + ((size_t)m_tile * kchunks) * A_TILE;
    const double* wsrc = Wt + ((size_t)n_tile * kchunks) * C::B_STAGE;
    const int* klist = SP ? kc_idx + (size_t)n_tile * kchunks : nullptr;
    const int kcount = SP ? __ldg(kc_cnt + n_tile) : kchunks;

    // one elected thread feeds the ring: two TMA bulk copies per stage, completion counted in bytes on the mbarrier
    auto produce = [&](int c) {
        if (tid == 0) {
            const int s = c % STAGES;
            double* As = stages + (size_t)s * C::STAGE;
            const uint32_t bar = smem_u32(&bars[s]);
            const int ci = SP ? __ldg(klist + c) : c;
            mbar_expect_tx(bar, (A_TILE + C::B_STAGE) * 8);
            tma_bulk_g2s(smem_u32(As), asrc + (size_t)ci * A_TILE, A_TILE * 8, bar);
            tma_bulk_g2s(smem_u32(As + C::A_STAGE), wsrc + (size_t)ci * C::B_STAGE, C::B_STAGE * 8, bar);
        }
    };

    double acc[4][NA][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NA; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int c = 0; c < STAGES - 1 && c < kcount; ++c) produce(c);

    for (int c = 0; c < kcount; ++c) {
        const int s = c % STAGES;
        mbar_wait(smem_u32(&bars[s]), (uint32_t)((c / STAGES) & 1));
        __syncthreads();  // every warp is done with the stage refilled below
        if (c + STAGES - 1 < kcount) produce(c + STAGES - 1);

        const double* As = stages + (size_t)s * C::STAGE;
        const double* Bs = As + C::A_STAGE;
        const double* ap = As + (wm * 32 + g) * SA;
        const int sw = (g & 3) << 2;  // rows wm*32 + i*8 + g: (row & 3) == (g & 3)
        const double* bp = Bs + t * C::SB + wn * (8 * NA) + g;
#pragma unroll
        for (int k4 = 0; k4 < KC / 4; ++k4) {
            double a[4], b[NA];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = ap[i * 8 * SA + ((k4 * 4 + t) ^ sw)];
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
cudaError_t launch_na(const ModelDev& md, long nk, const double* Qt, double* Hp, cudaStream_t st) {
    using C = Cfg<NA>;
    const long m_tiles = (nk + BM - 1) / BM;
    if (m_tiles <= 0) return cudaSuccess;
    const size_t smem = C::smem_bytes();
    const long grid = m_tiles * md.n_tiles;
    if (grid > 2147483647L) return cudaErrorInvalidConfiguration;
    if (md.kc_cnt) {
        cudaError_t err = cudaFuncSetAttribute(hk_gemm_kernel<NA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        hk_gemm_kernel<NA, true><<<(unsigned)grid, THREADS, smem, st>>>(Qt, nk, md.Wt, md.kchunks, md.n_tiles, md.n * md.n, Hp,
                                                                      md.kc_cnt, md.kc_idx);
        return cudaGetLastError();
    }
    cudaError_t err = cudaFuncSetAttribute(hk_gemm_kernel<NA, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    hk_gemm_kernel<NA, false><<<(unsigned)grid, THREADS, smem, st>>>(Qt, nk, md.Wt, md.kchunks, md.n_tiles, md.n * md.n, Hp,
                                                                   nullptr, nullptr);
    return cudaGetLastError();
}

}  // namespace

size_t hk_gemm_smem_bytes(const ModelDev& md) {
    switch (md.na) {
        case 4: return Cfg<4>::smem_bytes();
        case 8: return Cfg<8>::smem_bytes();
        case 9: return Cfg<9>::smem_bytes();
        default: return 0;
    }
}

size_t hk_gemm_q_doubles(const ModelDev& md, long nk) {
    return (size_t)((nk + BM - 1) / BM) * (size_t)md.kchunks * A_TILE;
}

cudaError_t launch_hk_phase(const ModelDev& md, const double* k, long nk, double* Qt, cudaStream_t st) {
    const long m_tiles = (nk + BM - 1) / BM;
    if (m_tiles <= 0 || md.kchunks <= 0) return cudaSuccess;
    if (m_tiles > 2147483647L) return cudaErrorInvalidConfiguration;
    int ysplit = 1;
    while (m_tiles * ysplit < 592 && ysplit * 2 <= md.kchunks) ysplit *= 2;  // >= 4 CTAs per SM worth of work
    hk_phase_kernel<<<dim3((unsigned)m_tiles, (unsigned)ysplit), THREADS, (size_t)BM * md.dim * 8, st>>>(
        k, nk, md.Rd, md.Pw, md.kind, md.dim, md.kchunks, Qt);
    return cudaGetLastError();
}

cudaError_t launch_hk_gemm(const ModelDev& md, long nk, const double* Qt, double* Hp, cudaStream_t st) {
    switch (md.na) {
        case 4: return launch_na<4>(md, nk, Qt, Hp, st);
        case 8: return launch_na<8>(md, nk, Qt, Hp, st);
        case 9: return launch_na<9>(md, nk, Qt, Hp, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace tbk
