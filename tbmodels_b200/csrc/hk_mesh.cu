// hk_mesh.cu -- H(k) on a regular k-mesh: the Fourier sum factorised over the last mesh dimension.
//
// Row f4 of SURVEY.md section 8 ("mesh-native k generation"): band structures are evaluated on tensor-product meshes
// (BASELINE.json C1 / C3 / C5 are k-grids), and on a mesh the sum of Model.hamilton (reference
// src/tbmodels/_tb_model.py:1111-1123) separates.  Write k = (kappa, k_z) with kappa the leading D-1 coordinates and
// sort the stored lattice vectors into classes c by their last component z_c.  For one mesh LINE (fixed kappa)
//     G_c(kappa) = sum_{r in c} e^{2 pi i kappa.R_r} T_r ,      H(kappa, k_z) = sum_c e^{2 pi i k_z z_c} G_c + h.c.
// and in the packed Hermitian ("hp") form of tbk_math.cuh, with phi_r = 2 pi kappa.R_r and W the Hermitian-split weights
//     A_c = hp(G_c + G_c^H)    = sum_{r in c}  cos(phi_r) W[2r] + sin(phi_r) W[2r+1]
//     B_c = hp(i(G_c - G_c^H)) = sum_{r in c} -sin(phi_r) W[2r] + cos(phi_r) W[2r+1]
//     Hp(kappa, k_z) = sum_c cos(2 pi k_z z_c) A_c + sin(2 pi k_z z_c) B_c .
// Stage A (mesh_phase_kernel + the unchanged DMMA GEMM of hk_gemm.cu) evaluates the 2C rows (A_c, B_c) of every line as
// 2C pseudo k-points; stage B (mesh_lines_kernel) expands each line along k_z with K = 2C instead of K = 2 n_R:
// per k-point 2C N^2 FMAs instead of 2 n_R N^2 (C3: 18 vs 502), the same terms as the reference's sum, reassociated.
// mesh_kpoints_kernel writes the explicit k-points of a mesh range for the models that take the ordinary path.
#include "tbk_kernels.h"
#include "tbk_math.cuh"

namespace tbk {

namespace {

constexpr int THREADS = 256;
constexpr int BM = kGemmBM;
constexpr int KC = kGemmKC;
constexpr int A_TILE = BM * KC;

struct MeshDims {
    long n[kMaxDim];      // mesh points per dimension
    double shift[kMaxDim];  // k_d = (i_d + shift_d) / n_d
    int dim;
};

// Q tiles of the pseudo k-points (same swizzled 128 x 16 stage tiles as hk_phase_kernel): row = line * K2 + j,
// j = 2 c (A_c) or 2 c + 1 (B_c); entry pair (2 rq, 2 rq + 1) of R vector rq is (cos, sin) / (-sin, cos) of
// 2 pi kappa.R if the vector belongs to class c, else zero.
__global__ void __launch_bounds__(THREADS)
mesh_phase_kernel(MeshDims md, long line0, long rows, int K2, const double* __restrict__ Rd,
                  const int* __restrict__ Rc, int kchunks, double* __restrict__ Qt) {
    const int tid = threadIdx.x;
    const long m_tile = blockIdx.x;
    const int rq = tid & 7;
    const int mq = tid >> 3;
    const int pd = md.dim - 1;  // prefix dimensions
    double kap[BM / 32][kMaxDim];
    int cls[BM / 32], part[BM / 32];
#pragma unroll
    for (int it = 0; it < BM / 32; ++it) {
        const long row = m_tile * BM + mq + it * 32;
        long line = line0 + ((row < rows) ? row : 0) / K2;
        const int j = (int)(((row < rows) ? row : 0) % K2);
        cls[it] = (row < rows) ? (j >> 1) : -2;  // rows past the batch: all zero, never stored by the GEMM
        part[it] = j & 1;
#pragma unroll
        for (int d = kMaxDim - 2; d >= 0; --d) {  // C order: the last prefix dimension runs fastest
            kap[it][d] = 0.0;
            if (d < pd) {
                const long i = line % md.n[d];
                line /= md.n[d];
                kap[it][d] = ((double)i + md.shift[d]) / (double)md.n[d];
            }
        }
    }
    for (int c = blockIdx.y; c < kchunks; c += gridDim.y) {
        const int r = c * 8 + rq;
        const double* rv = Rd + (size_t)r * md.dim;
        const int rc = __ldg(Rc + r);
        double* tile = Qt + ((size_t)m_tile * kchunks + c) * A_TILE;
#pragma unroll
        for (int it = 0; it < BM / 32; ++it) {
            const int m = mq + it * 32;
            double2 q = make_double2(0.0, 0.0);
            if (rc == cls[it]) {
                double x = 0.0;
#pragma unroll
                for (int d = 0; d < kMaxDim - 1; ++d)
                    if (d < pd) x = fma(kap[it][d], __ldg(rv + d), x);
                double sn, cs;
                sincospi_lean(2.0 * x, sn, cs);
                q = part[it] ? make_double2(-sn, cs) : make_double2(cs, sn);
            }
            *reinterpret_cast<double2*>(tile + m * KC + ((2 * rq) ^ ((m & 3) << 2))) = q;
        }
    }
}

// Qz[i][2c] = cos(2 pi k_z z_c), Qz[i][2c+1] = sin(2 pi k_z z_c), k_z = (i + shift) / n_z.
__global__ void mesh_qz_kernel(long nz, double shift, int C, const double* __restrict__ zc, double* __restrict__ Qz) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nz * C) return;
    const long i = idx / C;
    const int c = (int)(idx - i * C);
    const double kz = ((double)i + shift) / (double)nz;
    double sn, cs;
    sincospi_lean(2.0 * kz * zc[c], sn, cs);
    *reinterpret_cast<double2*>(Qz + (i * C + c) * 2) = make_double2(cs, sn);
}

// Stage B: Hp[(line, i)][e] = sum_j Qz[i][j] AB[line][j][e].  One CTA per (line, 128-column tile): the AB tile stays in
// shared memory while the CTA walks the line in 64-point steps; 256 threads, 4 x 8 outputs each.
constexpr int TC = 128;  // columns per CTA
constexpr int TR = 64;   // mesh points per step

__global__ void __launch_bounds__(THREADS)
mesh_lines_kernel(const double* __restrict__ AB, const double* __restrict__ Qz, long nz, int K2, int NN, int col_tiles,
                  double* __restrict__ Hp) {
    extern __shared__ __align__(16) double sm[];
    double* ABs = sm;                       // [K2][TC]
    double* Qs = sm + (size_t)K2 * TC;      // [TR][K2 + 1]
    const int tid = threadIdx.x;
    const long line = blockIdx.x / col_tiles;
    const int ct = (int)(blockIdx.x - line * col_tiles);
    const int col0 = ct * TC;
    const int ncol = (NN - col0 < TC) ? (NN - col0) : TC;
    const double* ab = AB + (size_t)line * K2 * NN + col0;
    for (int idx = tid; idx < K2 * TC; idx += THREADS) {
        const int j = idx / TC, cc = idx - j * TC;
        ABs[idx] = (cc < ncol) ? ab[(size_t)j * NN + cc] : 0.0;
    }
    // thread (tx, ty): mesh points ty * 4 .. + 4 of the step, columns 32 b + 2 tx + {0, 1}, b = 0 .. 3 -- the 16 lanes of a
    // half warp read / write 256 contiguous bytes per instruction
    const int tx = tid & 15, ty = tid >> 4;
    const int lane = tid & 31, warp = tid >> 5;
    const int QS = K2 + 1;
    const bool vec_ok = (NN & 1) == 0;
    for (long i0 = 0; i0 < nz; i0 += TR) {
        __syncthreads();  // ABs loaded / previous step's Qs consumed
        for (int rr = warp; rr < TR; rr += THREADS / 32)
            for (int j = lane; j < K2; j += 32) Qs[rr * QS + j] = (i0 + rr < nz) ? Qz[(i0 + rr) * K2 + j] : 0.0;
        __syncthreads();
        double acc[4][8];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;
        const double* qp = Qs + (ty * 4) * QS;
        const double* wp = ABs + 2 * tx;
#pragma unroll 2
        for (int j = 0; j < K2; ++j) {
            double q[4], w[8];
#pragma unroll
            for (int a = 0; a < 4; ++a) q[a] = qp[a * QS + j];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const double2 t2 = *reinterpret_cast<const double2*>(wp + j * TC + 32 * b);
                w[2 * b] = t2.x;
                w[2 * b + 1] = t2.y;
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] = fma(q[a], w[b], acc[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const long i = i0 + ty * 4 + a;
            if (i >= nz) continue;
            double* out = Hp + ((size_t)line * nz + i) * NN + col0 + 2 * tx;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int cc = 32 * b + 2 * tx;
                if (vec_ok && cc + 1 < ncol) {
                    *reinterpret_cast<double2*>(out + 32 * b) = make_double2(acc[a][2 * b], acc[a][2 * b + 1]);
                } else {
                    if (cc < ncol) out[32 * b] = acc[a][2 * b];
                    if (cc + 1 < ncol) out[32 * b + 1] = acc[a][2 * b + 1];
                }
            }
        }
    }
}

__global__ void mesh_kpoints_kernel(MeshDims md, long first, long count, double* __restrict__ k) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    long f = first + idx;
    for (int d = md.dim - 1; d >= 0; --d) {
        const long i = f % md.n[d];
        f /= md.n[d];
        k[idx * md.dim + d] = ((double)i + md.shift[d]) / (double)md.n[d];
    }
}

MeshDims make_dims(int dim, const int64_t* dims, const double* shift) {
    MeshDims md;
    md.dim = dim;
    for (int d = 0; d < kMaxDim; ++d) {
        md.n[d] = d < dim ? (long)dims[d] : 1;
        md.shift[d] = (d < dim && shift) ? shift[d] : 0.0;
    }
    return md;
}

}  // namespace

size_t mesh_lines_smem_bytes(int K2) { return ((size_t)K2 * TC + (size_t)TR * (K2 + 1)) * 8; }

cudaError_t launch_mesh_phase(const ModelDev& m, const int64_t* dims, const double* shift, long line0, long n_lines,
                              double* Qt, cudaStream_t st) {
    const int K2 = 2 * m.nclass;
    const long rows = n_lines * K2;
    const long m_tiles = (rows + BM - 1) / BM;
    if (m_tiles <= 0 || m.kchunks <= 0) return cudaSuccess;
    if (m_tiles > 2147483647L) return cudaErrorInvalidConfiguration;
    int ysplit = 1;
    while (m_tiles * ysplit < 592 && ysplit * 2 <= m.kchunks) ysplit *= 2;
    mesh_phase_kernel<<<dim3((unsigned)m_tiles, (unsigned)ysplit), THREADS, 0, st>>>(
        make_dims(m.dim, dims, shift), line0, rows, K2, m.Rd, m.Rc, m.kchunks, Qt);
    return cudaGetLastError();
}

cudaError_t launch_mesh_qz(const ModelDev& m, long nz, double shift, double* Qz, cudaStream_t st) {
    const long total = nz * m.nclass;
    if (total <= 0) return cudaSuccess;
    mesh_qz_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(nz, shift, m.nclass, m.zc, Qz);
    return cudaGetLastError();
}

cudaError_t launch_mesh_lines(const ModelDev& m, const double* AB, const double* Qz, long nz, long n_lines, double* Hp,
                              cudaStream_t st) {
    const int K2 = 2 * m.nclass;
    const int NN = m.n * m.n;
    const int col_tiles = (NN + TC - 1) / TC;
    const long grid = n_lines * col_tiles;
    if (grid <= 0) return cudaSuccess;
    if (grid > 2147483647L) return cudaErrorInvalidConfiguration;
    const size_t smem = mesh_lines_smem_bytes(K2);
    cudaError_t err = cudaFuncSetAttribute(mesh_lines_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    mesh_lines_kernel<<<(unsigned)grid, THREADS, smem, st>>>(AB, Qz, nz, K2, NN, col_tiles, Hp);
    return cudaGetLastError();
}

cudaError_t launch_mesh_kpoints(int dim, const int64_t* dims, const double* shift, long first, long count, double* k,
                                cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    mesh_kpoints_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(make_dims(dim, dims, shift), first, count, k);
    return cudaGetLastError();
}

}  // namespace tbk
