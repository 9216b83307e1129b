// k2_rotate.cu -- eigen-rotation  rot[r,k] = sum_j g[r,j] * U^T[k,j]  as an FP64 tensor-core GEMM.
//
// Replaces, for the B200 path:
//   rotate_snp_block_with_ut_blas (cblas_sgemm, RowMajor NoTrans x Trans)   src/stats/lmm.rs:728-783
//   lmm_rotate_x_y_with_ut_f64                                               src/stats/reml.rs:109-198
//
// Arithmetic contract (DESIGN.md "Rotation arithmetic"): inputs are f32-valued (genotypes centred in
// f32, U^T rounded to f32 by the caller exactly like python/janusx/pyBLUP/assoc.py:1818), products and
// sums are FP64 (every product of two f32 values is exact in f64), the result is rounded once to f32 --
// the reference's storage type for the rotated block.
//
// Kernel: persistent CTAs (one per SM), 8 consumer warps + 1 TMA producer warp.  CTA tile 128 (SNP
// rows) x 128 (eigen-directions), k-slab = 16 doubles = one 128-byte SWIZZLE_128B row, 6-stage
// mbarrier pipeline (192 KB smem).  On sm_100a every f64 mma.sync shape lowers to DMMA.8x8x4, so the
// kernel issues m8n8k4 directly: warp tile 64x32 = 8x4 DMMA tiles, 64 f64 accumulators per lane.
// The four k-indices a lane feeds to one DMMA are chosen as k = 8*(t>>1) + 2*j + (t&1) (t = lane&3,
// j = sub-step) so that with the 128-byte swizzle every 64-bit fragment load of a half-warp hits 16
// distinct 8-byte banks: conflict-free LDS.64 for both operands straight from the TMA layout.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "jxb_common.cuh"

namespace jxb {

namespace {

constexpr int BM = kRotBM, BN = kRotBN, BK = kRotBK, STAGES = kRotStages;
constexpr int CONSUMER_WARPS = 8;
constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
constexpr int A_BYTES = BM * BK * 8;
constexpr int B_BYTES = BN * BK * 8;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int GROUP_M = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void tile_coords(int tile, int mt_count, int nt_count, int& mt, int& nt) {
    const int per_group = GROUP_M * nt_count;
    const int group = tile / per_group;
    const int first_m = group * GROUP_M;
    const int gsize = min(GROUP_M, mt_count - first_m);
    const int in_group = tile - group * per_group;
    mt = first_m + in_group % gsize;
    nt = in_group / gsize;
}

__global__ void __launch_bounds__(THREADS, 1)
rotate_dmma_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_ut,
                   float* __restrict__ rot, size_t ldc, int transposed, int n, int kblocks, int max_rows,
                   const int32_t* __restrict__ n_rows_dev) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;            // SWIZZLE_128B wants 1024-byte alignment
    const uint32_t bars = tiles + STAGES * STAGE_BYTES;       // full[STAGES], empty[STAGES]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int rows = n_rows_dev ? min(*n_rows_dev, max_rows) : max_rows;
    const int mt_count = (rows + BM - 1) / BM;
    const int nt_count = (n + BN - 1) / BN;
    const int n_tiles = mt_count * nt_count;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bars + 8 * s, 1);                          // producer arrive + tx bytes
            mbar_init(bars + 8 * (STAGES + s), CONSUMER_WARPS);  // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == CONSUMER_WARPS) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                int mt, nt;
                tile_coords(tile, mt_count, nt_count, mt, nt);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(bars + 8 * (STAGES + stage), phase ^ 1u);
                    const uint32_t full = bars + 8 * stage;
                    mbar_expect_tx(full, STAGE_BYTES);
                    const uint32_t sa = tiles + stage * STAGE_BYTES;
                    tma_load_2d(sa, &tm_g, kb * BK, mt * BM, full);
                    tma_load_2d(sa + A_BYTES, &tm_ut, kb * BK, nt * BN, full);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
        return;
    }

    // ===== consumers: 2 (M) x 4 (N) warps, warp tile 64 x 32 =====
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, t = lane & 3;
    // byte offset of this lane's element inside a 128-byte row, per k4 sub-step j (see header comment)
    uint32_t koff[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) koff[j] = (uint32_t)((((4 * (t >> 1) + j) ^ g) << 4) + ((t & 1) << 3));
    const uint32_t a_row = (uint32_t)((wm * 64 + g) * 128);
    const uint32_t b_row = (uint32_t)(A_BYTES + (wn * 32 + g) * 128);

    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int mt, nt;
        tile_coords(tile, mt_count, nt_count, mt, nt);
        double acc[8][4][2];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

        for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(bars + 8 * stage, phase);
            const uint32_t sa = tiles + stage * STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double af[8], bf[4];
#pragma unroll
                for (int mi = 0; mi < 8; ++mi) af[mi] = lds_f64(sa + a_row + mi * 1024 + koff[j]);
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) bf[ni] = lds_f64(sa + b_row + ni * 1024 + koff[j]);
#pragma unroll
                for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * (STAGES + stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }

        // epilogue: round once to f32 (the reference's rotated-block storage type) and store, either
        // row-major [row][col] or SNP-minor [col][row] (what the per-SNP solve kernel reads coalesced)
        const int row0 = mt * BM + wm * 64 + g;
        const int col0 = nt * BN + wn * 32 + 2 * t;
        if (transposed) {
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) {
                const int row = row0 + mi * 8;
                if (row >= rows) continue;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const int col = col0 + ni * 8;
                    if (col < n) rot[(size_t)col * ldc + row] = (float)acc[mi][ni][0];
                    if (col + 1 < n) rot[(size_t)(col + 1) * ldc + row] = (float)acc[mi][ni][1];
                }
            }
        } else {
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) {
                const int row = row0 + mi * 8;
                if (row >= rows) continue;
                float* dst = rot + (size_t)row * ldc;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const int col = col0 + ni * 8;
                    if (col + 1 < n) {
                        *reinterpret_cast<float2*>(dst + col) =
                            make_float2((float)acc[mi][ni][0], (float)acc[mi][ni][1]);
                    } else if (col < n) {
                        dst[col] = (float)acc[mi][ni][0];
                    }
                }
            }
        }
    }
}

// Plain CUDA-core FP64 tile kernel (no TMA, no DMMA): the in-tree cross-check for the DMMA kernel and
// the variant used when tensor maps cannot be built.  64x64 tile, 16x16 threads, 4x4 outputs each.
__global__ void __launch_bounds__(256) rotate_simple_kernel(const double* __restrict__ g64, size_t ldk,
                                                            const double* __restrict__ ut, float* __restrict__ rot,
                                                            size_t ldc, int transposed, int n, int max_rows,
                                                            const int32_t* __restrict__ n_rows_dev) {
    __shared__ double sA[64][17];
    __shared__ double sB[64][17];
    const int rows = n_rows_dev ? min(*n_rows_dev, max_rows) : max_rows;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    if (m0 >= rows) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    const int kpad = (int)ldk;
    for (int k0 = 0; k0 < kpad; k0 += 16) {
        for (int e = threadIdx.x; e < 64 * 16; e += 256) {
            const int r = e >> 4, c = e & 15;
            const int gr = m0 + r, gc = n0 + r;
            sA[r][c] = (gr < rows) ? g64[(size_t)gr * ldk + k0 + c] : 0.0;
            sB[r][c] = (gc < n) ? ut[(size_t)gc * ldk + k0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[ty * 4 + i][kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[tx * 4 + j][kk];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = m0 + ty * 4 + i;
        if (row >= rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + tx * 4 + j;
            if (col < n) rot[transposed ? (size_t)col * ldc + row : (size_t)row * ldc + col] = (float)acc[i][j];
        }
    }
}

// A6: X_rot[i,c] = sum_j U^T[i,j] x[j,c], y_rot[i] = sum_j U^T[i,j] y[j]  (f64 accumulate; U^T holds
// f32 values widened exactly).  One warp per eigen-direction i.
__global__ void __launch_bounds__(256) rotate_xy_kernel(const double* __restrict__ ut, size_t ldk, int n,
                                                        const double* __restrict__ x, int q,
                                                        const double* __restrict__ y, double* __restrict__ x_rot,
                                                        double* __restrict__ y_rot) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < n; i += nwarps) {
        const double* u = ut + (size_t)i * ldk;
        for (int c = 0; c <= q; ++c) {
            double acc = 0.0;
            if (c < q) {
                for (int j = lane; j < n; j += 32) acc = fma(u[j], x[(size_t)j * q + c], acc);
            } else {
                for (int j = lane; j < n; j += 32) acc = fma(u[j], y[j], acc);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) {
                if (c < q) x_rot[(size_t)i * q + c] = acc; else y_rot[i] = acc;
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int encode_f64_rows(CUtensorMap* tm, void* base, size_t rows, size_t ldk, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(-101, "cuTensorMapEncodeTiled is unavailable from the CUDA driver");
    cuuint64_t dims[2] = {(cuuint64_t)ldk, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ldk * 8};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-102, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return 0;
}

}  // namespace

int make_tensor_maps(Model& m) {
    if (!m.ut || !m.g64) return 0;
    if (!m.tmap_ut) m.tmap_ut = aligned_alloc(64, sizeof(CUtensorMap));
    if (!m.tmap_g) m.tmap_g = aligned_alloc(64, sizeof(CUtensorMap));
    int rc = encode_f64_rows((CUtensorMap*)m.tmap_ut, m.ut, m.n_pad, m.ldk, BN);
    if (rc) return rc;
    return encode_f64_rows((CUtensorMap*)m.tmap_g, m.g64, round_up(m.cap_rows, BM), m.ldk, BM);
}

int launch_rotate(Model& m, size_t max_rows, const int32_t* n_rows_dev, float* out, size_t ld, int transposed,
                  cudaStream_t st, int variant) {
    if (max_rows == 0) return 0;
    if (!m.ut) return fail(-3, "model was created without U^T; rotation is unavailable");
    if (variant == 1) {
        dim3 grid((unsigned)((m.n + 63) / 64), (unsigned)((max_rows + 63) / 64));
        rotate_simple_kernel<<<grid, 256, 0, st>>>(m.g64, m.ldk, m.ut, out, ld, transposed, (int)m.n, (int)max_rows,
                                                   n_rows_dev);
        JXB_CUDA_OK(cudaGetLastError());
        return 0;
    }
    if (!m.tmap_ut || !m.tmap_g) {
        int rc = make_tensor_maps(m);
        if (rc) return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
        JXB_CUDA_OK(cudaFuncSetAttribute(rotate_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m.device);
    const size_t tiles = ((max_rows + BM - 1) / BM) * ((m.n + BN - 1) / BN);
    const int grid = (int)std::min<size_t>((size_t)sms, tiles);
    const int kblocks = (int)(m.ldk / BK);
    rotate_dmma_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(*(const CUtensorMap*)m.tmap_g, *(const CUtensorMap*)m.tmap_ut,
                                                          out, ld, transposed, (int)m.n, kblocks, (int)max_rows,
                                                          n_rows_dev);
    JXB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_rotate_xy(const Model& m, const float* ut_f32, const double* x, size_t q, const double* y,
                     double* x_rot, double* y_rot, cudaStream_t st) {
    (void)ut_f32;
    if (!m.ut) return fail(-3, "model was created without U^T; rotation is unavailable");
    const int blocks = (int)std::min<size_t>((m.n + 7) / 8, 148 * 4);
    rotate_xy_kernel<<<blocks, 256, 0, st>>>(m.ut, m.ldk, (int)m.n, x, (int)q, y, x_rot, y_rot);
    JXB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace jxb
