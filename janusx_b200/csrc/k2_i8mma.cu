// k2_i8mma.cu -- hand-written tcgen05 (sm_100a) kernel for the int8-sliced exact rotation (see k2_int8.cu for
// the arithmetic): all digit planes of a U^T column block are stacked into ONE MMA N dimension so a 128-row
// genotype tile meets every slice in a single tcgen05.mma, the int32 accumulators live in TMEM, and the
// epilogue recombines the slices by Horner in f64 straight out of TMEM -- the int32 slice results never
// touch HBM.
//
// Tiling (pass D, 7 slices): CTA tile = 256 genotype rows (two M=128 sub-tiles) x 32 eigen-directions;
//   B tile  = [7 slices][32 rows][128 B of K]  -> MMA N = 224, one TMA box from the 3-D plane tensor
//   A tiles = 2 x [128 rows][128 B of K]       -> two TMA boxes
//   TMEM    = 2 accumulators x 224 columns (int32) of the 512-column allocation
//   k-slab  = 128 bytes (SWIZZLE_128B), 4 tcgen05.mma.kind::i8 (K = 32) per sub-tile per slab, 3 stages (180 KB)
// Pass 2 (top 3 slices, hom indicator) uses 64 eigen-directions per tile (N = 192); pass M (missing indicator)
// is pass D's shape.  Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld 32x32b, one TMEM lane quarter each).  Persistent CTAs, grouped raster.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "jxb_common.cuh"
#include "tc05.cuh"

namespace jxb {

namespace {

using namespace tc05;

constexpr int TM = 128;
constexpr int MSUB = 2;
constexpr int A_STAGE = MSUB * TM * KSLAB;       // 32 KB
constexpr int NTHREADS = 192;
// Two builds of the kernel: the stand-alone one (3 stages = 180 KB of shared memory, registers unconstrained) and the
// SLIM one that shares an SM with three CTAs of the streamed FP64 solve (cabi.cu scan_streamed): 2 stages (121 KB) and
// at most 64 registers per thread (192 x 64 = 12 K of the SM's 64 K).
constexpr int kStagesFull = 3, kStagesSlim = 2;
constexpr int GROUP_M = 16;
constexpr int ACC_COLS = 256;                    // TMEM column stride between the two accumulators


__device__ __forceinline__ void tile_coords(int tile, int mt_count, int nt_count, int& mt, int& nt) {
    const int per_group = GROUP_M * nt_count;
    const int group = tile / per_group;
    const int first_m = group * GROUP_M;
    const int gsize = min(GROUP_M, mt_count - first_m);
    const int in_group = tile - group * per_group;
    mt = first_m + in_group % gsize;
    nt = in_group / gsize;
}

// mode 0: corr  = coef[2] * T * is        (top slices of the hom indicator; T carries the 256^slice0 factor)
// mode 1: corr += coef[3] * T * is        (missing indicator)
// mode 2: rot   = f32(coef[0] * rk + coef[1] * T * is + corr)
// rows [row0, rows) of the operand planes are rotated (row0 is a multiple of the 256-row CTA tile); `corr` holds the
// correction terms of these rows only, indexed r - row0.
template <int NSL, int CG, bool SLIM>
__global__ void __maxnreg__(SLIM ? 48 : 232)
i8_rotate_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int a_plane,
                 int slice0, int mode, int row0, int rows, int n, int kslabs, const double* __restrict__ coef,
                 const double* __restrict__ inv_scale, const double* __restrict__ rk, double* __restrict__ corr,
                 size_t ld_corr, float* __restrict__ rot, size_t ldc, int transposed) {
    constexpr int STAGES = SLIM ? kStagesSlim : kStagesFull;
    constexpr int NB = NSL * CG;                    // MMA N
    constexpr int B_STAGE = NB * KSLAB;
    constexpr int STAGE = A_STAGE + B_STAGE;
    static_assert(NB % 16 == 0 && NB <= 256, "invalid MMA N");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles_s = (raw + 1023u) & ~1023u;
    const uint32_t bars = tiles_s + STAGES * STAGE;     // full[S], empty[S], tmem_full, tmem_empty
    const uint32_t bar_tfull = bars + 8 * (2 * STAGES);
    const uint32_t bar_tempty = bars + 8 * (2 * STAGES + 1);
    const uint32_t tmem_slot = bars + 8 * (2 * STAGES + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int mt_count = (rows - row0 + MSUB * TM - 1) / (MSUB * TM);
    const int nt_count = (n + CG - 1) / CG;
    const int n_tiles = mt_count * nt_count;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bars + 8 * s, 1);
            mbar_init(bars + 8 * (STAGES + s), 1);       // released by tcgen05.commit
        }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 4);                        // one arrive per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                int mt, nt;
                tile_coords(tile, mt_count, nt_count, mt, nt);
                for (int kb = 0; kb < kslabs; ++kb) {
                    mbar_wait(bars + 8 * (STAGES + stage), phase ^ 1u);
                    const uint32_t full = bars + 8 * stage;
                    mbar_expect_tx(full, STAGE);
                    const uint32_t sa = tiles_s + stage * STAGE;
                    tma_load_3d(sa, &tm_a, kb * KSLAB, row0 + mt * MSUB * TM, a_plane, full);
                    tma_load_3d(sa + TM * KSLAB, &tm_a, kb * KSLAB, row0 + mt * MSUB * TM + TM, a_plane, full);
                    tma_load_3d(sa + A_STAGE, &tm_b, kb * KSLAB, nt * CG, slice0, full);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(TM, NB);
            int stage = 0;
            uint32_t phase = 0, tphase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(bar_tempty, tphase ^ 1u);          // epilogue has drained the accumulators
                tc_fence_after();
                for (int kb = 0; kb < kslabs; ++kb) {
                    mbar_wait(bars + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = tiles_s + stage * STAGE;
                    const uint64_t db = make_desc(sa + A_STAGE);
#pragma unroll
                    for (int sub = 0; sub < MSUB; ++sub) {
                        const uint64_t da = make_desc(sa + sub * TM * KSLAB);
#pragma unroll
                        for (int s = 0; s < KSLAB / 32; ++s)
                            tc_mma_i8(tmem_base + sub * ACC_COLS, da + 2 * s, db + 2 * s, idesc, (kb | s) ? 1u : 0u);
                    }
                    tc_commit(bars + 8 * (STAGES + stage));   // frees the smem stage when these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                tc_commit(bar_tfull);                          // accumulators complete
                tphase ^= 1u;
            }
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        const int quarter = warp & 3;
        uint32_t tphase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            int mt, nt;
            tile_coords(tile, mt_count, nt_count, mt, nt);
            mbar_wait(bar_tfull, tphase);
            tc_fence_after();
            const int k0 = nt * CG;
#pragma unroll
            for (int sub = 0; sub < MSUB; ++sub) {
                const int r = row0 + mt * MSUB * TM + sub * TM + quarter * 32 + lane;
                const bool live = r < rows;
                double c_a = 0.0, c_t = 0.0;
                if (live) {
                    const double* cf = coef + 4 * (size_t)r;
                    c_a = cf[0];
                    c_t = mode == 0 ? cf[2] : (mode == 1 ? cf[3] : cf[1]);
                }
                double* crow = corr + (size_t)(r - row0) * ld_corr;
                const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + sub * ACC_COLS;
#pragma unroll 1
                for (int cb = 0; cb < CG / 8; ++cb) {
                    // Horner over the digit planes, most significant first; the load of plane l-1 is in flight while
                    // plane l is folded in (two 8-register buffers instead of NSL: the SLIM build has 64 registers)
                    int32_t v[2][8];
                    double t[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) t[j] = 0.0;
                    tc_ld8(tbase + (NSL - 1) * CG + cb * 8, v[(NSL - 1) & 1]);
#pragma unroll
                    for (int l = NSL - 1; l >= 0; --l) {
                        tc_wait_ld();
                        if (l > 0) tc_ld8(tbase + (l - 1) * CG + cb * 8, v[(l - 1) & 1]);
#pragma unroll
                        for (int j = 0; j < 8; ++j) t[j] = t[j] * 256.0 + (double)v[l & 1][j];
                    }
                    const int kc = k0 + cb * 8;
                    if (live && kc < n) {
                        double o[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int k = min(kc + j, n - 1);
                            const double tt = t[j] * inv_scale[k];
                            if (mode == 0) {
                                o[j] = c_t * (tt * 4294967296.0);              // slice0 = 4: 256^4
                            } else if (mode == 1) {
                                o[j] = crow[k] + c_t * tt;
                            } else {
                                o[j] = c_a * rk[k] + c_t * tt + crow[k];
                            }
                        }
                        if (mode == 2 && transposed) {
                            // SNP-minor output for the thread-per-SNP solve: rotT[k][r], lanes = consecutive r
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (kc + j < n) rot[(size_t)(kc + j) * ldc + r] = (float)o[j];
                        } else if (mode == 2) {
                            float* dst = rot + (size_t)r * ldc + kc;
                            if (kc + 8 <= n) {
                                reinterpret_cast<float4*>(dst)[0] = make_float4((float)o[0], (float)o[1], (float)o[2], (float)o[3]);
                                reinterpret_cast<float4*>(dst)[1] = make_float4((float)o[4], (float)o[5], (float)o[6], (float)o[7]);
                            } else {
                                for (int j = 0; j < 8 && kc + j < n; ++j) dst[j] = (float)o[j];
                            }
                        } else {
                            double* dst = crow + kc;
                            if (kc + 8 <= n) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) reinterpret_cast<double2*>(dst)[j] = make_double2(o[2 * j], o[2 * j + 1]);
                            } else {
                                for (int j = 0; j < 8 && kc + j < n; ++j) dst[j] = o[j];
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty);
            tphase ^= 1u;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}


template <int NSL, int CG, bool SLIM>
int launch_pass(Model& m, const CUtensorMap& tm_a, const CUtensorMap& tm_b, int a_plane, int slice0, int mode,
                size_t row0, size_t rows, float* out, size_t ld_out, int transposed, cudaStream_t st) {
    constexpr int STAGE = A_STAGE + NSL * CG * KSLAB;
    constexpr int SMEM = (SLIM ? kStagesSlim : kStagesFull) * STAGE + 1024 + 256;
    static bool attr = false;
    if (!attr) {
        JXB_CUDA_OK(cudaFuncSetAttribute(i8_rotate_kernel<NSL, CG, SLIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        JXB_CUDA_OK(cudaFuncSetAttribute(i8_rotate_kernel<NSL, CG, SLIM>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m.device);
    const size_t tiles = ((rows - row0 + MSUB * TM - 1) / (MSUB * TM)) * ((m.n + CG - 1) / CG);
    const int grid = (int)std::min<size_t>((size_t)sms, tiles);
    i8_rotate_kernel<NSL, CG, SLIM><<<grid, NTHREADS, SMEM, st>>>(tm_a, tm_b, a_plane, slice0, mode, (int)row0, (int)rows,
                                                                 (int)m.n, (int)(m.ld8 / KSLAB), m.coef, m.q8_inv_scale,
                                                                 m.q8_rk, m.corr64, m.ld_corr, out, ld_out, transposed);
    note_launch(1);
    JXB_CUDA_OK(cudaGetLastError());
    return 0;
}

// tensor maps of the operand / digit planes and the f64 correction buffer for `corr_rows` rows
int prepare_rotate_tc_impl(Model& m, size_t corr_rows) {
    const size_t ld_corr = round_up(m.n, 32);
    if (!m.corr64 || m.corr_rows < corr_rows || m.ld_corr != ld_corr) {
        if (m.corr64) { cudaDeviceSynchronize(); cudaFree(m.corr64); }
        m.corr64 = nullptr;
        JXB_CUDA_OK(cudaMalloc((void**)&m.corr64, corr_rows * ld_corr * sizeof(double)));
        m.corr_rows = corr_rows;
        m.ld_corr = ld_corr;
    }
    if (!m.tmap_a8 || m.tmap_a8_rows != m.a8_rows) {
        if (!m.tmap_a8) m.tmap_a8 = aligned_alloc(64, sizeof(CUtensorMap));
        int rc = encode_planes((CUtensorMap*)m.tmap_a8, m.a8, m.ld8, m.a8_rows, 3, TM, 1);
        if (rc) return rc;
        m.tmap_a8_rows = m.a8_rows;
    }
    if (!m.tmap_q8_7) {
        m.tmap_q8_7 = aligned_alloc(64, sizeof(CUtensorMap));
        m.tmap_q8_3 = aligned_alloc(64, sizeof(CUtensorMap));
        int rc = encode_planes((CUtensorMap*)m.tmap_q8_7, m.q8, m.ld8, m.q8_rows, 7, 32, 7);
        if (!rc) rc = encode_planes((CUtensorMap*)m.tmap_q8_3, m.q8, m.ld8, m.q8_rows, 7, 64, 3);
        if (rc) return rc;
    }
    return 0;
}

template <bool SLIM>
int rotate_rows(Model& m, size_t row0, size_t rows, bool has_missing, bool transposed_out, cudaStream_t st) {
    const CUtensorMap& ta = *(const CUtensorMap*)m.tmap_a8;
    int rc = launch_pass<3, 64, SLIM>(m, ta, *(const CUtensorMap*)m.tmap_q8_3, /*a_plane=*/1, /*slice0=*/4, /*mode=*/0, row0, rows,
                                      m.rot, m.ldc, 0, st);
    if (!rc && has_missing)
        rc = launch_pass<7, 32, SLIM>(m, ta, *(const CUtensorMap*)m.tmap_q8_7, /*a_plane=*/2, 0, /*mode=*/1, row0, rows, m.rot,
                                      m.ldc, 0, st);
    // transposed: rotT[n][ldr] with ldr = cap_rows shares the rot allocation (cap_rows * ldc >= n * cap_rows floats)
    if (!rc)
        rc = launch_pass<7, 32, SLIM>(m, ta, *(const CUtensorMap*)m.tmap_q8_7, /*a_plane=*/0, 0, /*mode=*/2, row0, rows, m.rot,
                                      transposed_out ? m.cap_rows : m.ldc, transposed_out ? 1 : 0, st);
    return rc;
}

}  // namespace

int prepare_rotate_tc(Model& m, size_t corr_rows) { return prepare_rotate_tc_impl(m, corr_rows); }

// Hand-written tcgen05 path: pass 2 (hom indicator, top 3 slices) -> [pass M (missing indicator)] -> pass D.
int launch_rotate_int8_tc(Model& m, size_t rows, bool has_missing, bool transposed_out, cudaStream_t st) {
    if (rows == 0) return 0;
    int rc = prepare_rotate_tc(m, m.a8_rows);
    if (rc) return rc;
    return rotate_rows<false>(m, 0, rows, has_missing, transposed_out, st);
}

// Row slab [row0, row1) of the decoded batch -> rows [row0, row1) of m.rot (row-major); row0 must be a multiple of 256.
// The correction buffer (prepare_rotate_tc) only needs row1 - row0 rows.  slim: the build that shares an SM with the
// streamed solve (2 stages, <= 64 registers).
int launch_rotate_int8_tc_slab(Model& m, size_t row0, size_t row1, bool has_missing, bool slim, cudaStream_t st) {
    if (row1 <= row0) return 0;
    if (row0 % (MSUB * TM)) return fail(-2, "rotation slabs must start at a multiple of 256 rows");
    if (m.corr_rows < row1 - row0) return fail(-2, "rotation correction buffer is smaller than the slab");
    return slim ? rotate_rows<true>(m, row0, row1, has_missing, false, st) : rotate_rows<false>(m, row0, row1, has_missing, false, st);
}

// registers per thread / shared memory per CTA of the SLIM pass-D kernel (the larger of the slim instantiations)
int rotate_slim_resources(int* regs, int* smem) {
    cudaFuncAttributes f7, f3;
    if (cudaFuncGetAttributes(&f7, i8_rotate_kernel<7, 32, true>) != cudaSuccess) return -1;
    if (cudaFuncGetAttributes(&f3, i8_rotate_kernel<3, 64, true>) != cudaSuccess) return -1;
    *regs = std::max(f7.numRegs, f3.numRegs);
    *smem = (int)std::max(f7.sharedSizeBytes, f3.sharedSizeBytes) + kStagesSlim * (A_STAGE + 7 * 32 * KSLAB) + 1024 + 256;
    return 0;
}

}  // namespace jxb
