// grm.cu -- SURVEY 8(f) N1: the centred additive GRM (VanRaden method 1) from packed PLINK rows, on the int8
// tensor cores (tcgen05.mma.kind::i8, sm_100a).
//
// Replaces, for the B200 path:
//   grm_packed_f32 / grm_packed_f64 core                   src/stats/grm.rs:204-608, 3053-3623
//   decode_additive_grm_block_f32 (method 1)               src/decode/decode.rs:728-900
//   grm_scale_and_symmetrize_raw_f64                       src/stats/grm.rs:2771-2786
//
// Reference arithmetic: z[s][i] = LUT[code] with LUT = {0-mu, 0 (missing), 1-mu, 2-mu} in f32, mu = 2*clamp(p,0,1)
// (p = the prepared row allele frequency), K = sum over SNP blocks of SYRK_f32(Z_block) / sum_s 2p(1-p).  The f32
// SYRK makes the reference's K order-dependent at the 1e-6 level.  Here the contraction is EXACT integer work:
//   k_s   = rint(mu_s * 2^21)                       (mu on a 2^-21 grid: |delta z| <= 2^-22, below f32 resolution of z)
//   Z~    = 2^21 z  in  {-k, 0, 2^21-k, 2^22-k}      a 24-bit integer = 3 balanced base-256 digits (int8 planes)
//   C    += sum_{a,l} 256^(a+l) Z~_a Z~_l^T          int8 x int8 -> int32 in TMEM, recombined in f64 in the epilogue
//   K     = C * 2^-42 / varsum, symmetrised          varsum = sum_s f64(2p(1-p) in f32) exactly as the reference
// so K is the f64-accurate Gram matrix of the grid-rounded genotypes, independent of batch size and GPU count.
//
// Layout: Zt[3 planes][n samples][ldk] int8 with K (= SNP index inside the batch) contiguous: both MMA operands are
// K-major tiles of the SAME tensor.  CTA tile = 256 samples (two M=128 sub-tiles of plane `a`) x 64 samples x 3 planes
// stacked in MMA N (=192); only tiles touching the lower triangle are computed; 3 launches (a = 0,1,2) per batch.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "../../include/jxb200.h"
#include "jxb_common.cuh"
#include "tc05.cuh"

namespace jxb {

namespace {

using namespace tc05;

constexpr int FBITS = 21;                        // fixed-point fraction bits of mu
constexpr int TM = 128;
constexpr int MSUB = 2;
constexpr int CG = 64;                           // j-samples per tile
constexpr int NPL = 3;                           // digit planes
constexpr int NB = NPL * CG;                     // MMA N = 192
constexpr int STAGES = 4;
constexpr int A_STAGE = MSUB * TM * KSLAB;       // 32 KB
constexpr int B_STAGE = NB * KSLAB;              // 24 KB
constexpr int STAGE = A_STAGE + B_STAGE;
constexpr int SMEM = STAGES * STAGE + 1024 + 256;
constexpr int NTHREADS = 192;
constexpr int GROUP_M = 8;                       // row-blocks per raster group (L2 reuse of the A tiles)
constexpr int ACC_COLS = 256;
constexpr int NT_PER_MT = MSUB * TM / CG;        // 4 column tiles per row-block on the diagonal
constexpr size_t MAX_BATCH = 65536;              // 65536 * 128 * 128 = 2^30 < 2^31: the int32 accumulators cannot wrap

struct GrmHandle {
    int device = 0;
    size_t n_full = 0, n = 0;
    int64_t* sample_idx = nullptr;   // device, nullable
    double* C = nullptr;             // [n][n] accumulated 2^42-scaled Gram matrix (lower-triangle tiles)
    double varsum = 0.0;
    size_t snps = 0, snps_used = 0;
    bool finished = false;
    // batch workspace
    uint8_t* packed = nullptr; size_t packed_cap = 0;
    int32_t* counts = nullptr; float* af = nullptr; float* miss = nullptr; size_t rows_cap = 0;
    int8_t* zt = nullptr; size_t ldk = 0;    // [NPL][n][ldk]
    CUtensorMap tm_a, tm_b; size_t tm_ldk = 0;
    int* wave_ctr = nullptr;                 // [NPL] grid-wide tile-start counters, zeroed before every batch
    cudaStream_t st = nullptr;
};

// ---- transposing decode: packed [snp][bps] -> Zt[plane][sample][snp] ------------------------------------------
// One thread = one sample x 32 SNPs (two 16-byte stores per plane = one full 32-byte sector per row).
__global__ void __launch_bounds__(128) grm_decode_t_kernel(const uint8_t* __restrict__ packed, size_t bps, int rows,
                                                           const int64_t* __restrict__ sample_idx, int n,
                                                           const float* __restrict__ af,
                                                           const int32_t* __restrict__ counts, int8_t* __restrict__ zt,
                                                           size_t ldk) {
    __shared__ int8_t dig[32][4][NPL];
    const int s0 = blockIdx.y * 32;
    if (threadIdx.x < 32) {
        const int s = s0 + threadIdx.x;
        int k = 0;
        const bool live = s < rows && (counts == nullptr || counts[4 * s + 3] != 0);   // QC-dropped rows decode as 0
        if (live) {
            float p = af[s];
            p = fminf(fmaxf(p, 0.0f), 1.0f);           // decode.rs:813 clamp(0,1); NaN -> treated as 0 below
            if (!(p == p)) p = 0.0f;
            const float mean_g = 2.0f * p;
            k = (int)llrint((double)mean_g * (double)(1 << FBITS));
        }
        const int vals[4] = {-k, 0, (1 << FBITS) - k, (2 << FBITS) - k};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int v = live ? vals[c] : 0;
#pragma unroll
            for (int l = 0; l < NPL; ++l) {
                const int d = (int)(int8_t)(v & 0xFF);
                dig[threadIdx.x][c][l] = (int8_t)d;
                v = (v - d) >> 8;
            }
        }
    }
    __syncthreads();
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    const long long src = sample_idx ? sample_idx[i] : (long long)i;
    const uint8_t* col = packed + (size_t)(src >> 2);
    const int sh = 2 * (int)(src & 3);
    uint32_t w[NPL][8];
#pragma unroll
    for (int l = 0; l < NPL; ++l)
#pragma unroll
        for (int q = 0; q < 8; ++q) w[l][q] = 0u;
#pragma unroll 8
    for (int t = 0; t < 32; ++t) {
        const int s = s0 + t;
        int code = 1;                                   // padding columns decode as "missing" = 0
        if (s < rows) code = (__ldg(col + (size_t)s * bps) >> sh) & 3;
#pragma unroll
        for (int l = 0; l < NPL; ++l)
            w[l][t >> 2] |= (uint32_t)(uint8_t)dig[t][code][l] << (8 * (t & 3));
    }
#pragma unroll
    for (int l = 0; l < NPL; ++l) {
        uint4* dst = reinterpret_cast<uint4*>(zt + ((size_t)l * n + i) * ldk + s0);
        dst[0] = make_uint4(w[l][0], w[l][1], w[l][2], w[l][3]);
        dst[1] = make_uint4(w[l][4], w[l][5], w[l][6], w[l][7]);
    }
}

// ---- tile raster over the tiles that touch the lower triangle -------------------------------------------------
// Groups of GROUP_M row-blocks; inside a group column tiles ascend and the row-block index is the fast one, so the
// 148 tiles in flight share ~8 A tiles and ~19 B tiles (L2 reuse).  Only valid tiles are numbered: every CTA gets
// equal-length work items back to back, which keeps the CTAs of a wave in k-lockstep -- the property the L2 reuse
// depends on (numbering the skipped above-diagonal slots de-phased the wave and doubled DRAM traffic).
struct Raster {
    int mt_count, nt_count;
    // tiles of group starting at row-block f: a rectangle of `full` columns, then 4-column steps losing one row-block each
    __device__ static int diag_cols(int f, int q, int nt_count) {
        return max(0, min(nt_count, NT_PER_MT * (f + q + 1)) - NT_PER_MT * (f + q));
    }
    __device__ int group_tiles(int f, int gsize) const {
        int t = gsize * min(nt_count, NT_PER_MT * (f + 1));
        for (int q = 1; q < gsize; ++q) t += (gsize - q) * diag_cols(f, q, nt_count);
        return t;
    }
    __device__ int total() const {
        int t = 0;
        for (int f = 0; f < mt_count; f += GROUP_M) t += group_tiles(f, min(GROUP_M, mt_count - f));
        return t;
    }
    __device__ void coords(int tile, int& mt, int& nt) const {
        int f = 0, gsize = 0;
        for (;; f += GROUP_M) {
            gsize = min(GROUP_M, mt_count - f);
            const int gt = group_tiles(f, gsize);
            if (tile < gt) break;
            tile -= gt;
        }
        const int full = gsize * min(nt_count, NT_PER_MT * (f + 1));
        if (tile < full) {
            mt = f + tile % gsize;
            nt = tile / gsize;
            return;
        }
        tile -= full;
        for (int q = 1; q < gsize; ++q) {
            const int rows_q = gsize - q;
            const int cnt = rows_q * diag_cols(f, q, nt_count);
            if (tile < cnt) {
                mt = f + q + tile % rows_q;
                nt = NT_PER_MT * (f + q) + tile / rows_q;
                return;
            }
            tile -= cnt;
        }
        mt = f; nt = 0;   // unreachable for tile < total()
    }
};

// C[i][j] += scale * sum_l 256^l * (Zt[a_plane][i][:] . Zt[l][j][:])   for every tile touching j <= i
__global__ void __launch_bounds__(NTHREADS, 1)
grm_i8_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int a_plane, int n,
              int kslabs, double scale, double* __restrict__ C, int* __restrict__ wave_ctr) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles_s = (raw + 1023u) & ~1023u;
    const uint32_t bars = tiles_s + STAGES * STAGE;
    const uint32_t bar_tfull = bars + 8 * (2 * STAGES);
    const uint32_t bar_tempty = bars + 8 * (2 * STAGES + 1);
    const uint32_t tmem_slot = bars + 8 * (2 * STAGES + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    Raster rs;
    rs.mt_count = (n + MSUB * TM - 1) / (MSUB * TM);
    rs.nt_count = (n + CG - 1) / CG;
    const int n_tiles = rs.total();

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bars + 8 * s, 1);
            mbar_init(bars + 8 * (STAGES + s), 1);
        }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 4);
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
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                int mt, nt;
                rs.coords(tile, mt, nt);
                if (wave_ctr) {
                    // re-phase the wave: every CTA (all co-resident, grid <= SM count) starts its next tile together,
                    // so the slabs the wave streams stay inside the L2 residency window
                    const int target = min(n_tiles, (tile / (int)gridDim.x + 1) * (int)gridDim.x);
                    atomicAdd(wave_ctr, 1);
                    int seen;
                    do {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(wave_ctr) : "memory");
                        if (seen < target) __nanosleep(200);
                    } while (seen < target);
                }
                for (int kb = 0; kb < kslabs; ++kb) {
                    mbar_wait(bars + 8 * (STAGES + stage), phase ^ 1u);
                    const uint32_t full = bars + 8 * stage;
                    mbar_expect_tx(full, STAGE);
                    const uint32_t sa = tiles_s + stage * STAGE;
                    tma_load_3d(sa, &tm_a, kb * KSLAB, mt * MSUB * TM, a_plane, full);
                    tma_load_3d(sa + TM * KSLAB, &tm_a, kb * KSLAB, mt * MSUB * TM + TM, a_plane, full);
                    tma_load_3d(sa + A_STAGE, &tm_b, kb * KSLAB, nt * CG, 0, full);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(TM, NB);
            int stage = 0;
            uint32_t phase = 0, tphase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                int mt, nt;
                rs.coords(tile, mt, nt);
                mbar_wait(bar_tempty, tphase ^ 1u);
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
                    tc_commit(bars + 8 * (STAGES + stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                tc_commit(bar_tfull);
                tphase ^= 1u;
            }
        }
    } else {
        const int quarter = warp & 3;
        uint32_t tphase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            int mt, nt;
            rs.coords(tile, mt, nt);
            mbar_wait(bar_tfull, tphase);
            tc_fence_after();
            const int j0 = nt * CG;
#pragma unroll
            for (int sub = 0; sub < MSUB; ++sub) {
                const int i = mt * MSUB * TM + sub * TM + quarter * 32 + lane;
                const bool live = i < n;
                const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + sub * ACC_COLS;
#pragma unroll
                for (int cb = 0; cb < CG / 8; ++cb) {
                    int32_t v[NPL][8];
#pragma unroll
                    for (int l = 0; l < NPL; ++l) tc_ld8(tbase + l * CG + cb * 8, v[l]);
                    tc_wait_ld();
                    const int jc = j0 + cb * 8;
                    if (live && jc <= i) {               // columns beyond the diagonal are never read back
                        double* dst = C + (size_t)i * n + jc;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (jc + j < n) {
                                double t = 0.0;
#pragma unroll
                                for (int l = NPL - 1; l >= 0; --l) t = t * 256.0 + (double)v[l][j];
                                dst[j] += scale * t;
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

// K = C * inv, mirrored: 32x32 tiles of the lower triangle, transposed through shared memory
__global__ void __launch_bounds__(256) grm_finish_kernel(double* __restrict__ C, int n, double inv) {
    __shared__ double t[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int i = bi * 32 + r, j = bj * 32 + tx;
        double v = 0.0;
        if (i < n && j < n && j <= i) {
            v = C[(size_t)i * n + j] * inv;
            C[(size_t)i * n + j] = v;
        }
        t[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int j = bj * 32 + r, i = bi * 32 + tx;     // write C[j][i] = K[i][j] for j < i
        if (i < n && j < n && j < i) C[(size_t)j * n + i] = t[tx][r];
    }
}

int ensure_batch(GrmHandle& g, size_t rows, size_t bps) {
    if (rows > g.rows_cap) {
        if (g.counts) cudaFree(g.counts);
        if (g.af) cudaFree(g.af);
        if (g.miss) cudaFree(g.miss);
        g.counts = nullptr; g.af = nullptr; g.miss = nullptr;
        JXB_CUDA_OK(cudaMalloc((void**)&g.counts, rows * 4 * sizeof(int32_t)));
        JXB_CUDA_OK(cudaMalloc((void**)&g.af, rows * sizeof(float)));
        JXB_CUDA_OK(cudaMalloc((void**)&g.miss, rows * sizeof(float)));
        g.rows_cap = rows;
    }
    if (bps && rows * bps > g.packed_cap) {
        if (g.packed) cudaFree(g.packed);
        g.packed = nullptr;
        JXB_CUDA_OK(cudaMalloc((void**)&g.packed, rows * bps));
        g.packed_cap = rows * bps;
    }
    const size_t ldk = round_up(rows, (size_t)KSLAB);
    if (ldk > g.ldk) {
        if (g.zt) cudaFree(g.zt);
        g.zt = nullptr;
        JXB_CUDA_OK(cudaMalloc((void**)&g.zt, (size_t)NPL * g.n * ldk));
        g.ldk = ldk;
    }
    if (g.tm_ldk != g.ldk) {
        int rc = encode_planes(&g.tm_a, g.zt, g.ldk, g.n, NPL, TM, 1);
        if (!rc) rc = encode_planes(&g.tm_b, g.zt, g.ldk, g.n, NPL, CG, NPL);
        if (rc) return rc;
        g.tm_ldk = g.ldk;
    }
    return 0;
}

int update_device(GrmHandle& g, const uint8_t* packed_dev, size_t bps, size_t rows, const float* af_dev,
                  const int32_t* keep_counts_dev) {
    cudaStream_t st = g.st;
    const size_t kcols = round_up(rows, (size_t)KSLAB);
    dim3 grid((unsigned)((g.n + 127) / 128), (unsigned)(kcols / 32));
    grm_decode_t_kernel<<<grid, 128, 0, st>>>(packed_dev, bps, (int)rows, g.sample_idx, (int)g.n, af_dev, keep_counts_dev,
                                              g.zt, g.ldk);
    note_launch(1);
    JXB_CUDA_OK(cudaGetLastError());
    static bool attr = false;
    if (!attr) {
        JXB_CUDA_OK(cudaFuncSetAttribute(grm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g.device);
    const size_t mt = (g.n + MSUB * TM - 1) / (MSUB * TM);
    const int grid_mm = (int)std::min<size_t>((size_t)sms, mt * (mt + 1) / 2 * NT_PER_MT + 1);
    double scale = std::ldexp(1.0, -2 * FBITS);
    static const bool wave_sync = !(getenv("JXB_GRM_WAVE_SYNC") && atoi(getenv("JXB_GRM_WAVE_SYNC")) == 0);
    if (!g.wave_ctr) JXB_CUDA_OK(cudaMalloc((void**)&g.wave_ctr, NPL * sizeof(int)));
    JXB_CUDA_OK(cudaMemsetAsync(g.wave_ctr, 0, NPL * sizeof(int), st));
    for (int a = 0; a < NPL; ++a) {
        grm_i8_kernel<<<grid_mm, NTHREADS, SMEM, st>>>(g.tm_a, g.tm_b, a, (int)g.n, (int)(kcols / KSLAB), scale, g.C,
                                                       (wave_sync && grid_mm <= sms) ? g.wave_ctr + a : nullptr);
        note_launch(1);
        JXB_CUDA_OK(cudaGetLastError());
        scale *= 256.0;
    }
    return 0;
}

}  // namespace

}  // namespace jxb

using jxb::fail;
using jxb::GrmHandle;

extern "C" int jxb_grm_create(int device, size_t n_full, const int64_t* sample_idx_host, size_t n_sel, int method,
                              jxb_grm** out) {
    if (!out) return fail(-2, "null argument");
    *out = nullptr;
    if (n_full == 0) return fail(-2, "n_samples must be > 0");
    if (method != 1)
        return fail(-3, "unsupported method=" + std::to_string(method) +
                            "; this build implements 1 (centered additive) on the device");
    if (sample_idx_host && n_sel == 0) return fail(-2, "sample_indices must not be empty");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(-1, "no CUDA device is visible: janusx_b200 has no CPU fallback");
    }
    JXB_CUDA_OK(cudaSetDevice(device));
    GrmHandle* g = new GrmHandle();
    g->device = device;
    g->n_full = n_full;
    g->n = sample_idx_host ? n_sel : n_full;
    if (sample_idx_host) {
        for (size_t i = 0; i < n_sel; ++i)
            if (sample_idx_host[i] < 0 || (size_t)sample_idx_host[i] >= n_full) {
                delete g;
                return fail(-2, "sample index out of range: " + std::to_string((long long)sample_idx_host[i]) +
                                    " >= " + std::to_string(n_full));
            }
    }
    cudaError_t e = cudaStreamCreateWithFlags(&g->st, cudaStreamNonBlocking);
    if (e == cudaSuccess && sample_idx_host) {
        e = cudaMalloc((void**)&g->sample_idx, n_sel * sizeof(int64_t));
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(g->sample_idx, sample_idx_host, n_sel * sizeof(int64_t), cudaMemcpyHostToDevice, g->st);
    }
    if (e == cudaSuccess) e = cudaMalloc((void**)&g->C, g->n * g->n * sizeof(double));
    if (e == cudaSuccess) e = cudaMemsetAsync(g->C, 0, g->n * g->n * sizeof(double), g->st);
    if (e != cudaSuccess) {
        jxb_grm_destroy((jxb_grm*)g);
        return fail(-100, std::string("GRM allocation: ") + cudaGetErrorString(e));
    }
    *out = (jxb_grm*)g;
    return 0;
}

extern "C" void jxb_grm_destroy(jxb_grm* h) {
    GrmHandle* g = (GrmHandle*)h;
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->st) cudaStreamSynchronize(g->st);
    for (void* p : {(void*)g->sample_idx, (void*)g->C, (void*)g->packed, (void*)g->counts, (void*)g->af,
                    (void*)g->miss, (void*)g->zt, (void*)g->wave_ctr})
        if (p) cudaFree(p);
    if (g->st) cudaStreamDestroy(g->st);
    delete g;
}

static int grm_update_impl(GrmHandle* g, const uint8_t* packed, bool on_device, size_t bps, size_t rows,
                           const float* row_maf, const jxb_qc_cfg* qc) {
    if (!g || (!packed && rows)) return fail(-2, "null argument");
    if (g->finished) return fail(-2, "GRM already finished");
    if (bps != (g->n_full + 3) / 4)
        return fail(-2, "packed second dimension mismatch: got " + std::to_string(bps) + ", expected " +
                            std::to_string((g->n_full + 3) / 4));
    JXB_CUDA_OK(cudaSetDevice(g->device));
    if (row_maf && qc) return fail(-2, "QC thresholds apply only when the allele frequency is computed on the device");
    const cudaMemcpyKind to_dev = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    std::vector<float> af_host;
    std::vector<int32_t> counts_host;
    for (size_t r0 = 0; r0 < rows; r0 += jxb::MAX_BATCH) {
        const size_t cur = std::min(jxb::MAX_BATCH, rows - r0);
        int rc = jxb::ensure_batch(*g, cur, on_device ? 0 : bps);
        if (rc) return rc;
        const uint8_t* pk = packed + r0 * bps;
        if (!on_device) {
            JXB_CUDA_OK(cudaMemcpyAsync(g->packed, pk, cur * bps, cudaMemcpyHostToDevice, g->st));
            pk = g->packed;
        }
        af_host.resize(cur);
        if (row_maf) {
            JXB_CUDA_OK(cudaMemcpyAsync(g->af, row_maf + r0, cur * sizeof(float), to_dev, g->st));
        } else {
            // allele frequency over the selected samples, the A3 formula (src/stats/lmm.rs:1262-1323)
            jxb::Model dummy;
            rc = jxb::launch_count_qc(dummy, pk, bps, cur, g->n_full, g->sample_idx, g->n, qc ? qc->maf_thr : 0.0f,
                                      qc ? qc->miss_thr : 1.0f, qc ? qc->het_thr : 0.0f, g->counts, g->af, g->miss, g->st);
            jxb::note_launch(1);
            if (rc) return rc;
            if (qc) {
                counts_host.resize(cur * 4);
                JXB_CUDA_OK(cudaMemcpyAsync(counts_host.data(), g->counts, cur * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, g->st));
            }
        }
        JXB_CUDA_OK(cudaMemcpyAsync(af_host.data(), g->af, cur * sizeof(float), cudaMemcpyDeviceToHost, g->st));
        const bool masked = !row_maf && qc;
        rc = jxb::update_device(*g, pk, bps, cur, g->af, masked ? g->counts : nullptr);
        if (rc) return rc;
        JXB_CUDA_OK(cudaStreamSynchronize(g->st));
        // grm.rs:91-111 / decode.rs:813-815, 842: var = 2p(1-p) in f32, summed in f64 in SNP order
        for (size_t r = 0; r < cur; ++r) {
            if (masked && counts_host[4 * r + 3] == 0) continue;
            ++g->snps_used;
            float p = af_host[r];
            p = p != p ? 0.0f : std::min(std::max(p, 0.0f), 1.0f);
            const float var = 2.0f * p * (1.0f - p);
            const double v = (double)var;
            if (std::isfinite(v) && v > 0.0) g->varsum += v;
        }
        g->snps += cur;
    }
    return 0;
}

extern "C" int jxb_grm_update(jxb_grm* h, const uint8_t* packed_host, size_t bps, size_t rows,
                              const float* row_maf_host, const jxb_qc_cfg* qc) {
    return grm_update_impl((GrmHandle*)h, packed_host, false, bps, rows, row_maf_host, qc);
}

extern "C" int jxb_grm_update_dev(jxb_grm* h, const uint8_t* packed_dev, size_t bps, size_t rows,
                                  const float* row_maf_dev, const jxb_qc_cfg* qc) {
    return grm_update_impl((GrmHandle*)h, packed_dev, true, bps, rows, row_maf_dev, qc);
}

extern "C" int jxb_grm_finish(jxb_grm* h, double* k_host, double* varsum_out) {
    GrmHandle* g = (GrmHandle*)h;
    if (!g) return fail(-2, "null argument");
    JXB_CUDA_OK(cudaSetDevice(g->device));
    if (!g->finished) {
        if (g->snps == 0) return fail(-2, "packed must contain at least one SNP row");
        if (!(std::isfinite(g->varsum) && g->varsum > 0.0))
            return fail(-4, "invalid centered GRM denominator: sum(2p(1-p)) <= 0");
        const unsigned nb = (unsigned)((g->n + 31) / 32);
        jxb::grm_finish_kernel<<<dim3(nb, nb), 256, 0, g->st>>>(g->C, (int)g->n, 1.0 / g->varsum);
        jxb::note_launch(1);
        JXB_CUDA_OK(cudaGetLastError());
        g->finished = true;
    }
    if (k_host)
        JXB_CUDA_OK(cudaMemcpyAsync(k_host, g->C, g->n * g->n * sizeof(double), cudaMemcpyDeviceToHost, g->st));
    JXB_CUDA_OK(cudaStreamSynchronize(g->st));
    if (varsum_out) *varsum_out = g->varsum;
    return 0;
}

extern "C" size_t jxb_grm_rows_used(jxb_grm* h) { return h ? ((GrmHandle*)h)->snps_used : 0; }
extern "C" double* jxb_grm_device_matrix(jxb_grm* h) { return h ? ((GrmHandle*)h)->C : nullptr; }
extern "C" void* jxb_grm_stream(jxb_grm* h) { return h ? (void*)((GrmHandle*)h)->st : nullptr; }

// N1 -> N2 without a host round trip of the f64 matrix: finish, decompose in place, hand back S and the f32 U^T.
extern "C" int jxb_grm_eigh(jxb_grm* h, double diag_shift, double* evals_host, float* ut_f32_host) {
    GrmHandle* g = (GrmHandle*)h;
    if (!g || !evals_host || !ut_f32_host) return fail(-2, "null argument");
    int rc = jxb_grm_finish(h, nullptr, nullptr);
    if (rc) return rc;
    double* w = nullptr;
    float* u32 = nullptr;
    cudaError_t e = cudaMalloc((void**)&w, g->n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&u32, g->n * g->n * sizeof(float));
    if (e != cudaSuccess) {
        if (w) cudaFree(w);
        cudaGetLastError();
        return fail(-100, std::string("GRM eigh buffers: ") + cudaGetErrorString(e));
    }
    rc = jxb_eigh_dev(g->device, g->n, g->C, diag_shift, w, u32, (void*)g->st);
    if (!rc) {
        e = cudaMemcpyAsync(evals_host, w, g->n * sizeof(double), cudaMemcpyDeviceToHost, g->st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ut_f32_host, u32, g->n * g->n * sizeof(float), cudaMemcpyDeviceToHost, g->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g->st);
        if (e != cudaSuccess) rc = fail(-100, std::string("GRM eigh readback: ") + cudaGetErrorString(e));
    }
    cudaFree(w);
    cudaFree(u32);
    return rc;
}
