// k1_decode.cu -- 2-bit PLINK count + QC + compaction + decode/impute/centre (sm_100a).
//
// Replaces, for the B200 path:
//   count_packed_row_counts[_selected_with_excluded]   src/io/gfreader.rs:1378-1395, 1453-1528
//   the QC closure of the unified BED scan              src/stats/lmm.rs:1262-1323
//   decode_centered_block_packed_f32 + value LUT        src/decode/decode.rs:163-271
//   center_decoded_row_inplace_f32                      src/decode/decode.rs:181-189
//
// HBM-bound byte work: one warp streams one packed SNP row with 128-bit loads and popcounts the
// three code classes; a single-CTA ordered scan compacts the kept rows (SNP order preserved);
// one CTA per kept row expands codes through the per-SNP 4-entry value LUT, subtracts the f32
// mean (closed form from the counts -- bit-identical to the reference's sequential f64 sum while
// that sum is exact, i.e. af >= ~1e-4) and writes the row once, as f64, K-padded for the TMA
// tiles of the rotation GEMM.
#include <algorithm>

#include "jxb_common.cuh"

namespace jxb {

namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ void count_word(uint32_t x, int& miss, int& het, int& hom) {
    const uint32_t odd = (x >> 1) & 0x55555555u;
    const uint32_t even = x & 0x55555555u;
    miss += __popc(~odd & even);
    het += __popc(odd & ~even);
    hom += __popc(odd & even);
}

__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// A2 + A3.  counts[r] = {missing, het, hom_alt, keep}.
__global__ void __launch_bounds__(256) count_qc_kernel(const uint8_t* __restrict__ packed, size_t bps, int rows,
                                                       int n_full, const int64_t* __restrict__ sample_idx,
                                                       int n_sel, float maf_thr, float miss_thr, float het_thr,
                                                       int32_t* __restrict__ counts, float* __restrict__ af,
                                                       float* __restrict__ miss_rate) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < rows; r += nwarps) {
        const uint8_t* row = packed + (size_t)r * bps;
        int miss = 0, het = 0, hom = 0;
        int n;
        if (sample_idx == nullptr) {
            n = n_full;
            const int full_bytes = n_full >> 2;
            const int rem = n_full & 3;
            // head bytes up to 16-byte alignment, 128-bit body, byte tail
            const uintptr_t addr = (uintptr_t)row;
            int head = (int)((16 - (addr & 15)) & 15);
            if (head > full_bytes) head = full_bytes;
            for (int b = lane; b < head; b += 32) count_word(row[b], miss, het, hom);
            const int body = (full_bytes - head) >> 4;
            const uint4* v = reinterpret_cast<const uint4*>(row + head);
            for (int q = lane; q < body; q += 32) {
                const uint4 x = __ldg(v + q);
                count_word(x.x, miss, het, hom);
                count_word(x.y, miss, het, hom);
                count_word(x.z, miss, het, hom);
                count_word(x.w, miss, het, hom);
            }
            for (int b = head + (body << 4) + lane; b < full_bytes; b += 32) count_word(row[b], miss, het, hom);
            if (rem > 0 && lane == 0) {
                const uint32_t mask = (1u << (rem * 2)) - 1u;
                count_word((uint32_t)row[full_bytes] & mask, miss, het, hom);
            }
        } else {
            n = n_sel;
            for (int k = lane; k < n_sel; k += 32) {
                const size_t sid = (size_t)sample_idx[k];
                const unsigned code = (row[sid >> 2] >> ((sid & 3) * 2)) & 3u;
                miss += (code == 1u);
                het += (code == 2u);
                hom += (code == 3u);
            }
        }
        miss = warp_sum_i(miss);
        het = warp_sum_i(het);
        hom = warp_sum_i(hom);
        if (lane == 0) {
            // src/stats/lmm.rs:1278-1321, all in f32 like the reference
            int non_missing = n - miss;
            if (non_missing < 0) non_missing = 0;
            const float mr = (n > 0) ? __fdiv_rn((float)miss, (float)n) : 1.0f;
            float a = 0.0f;
            int keep = 1;
            if (mr > miss_thr) {
                keep = 0;
            } else if (non_missing == 0) {
                keep = (maf_thr > 0.0f) ? 0 : 1;
            } else {
                if (het_thr > 0.0f) {
                    const float hr = __fdiv_rn((float)het, (float)non_missing);
                    if (hr > het_thr) keep = 0;
                }
                if (keep) {
                    const int alt_sum = het + 2 * hom;
                    const float alt = __fdiv_rn((float)alt_sum, __fmul_rn(2.0f, (float)non_missing));
                    const float other = __fsub_rn(1.0f, alt);
                    const float maf_v = alt < other ? alt : other;
                    if (maf_v < maf_thr) keep = 0; else a = alt;
                }
            }
            counts[4 * r + 0] = miss;
            counts[4 * r + 1] = het;
            counts[4 * r + 2] = hom;
            counts[4 * r + 3] = keep;
            af[r] = a;
            miss_rate[r] = mr;
        }
    }
}

// Ordered compaction of kept rows (single CTA; rows per batch <= 2^20).
__global__ void __launch_bounds__(1024) compact_kernel(const int32_t* __restrict__ counts, int rows,
                                                       int32_t* __restrict__ src_row, int32_t* __restrict__ n_kept) {
    __shared__ int warp_tot[32];
    __shared__ int base_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (rows + 1023) / 1024;
    const int lo = tid * per;
    const int hi = min(rows, lo + per);
    int cnt = 0;
    for (int r = lo; r < hi; ++r) cnt += counts[4 * r + 3] != 0;
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int t = warp_tot[lane];
        int s = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFull, s, o);
            if (lane >= o) s += v;
        }
        warp_tot[lane] = s - t;  // exclusive
        if (lane == 31) base_s = s;
    }
    __syncthreads();
    int pos = warp_tot[wid] + inc - cnt;
    for (int r = lo; r < hi; ++r)
        if (counts[4 * r + 3] != 0) src_row[pos++] = r;
    if (tid == 0) n_kept[0] = base_s;
}

// src/decode/decode.rs:121-145 on the four raw LUT entries
__device__ __forceinline__ float model_apply(int model, float raw) {
    const double g = (double)raw;
    switch (model) {
        case 1: return (g > 0.0) ? 1.0f : 0.0f;
        case 2: return (fabs(g - 2.0) < 1e-6) ? 1.0f : 0.0f;
        case 3: return (fabs(g - 1.0) < 1e-6) ? 1.0f : 0.0f;
        default: return raw;
    }
}

// A4.  One CTA per kept row.  af/counts are indexed by SOURCE row.
__global__ void __launch_bounds__(256) decode_center_kernel(const uint8_t* __restrict__ packed, size_t bps,
                                                            const int32_t* __restrict__ src_row,
                                                            const int32_t* __restrict__ n_kept, int max_rows,
                                                            int n_full, const int64_t* __restrict__ sample_idx,
                                                            int n, const float* __restrict__ af_by_src,
                                                            const int32_t* __restrict__ counts_by_src, int model,
                                                            double* __restrict__ g64, size_t ldk,
                                                            float* __restrict__ g32, size_t ld32,
                                                            const float* __restrict__ row_lut) {
    const int rows = n_kept ? min(*n_kept, max_rows) : max_rows;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const int src = src_row ? src_row[r] : r;
        const uint8_t* row = packed + (size_t)src * bps;
        float c0, c1, c2, c3;
        if (row_lut) {
            // caller-supplied value LUT per source row, indexed by the 2-bit PLINK code (00, 01 = missing, 10, 11); no
            // centring.  Serves BedChunkReaderFromMeta (src/io/gfreader.rs:7623-7730: [(0-m), 0, (1-m), (2-m)]) and the raw
            // BedChunkReader.next_chunk (src/io/gfcore.rs:405-480: [0, imputed, 1, 2] or flipped [2, imputed, 1, 0])
            c0 = row_lut[4 * src + 0];
            c1 = row_lut[4 * src + 1];
            c2 = row_lut[4 * src + 2];
            c3 = row_lut[4 * src + 3];
        } else {
        // decode.rs:218-219: mean_g = (2.0 * maf as f64).max(0.0) as f32; row_flip (bit 1 of the keep word, set only from
        // prepared row metadata) reverses the raw LUT to [2, mean_g, 1, 0] (decode.rs:163-178)
        double mg = 2.0 * (double)af_by_src[src];
        if (!(mg > 0.0)) mg = 0.0;
        const float mean_g = (float)mg;
        const bool flip = (counts_by_src[4 * src + 3] & 2) != 0;
        const float l0 = model_apply(model, flip ? 2.0f : 0.0f), l1 = model_apply(model, mean_g);
        const float l2 = model_apply(model, 1.0f), l3 = model_apply(model, flip ? 0.0f : 2.0f);
        const int nmiss = counts_by_src[4 * src + 0], nhet = counts_by_src[4 * src + 1];
        const int nhom = counts_by_src[4 * src + 2];
        const int n0 = n - nmiss - nhet - nhom;
        // decode.rs:185: mean = (sum_j v_j as f64) / n as f32 -- closed form over the code counts
        const double sum = (double)n0 * (double)l0 + (double)nmiss * (double)l1 + (double)nhet * (double)l2 +
                           (double)nhom * (double)l3;
        const float mean = (n > 0) ? (float)(sum / (double)n) : 0.0f;
        c0 = __fsub_rn(l0, mean); c1 = __fsub_rn(l1, mean); c2 = __fsub_rn(l2, mean); c3 = __fsub_rn(l3, mean);
        }
        double* dst64 = g64 ? g64 + (size_t)r * ldk : nullptr;
        float* dst32 = g32 ? g32 + (size_t)r * ld32 : nullptr;
        if (sample_idx == nullptr) {
            const int nbytes = (n_full + 3) >> 2;
            for (int b = threadIdx.x; b < nbytes; b += blockDim.x) {
                const unsigned byte = row[b];
                float v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned code = (byte >> (2 * k)) & 3u;
                    v[k] = code == 0u ? c0 : (code == 1u ? c1 : (code == 2u ? c2 : c3));
                }
                const int j = b << 2;
                if (j + 3 < n) {
                    if (dst64) {
                        reinterpret_cast<double2*>(dst64 + j)[0] = make_double2((double)v[0], (double)v[1]);
                        reinterpret_cast<double2*>(dst64 + j)[1] = make_double2((double)v[2], (double)v[3]);
                    }
                    if (dst32) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) dst32[j + k] = v[k];
                    }
                } else {
                    for (int k = 0; k < 4 && j + k < n; ++k) {
                        if (dst64) dst64[j + k] = (double)v[k];
                        if (dst32) dst32[j + k] = v[k];
                    }
                }
            }
        } else {
            for (int j = threadIdx.x; j < n; j += blockDim.x) {
                const size_t sid = (size_t)sample_idx[j];
                const unsigned code = (row[sid >> 2] >> ((sid & 3) * 2)) & 3u;
                const float v = code == 0u ? c0 : (code == 1u ? c1 : (code == 2u ? c2 : c3));
                if (dst64) dst64[j] = (double)v;
                if (dst32) dst32[j] = v;
            }
        }
    }
}

__global__ void widen_kernel(const float* __restrict__ src, size_t ld_src, int rows, int n,
                             double* __restrict__ dst, size_t ldk) {
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
        const float* s = src + (size_t)r * ld_src;
        double* d = dst + (size_t)r * ldk;
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) d[j] = (double)s[j];
    }
}

}  // namespace

int launch_count_qc(const Model& m, const uint8_t* packed, size_t bps, size_t rows, size_t n_full,
                    const int64_t* sample_idx, size_t n_sel, float maf_thr, float miss_thr, float het_thr,
                    int32_t* counts, float* af, float* miss_rate, cudaStream_t st) {
    (void)m;
    if (rows == 0) return 0;
    const int blocks = (int)std::min<size_t>((rows + 7) / 8, 148 * 8);
    count_qc_kernel<<<blocks, 256, 0, st>>>(packed, bps, (int)rows, (int)n_full, sample_idx, (int)n_sel, maf_thr,
                                            miss_thr, het_thr, counts, af, miss_rate);
    JXB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_compact(const int32_t* counts, size_t rows, int32_t* src_row, int32_t* n_kept, cudaStream_t st) {
    compact_kernel<<<1, 1024, 0, st>>>(counts, (int)rows, src_row, n_kept);
    JXB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_decode_center(const uint8_t* packed, size_t bps, const int32_t* src_row, const int32_t* n_kept,
                         size_t max_rows, size_t n_full, const int64_t* sample_idx, size_t n,
                         const float* af_by_src, const int32_t* counts_by_src, int model_code, double* g64,
                         size_t ldk, float* g32, size_t ld32, cudaStream_t st, const float* row_lut) {
    if (max_rows == 0) return 0;
    const int blocks = (int)std::min<size_t>(max_rows, 148 * 16);
    decode_center_kernel<<<blocks, 256, 0, st>>>(packed, bps, src_row, n_kept, (int)max_rows, (int)n_full,
                                                 sample_idx, (int)n, af_by_src, counts_by_src, model_code, g64, ldk,
                                                 g32, ld32, row_lut);
    JXB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_widen_f32(const float* src, size_t ld_src, size_t rows, size_t n, double* dst, size_t ldk,
                     cudaStream_t st) {
    if (rows == 0) return 0;
    dim3 grid((unsigned)std::min<size_t>((n + 255) / 256, 64), (unsigned)std::min<size_t>(rows, 4096));
    widen_kernel<<<grid, 256, 0, st>>>(src, ld_src, (int)rows, (int)n, dst, ldk);
    JXB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace jxb
