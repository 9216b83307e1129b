// k2_int8.cu -- exact eigen-rotation through INT8 tensor cores by fixed-point slicing of U^T.
//
// Same arithmetic contract as k2_rotate.cu (f32-valued inputs, f64-accurate dot products, one rounding to
// f32), ~10x less tensor time than the FP64 DMMA GEMM:
//
//   g[r,j] in {c0, c1, c2, c3}   (centred f32 values of codes 00 / 01=missing / 10 / 11, src/decode/decode.rs:163-189)
//   rot[r,k] = sum_j g[r,j] U^T[k,j]
//            = c0 R_k + (c2-c0) T_D[r,k] + (c3-2c2+c0) T_2[r,k] + (c1-c0) T_m[r,k]
//   with  D = dosage (0/1/2, missing -> 0),  T_D = D U,  T_2 = [code==11] U,  T_m = [code==01] U,  R_k = sum_j U^T[k,j].
//
// U^T row k is written EXACTLY as a 54-bit fixed-point integer Q[k,j] = floor(U^T[k,j] 2^shift_k) (an f32 has 24
// significant bits; entries below 2^-29 of the row maximum lose their lowest bits: absolute error < 2^-53 of the
// row maximum per entry, i.e. below f64 summation noise) and split into 7 balanced base-256 digits (int8).
// D / indicators are int8 too, so every slice product D x digit accumulates EXACTLY in int32 on the tensor
// cores; the 7 (resp. 3, 7) int32 slice results are recombined by Horner in f64.  The coefficient of T_2 is a
// pure f32-rounding residual (|.| <= 2e-7), so T_2 only needs its top 3 digits.
//
// This file holds: the one-off slicing of U^T, the int8 decode of packed genotypes, the f64 recombination, and
// the slice GEMMs of rotation variant 2, which run through cuBLASLt (int8 tensor cores, loaded with dlopen so the
// library stays optional).  The default variant 3 (k2_i8mma.cu) shares the slicing and the decode with this file and
// replaces the library GEMMs + recombination kernel by one hand-written tcgen05 kernel per pass.
#include <cublasLt.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "jxb_common.cuh"

namespace jxb {

namespace {

constexpr int kSlices = 7;
constexpr int kTopSlices = 3;   // digits 4..6 for the rounding-residual term

// One CTA per eigenvector row: row maximum -> shift; digits; digit sums -> R_k.
__global__ void __launch_bounds__(256) slice_ut_kernel(const double* __restrict__ ut, size_t ldk, int n,
                                                       int8_t* __restrict__ q8, size_t ld8, size_t slice_stride,
                                                       double* __restrict__ inv_scale, double* __restrict__ rk) {
    __shared__ double red[256];
    __shared__ long long dsum[kSlices];
    const int k = blockIdx.x;
    const double* u = ut + (size_t)k * ldk;
    double mx = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) mx = fmax(mx, fabs(u[j]));
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    mx = red[0];
    int ex = 0;
    if (mx > 0.0) frexp(mx, &ex);            // mx = m 2^ex, m in [0.5,1)  =>  |u| < 2^ex
    const int shift = 53 - ex;               // |u| 2^shift < 2^53: exact in f64, fits 7 balanced digits
    const double scale = ldexp(1.0, shift);
    if (threadIdx.x < kSlices) dsum[threadIdx.x] = 0;
    __syncthreads();
    long long loc[kSlices];
#pragma unroll
    for (int l = 0; l < kSlices; ++l) loc[l] = 0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        long long rem = __double2ll_rd(u[j] * scale);
#pragma unroll
        for (int l = 0; l < kSlices; ++l) {
            const long long d = ((rem + 128) & 255) - 128;   // balanced digit in [-128, 127]
            q8[(size_t)l * slice_stride + (size_t)k * ld8 + j] = (int8_t)d;
            loc[l] += d;
            rem = (rem - d) >> 8;
        }
    }
#pragma unroll
    for (int l = 0; l < kSlices; ++l) atomicAdd((unsigned long long*)&dsum[l], (unsigned long long)loc[l]);
    __syncthreads();
    if (threadIdx.x == 0) {
        double acc = 0.0;
        for (int l = kSlices - 1; l >= 0; --l) acc = acc * 256.0 + (double)dsum[l];
        inv_scale[k] = ldexp(1.0, -shift);
        rk[k] = acc * ldexp(1.0, -shift);
    }
}

__device__ __forceinline__ float model_apply8(int model, float raw) {
    const double g = (double)raw;
    switch (model) {
        case 1: return (g > 0.0) ? 1.0f : 0.0f;
        case 2: return (fabs(g - 2.0) < 1e-6) ? 1.0f : 0.0f;
        case 3: return (fabs(g - 1.0) < 1e-6) ? 1.0f : 0.0f;
        default: return raw;
    }
}

// Packed row -> int8 operand rows (dosage, hom indicator, missing indicator) + per-row f64 coefficients.
// coef[r] = {c0, c2-c0, c3-2c2+c0, c1-c0}.  Same LUT / mean arithmetic as decode_center_kernel (k1_decode.cu).
__global__ void __launch_bounds__(256) decode_int8_kernel(const uint8_t* __restrict__ packed, size_t bps,
                                                          const int32_t* __restrict__ src_row,
                                                          const int32_t* __restrict__ n_kept, int max_rows, int n_full,
                                                          const int64_t* __restrict__ sample_idx, int n,
                                                          const float* __restrict__ af_by_src,
                                                          const int32_t* __restrict__ counts_by_src, int model,
                                                          int8_t* __restrict__ a_d, int8_t* __restrict__ a_2,
                                                          int8_t* __restrict__ a_m, size_t ld8,
                                                          double* __restrict__ coef, int32_t* __restrict__ any_missing) {
    const int rows = n_kept ? min(*n_kept, max_rows) : max_rows;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const int src = src_row ? src_row[r] : r;
        const uint8_t* row = packed + (size_t)src * bps;
        if (threadIdx.x == 0) {
            double mg = 2.0 * (double)af_by_src[src];
            if (!(mg > 0.0)) mg = 0.0;
            const float mean_g = (float)mg;
            const bool flip = (counts_by_src[4 * src + 3] & 2) != 0;     // prepared row_flip: LUT [2, mean_g, 1, 0]
            const float l0 = model_apply8(model, flip ? 2.0f : 0.0f), l1 = model_apply8(model, mean_g);
            const float l2 = model_apply8(model, 1.0f), l3 = model_apply8(model, flip ? 0.0f : 2.0f);
            const int nmiss = counts_by_src[4 * src + 0], nhet = counts_by_src[4 * src + 1];
            const int nhom = counts_by_src[4 * src + 2];
            const int n0 = n - nmiss - nhet - nhom;
            const double sum = (double)n0 * (double)l0 + (double)nmiss * (double)l1 + (double)nhet * (double)l2 +
                               (double)nhom * (double)l3;
            const float mean = (n > 0) ? (float)(sum / (double)n) : 0.0f;
            const double c0 = (double)__fsub_rn(l0, mean), c1 = (double)__fsub_rn(l1, mean);
            const double c2 = (double)__fsub_rn(l2, mean), c3 = (double)__fsub_rn(l3, mean);
            coef[4 * r + 0] = c0;
            coef[4 * r + 1] = c2 - c0;
            coef[4 * r + 2] = c3 - 2.0 * c2 + c0;
            coef[4 * r + 3] = c1 - c0;
            if (nmiss > 0) atomicOr(any_missing, 1);
        }
        int8_t* dd = a_d + (size_t)r * ld8;
        int8_t* d2 = a_2 + (size_t)r * ld8;
        int8_t* dm = a_m + (size_t)r * ld8;
        if (sample_idx == nullptr) {
            const int nbytes = (n_full + 3) >> 2;
            for (int b = threadIdx.x; b < nbytes; b += blockDim.x) {
                const unsigned byte = row[b];
                const int j = b << 2;
                uint32_t wd = 0, w2 = 0, wm = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned code = (byte >> (2 * k)) & 3u;
                    const unsigned dos = code == 2u ? 1u : (code == 3u ? 2u : 0u);
                    const bool live = j + k < n;
                    wd |= (live ? dos : 0u) << (8 * k);
                    w2 |= ((live && code == 3u) ? 1u : 0u) << (8 * k);
                    wm |= ((live && code == 1u) ? 1u : 0u) << (8 * k);
                }
                *reinterpret_cast<uint32_t*>(dd + j) = wd;      // ld8 is a multiple of 128: 4-byte stores stay in the row
                *reinterpret_cast<uint32_t*>(d2 + j) = w2;
                *reinterpret_cast<uint32_t*>(dm + j) = wm;
            }
        } else {
            for (int j = threadIdx.x; j < n; j += blockDim.x) {
                const size_t sid = (size_t)sample_idx[j];
                const unsigned code = (row[sid >> 2] >> ((sid & 3) * 2)) & 3u;
                dd[j] = (int8_t)(code == 2u ? 1 : (code == 3u ? 2 : 0));
                d2[j] = (int8_t)(code == 3u);
                dm[j] = (int8_t)(code == 1u);
            }
        }
    }
}

// rot[r,k] (f32) from the int32 slice results.  cD: kSlices buffers [rows][n], c2: kTopSlices buffers, cM: kSlices
// buffers (null when the batch has no missing call).  One thread per output.
__global__ void __launch_bounds__(256) recombine_kernel(const int32_t* __restrict__ c_d, const int32_t* __restrict__ c_2,
                                                        const int32_t* __restrict__ c_m, size_t buf_stride, int rows,
                                                        int n, int n_c, const double* __restrict__ coef,
                                                        const double* __restrict__ inv_scale,
                                                        const double* __restrict__ rk, float* __restrict__ rot,
                                                        size_t ldc) {
    const size_t total = (size_t)rows * n;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / n), k = (int)(e - (size_t)r * n);
        const size_t ce = (size_t)r * n_c + k;
        double td = 0.0;
#pragma unroll
        for (int l = kSlices - 1; l >= 0; --l) td = td * 256.0 + (double)c_d[(size_t)l * buf_stride + ce];
        double t2 = 0.0;
#pragma unroll
        for (int l = kTopSlices - 1; l >= 0; --l) t2 = t2 * 256.0 + (double)c_2[(size_t)l * buf_stride + ce];
        t2 *= 4294967296.0;   // 256^4: the top digits are 4..6
        const double is = inv_scale[k];
        const double* cf = coef + 4 * (size_t)r;
        double v = cf[0] * rk[k] + cf[1] * (td * is) + cf[2] * (t2 * is);
        if (c_m) {
            double tm = 0.0;
#pragma unroll
            for (int l = kSlices - 1; l >= 0; --l) tm = tm * 256.0 + (double)c_m[(size_t)l * buf_stride + ce];
            v += cf[3] * (tm * is);
        }
        rot[(size_t)r * ldc + k] = (float)v;
    }
}

// ---- cuBLASLt (optional, dlopen) ------------------------------------------------------------------------
struct Lt {
    void* so = nullptr;
    cublasLtHandle_t handle = nullptr;
    decltype(&cublasLtCreate) create = nullptr;
    decltype(&cublasLtMatmulDescCreate) desc_create = nullptr;
    decltype(&cublasLtMatmulDescSetAttribute) desc_set = nullptr;
    decltype(&cublasLtMatrixLayoutCreate) layout_create = nullptr;
    decltype(&cublasLtMatrixLayoutDestroy) layout_destroy = nullptr;
    decltype(&cublasLtMatmulDescDestroy) desc_destroy = nullptr;
    decltype(&cublasLtMatmulPreferenceCreate) pref_create = nullptr;
    decltype(&cublasLtMatmulPreferenceSetAttribute) pref_set = nullptr;
    decltype(&cublasLtMatmulPreferenceDestroy) pref_destroy = nullptr;
    decltype(&cublasLtMatmulAlgoGetHeuristic) heuristic = nullptr;
    decltype(&cublasLtMatmul) matmul = nullptr;
    bool ok = false;
};

Lt& lt() {
    static Lt g;
    static bool tried = false;
    if (tried) return g;
    tried = true;
    for (const char* name : {"libcublasLt.so.12", "libcublasLt.so"}) {
        g.so = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (g.so) break;
    }
    if (!g.so) return g;
#define JXB_SYM(field, sym) g.field = (decltype(g.field))dlsym(g.so, #sym); if (!g.field) return g;
    JXB_SYM(create, cublasLtCreate)
    JXB_SYM(desc_create, cublasLtMatmulDescCreate)
    JXB_SYM(desc_set, cublasLtMatmulDescSetAttribute)
    JXB_SYM(layout_create, cublasLtMatrixLayoutCreate)
    JXB_SYM(layout_destroy, cublasLtMatrixLayoutDestroy)
    JXB_SYM(desc_destroy, cublasLtMatmulDescDestroy)
    JXB_SYM(pref_create, cublasLtMatmulPreferenceCreate)
    JXB_SYM(pref_set, cublasLtMatmulPreferenceSetAttribute)
    JXB_SYM(pref_destroy, cublasLtMatmulPreferenceDestroy)
    JXB_SYM(heuristic, cublasLtMatmulAlgoGetHeuristic)
    JXB_SYM(matmul, cublasLtMatmul)
#undef JXB_SYM
    if (g.create(&g.handle) != CUBLAS_STATUS_SUCCESS) return g;
    g.ok = true;
    return g;
}

// C[rows][n] (row-major int32) = A[rows][K] (int8, K-major) x S[n][K]^T (int8, K-major).
// Column-major view: C_cm[n x rows] = S_cm^T[n x K] * A_cm[K x rows]  ("TN", the layout IMMA wants).
int int8_gemm_tn(const int8_t* s, const int8_t* a, int32_t* c, int n, int n_c, int rows, int kdim, size_t ld8, void* workspace,
                 size_t ws_bytes, cudaStream_t st) {
    Lt& L = lt();
    if (!L.ok) return fail(-110, "cuBLASLt could not be loaded: the int8 rotation variant is unavailable");
    cublasLtMatmulDesc_t desc = nullptr;
    cublasLtMatrixLayout_t la = nullptr, lb = nullptr, lc = nullptr;
    cublasLtMatmulPreference_t pref = nullptr;
    int rc = 0;
    cublasOperation_t opT = CUBLAS_OP_T, opN = CUBLAS_OP_N;
    if (L.desc_create(&desc, CUBLAS_COMPUTE_32I, CUDA_R_32I) != CUBLAS_STATUS_SUCCESS) return fail(-111, "cublasLt desc");
    L.desc_set(desc, CUBLASLT_MATMUL_DESC_TRANSA, &opT, sizeof opT);
    L.desc_set(desc, CUBLASLT_MATMUL_DESC_TRANSB, &opN, sizeof opN);
    L.layout_create(&la, CUDA_R_8I, (uint64_t)kdim, (uint64_t)n, (int64_t)ld8);      // S_cm: K x n
    L.layout_create(&lb, CUDA_R_8I, (uint64_t)kdim, (uint64_t)rows, (int64_t)ld8);   // A_cm: K x rows
    L.layout_create(&lc, CUDA_R_32I, (uint64_t)n, (uint64_t)rows, (int64_t)n_c);     // C_cm: n x rows, ld n_c
    L.pref_create(&pref);
    L.pref_set(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws_bytes, sizeof ws_bytes);
    cublasLtMatmulHeuristicResult_t heur;
    int found = 0;
    if (L.heuristic(L.handle, desc, la, lb, lc, lc, pref, 1, &heur, &found) != CUBLAS_STATUS_SUCCESS || found == 0) {
        rc = fail(-112, "cuBLASLt has no int8 TN algorithm for this shape");
    } else {
        const int32_t alpha = 1, beta = 0;
        cublasStatus_t s2 = L.matmul(L.handle, desc, &alpha, s, la, a, lb, &beta, c, lc, c, lc, &heur.algo, workspace,
                                     ws_bytes, st);
        if (s2 != CUBLAS_STATUS_SUCCESS) rc = fail(-113, "cublasLtMatmul(int8) failed with status " + std::to_string((int)s2));
    }
    L.pref_destroy(pref);
    L.layout_destroy(la); L.layout_destroy(lb); L.layout_destroy(lc);
    L.desc_destroy(desc);
    return rc;
}

}  // namespace

// One-off: slice the resident f64 copy of U^T into 7 int8 digit planes.
int prepare_int8_slices(Model& m, cudaStream_t st) {
    if (!m.ut) return fail(-3, "model was created without U^T; rotation is unavailable");
    if (m.q8) return 0;
    m.ld8 = round_up(m.n, 128);
    m.q8_rows = round_up(m.n, 128);
    const size_t plane = m.q8_rows * m.ld8;
    JXB_CUDA_OK(cudaMalloc((void**)&m.q8, (size_t)kSlices * plane));
    JXB_CUDA_OK(cudaMemsetAsync(m.q8, 0, (size_t)kSlices * plane, st));
    JXB_CUDA_OK(cudaMalloc((void**)&m.q8_inv_scale, m.n * sizeof(double)));
    JXB_CUDA_OK(cudaMalloc((void**)&m.q8_rk, m.n * sizeof(double)));
    slice_ut_kernel<<<(unsigned)m.n, 256, 0, st>>>(m.ut, m.ldk, (int)m.n, m.q8, m.ld8, plane, m.q8_inv_scale, m.q8_rk);
    note_launch(1);
    JXB_CUDA_OK(cudaGetLastError());
    return 0;
}

int ensure_int8_workspace(Model& m, size_t rows_cap) {
    if (m.a8 && m.a8_rows >= rows_cap) return 0;
    if (m.a8) { cudaFree(m.a8); cudaFree(m.coef); m.a8 = nullptr; m.coef = nullptr; }
    m.a8_rows = round_up(rows_cap, 128);
    JXB_CUDA_OK(cudaMalloc((void**)&m.a8, 3 * m.a8_rows * m.ld8));
    JXB_CUDA_OK(cudaMemset(m.a8, 0, 3 * m.a8_rows * m.ld8));
    JXB_CUDA_OK(cudaMalloc((void**)&m.coef, 4 * m.a8_rows * sizeof(double)));
    if (!m.flags8) JXB_CUDA_OK(cudaMalloc((void**)&m.flags8, 4 * sizeof(int32_t)));
    return 0;
}

int launch_decode_int8(Model& m, const uint8_t* packed, size_t bps, const int32_t* src_row, const int32_t* n_kept,
                       size_t max_rows, size_t n_full, const int64_t* sample_idx, const float* af_by_src,
                       const int32_t* counts_by_src, int model_code, cudaStream_t st) {
    if (max_rows == 0) return 0;
    JXB_CUDA_OK(cudaMemsetAsync(m.flags8, 0, sizeof(int32_t), st));
    const size_t plane = m.a8_rows * m.ld8;
    const int blocks = (int)std::min<size_t>(max_rows, 148 * 16);
    decode_int8_kernel<<<blocks, 256, 0, st>>>(packed, bps, src_row, n_kept, (int)max_rows, (int)n_full, sample_idx,
                                               (int)m.n, af_by_src, counts_by_src, model_code, m.a8, m.a8 + plane,
                                               m.a8 + 2 * plane, m.ld8, m.coef, m.flags8);
    note_launch(1);
    JXB_CUDA_OK(cudaGetLastError());
    return 0;
}

// Slice GEMMs (cuBLASLt) + recombination for `rows` decoded int8 rows -> m.rot (f32, row-major).
// Works in row sub-blocks so the int32 slice results stay bounded (<= 17 * sub * n * 4 bytes).
int launch_rotate_int8_lib(Model& m, size_t rows, bool has_missing, cudaStream_t st) {
    if (rows == 0) return 0;
    const size_t n = m.n;
    const size_t n_c = round_up(n, 16);
    const size_t sub = std::min<size_t>(rows, std::max<size_t>(256, ((size_t)3 << 30) / (17 * n * 4) / 128 * 128));
    const int nbuf = kSlices + kTopSlices + kSlices;
    const size_t buf_stride = sub * n_c;
    if (m.c32_elems < (size_t)nbuf * buf_stride) {
        if (m.c32) cudaFree(m.c32);
        m.c32 = nullptr;
        JXB_CUDA_OK(cudaMalloc((void**)&m.c32, (size_t)nbuf * buf_stride * sizeof(int32_t)));
        m.c32_elems = (size_t)nbuf * buf_stride;
    }
    if (!m.lt_ws) {
        m.lt_ws_bytes = (size_t)64 << 20;
        JXB_CUDA_OK(cudaMalloc(&m.lt_ws, m.lt_ws_bytes));
    }
    const size_t plane_q = m.q8_rows * m.ld8, plane_a = m.a8_rows * m.ld8;
    int32_t* c_d = m.c32;
    int32_t* c_2 = m.c32 + (size_t)kSlices * buf_stride;
    int32_t* c_m = m.c32 + (size_t)(kSlices + kTopSlices) * buf_stride;
    for (size_t r0 = 0; r0 < rows; r0 += sub) {
        const size_t rr = std::min(sub, rows - r0);
        const int8_t* a_d = m.a8 + r0 * m.ld8;
        const int8_t* a_2 = m.a8 + plane_a + r0 * m.ld8;
        const int8_t* a_m = m.a8 + 2 * plane_a + r0 * m.ld8;
        int rc = 0;
        for (int l = 0; l < kSlices && !rc; ++l)
            rc = int8_gemm_tn(m.q8 + (size_t)l * plane_q, a_d, c_d + (size_t)l * buf_stride, (int)n, (int)n_c, (int)rr, (int)m.ld8,
                              m.ld8, m.lt_ws, m.lt_ws_bytes, st);
        for (int l = 0; l < kTopSlices && !rc; ++l)
            rc = int8_gemm_tn(m.q8 + (size_t)(kSlices - kTopSlices + l) * plane_q, a_2, c_2 + (size_t)l * buf_stride,
                              (int)n, (int)n_c, (int)rr, (int)m.ld8, m.ld8, m.lt_ws, m.lt_ws_bytes, st);
        for (int l = 0; l < kSlices && has_missing && !rc; ++l)
            rc = int8_gemm_tn(m.q8 + (size_t)l * plane_q, a_m, c_m + (size_t)l * buf_stride, (int)n, (int)n_c, (int)rr, (int)m.ld8,
                              m.ld8, m.lt_ws, m.lt_ws_bytes, st);
        if (rc) return rc;
        note_launch(kSlices + kTopSlices + (has_missing ? kSlices : 0));
        const int blocks = (int)std::min<size_t>((rr * n + 255) / 256, (size_t)148 * 32);
        recombine_kernel<<<blocks, 256, 0, st>>>(c_d, c_2, has_missing ? c_m : nullptr, buf_stride, (int)rr, (int)n, (int)n_c,
                                                 m.coef + 4 * r0, m.q8_inv_scale, m.q8_rk, m.rot + r0 * m.ldc, m.ldc);
        note_launch(1);
        JXB_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

}  // namespace jxb
