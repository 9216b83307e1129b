// k3_inst.cu -- one static-covariate-count instantiation of the K3 kernels per translation unit
// (compiled with -DJXB_P=1..8 by janusx_b200/build.py so the heavy unrolled kernels build in parallel).
#include <algorithm>

#include "k3_solve.cuh"

#ifndef JXB_P
#error "compile with -DJXB_P=<covariate columns>"
#endif

#define JXB_CAT2(a, b) a##b
#define JXB_CAT(a, b) JXB_CAT2(a, b)

namespace jxb {

// warps per CTA: the staging buffers of 8 warps no longer fit 227 KB of shared memory from 6 covariates on
constexpr int kWarps = (JXB_P <= 5) ? 8 : 4;

int JXB_CAT(k3_launch_solve_p, JXB_P)(const k3::ModelView& mv, int blocks, const float* rot, size_t ldc,
                                      int max_rows, const int32_t* n_rows_dev, const SolveParams& sp, double* out,
                                      int out_cols, int32_t* evals, int32_t* queue, cudaStream_t st) {
    constexpr int kSmem = kWarps * JXB_K3_BUFS * k3::WarpDims<JXB_P, true>::SMEM_DOUBLES * (int)sizeof(double);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k3::solve_warp_kernel<JXB_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        attr = true;
    }
    k3::solve_warp_kernel<JXB_P><<<blocks, kWarps * 32, kSmem, st>>>(mv, rot, ldc, max_rows, n_rows_dev, sp, out, out_cols,
                                                            evals, queue);
    return 0;
}

int JXB_CAT(k3_solve_blocks_per_sm_p, JXB_P)() {
    constexpr int kSmem = kWarps * JXB_K3_BUFS * k3::WarpDims<JXB_P, true>::SMEM_DOUBLES * (int)sizeof(double);
    int nb = 1;
    cudaFuncSetAttribute(k3::solve_warp_kernel<JXB_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k3::solve_warp_kernel<JXB_P>, kWarps * 32, kSmem);
    return nb < 1 ? 1 : nb;
}

int JXB_CAT(k3_launch_solve_thread_p, JXB_P)(const k3::ModelView& mv, const float* rotT, size_t ldr, int max_rows,
                                             const int32_t* n_rows_dev, const SolveParams& sp, double* out,
                                             int out_cols, int32_t* evals, const void* log_table, cudaStream_t st) {
    const int blocks = (max_rows + 127) / 128;
    constexpr int kSmem = 4 * (int)sizeof(k3::ThreadTile<JXB_P>);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k3::solve_thread_kernel<JXB_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        attr = true;
    }
    k3::solve_thread_kernel<JXB_P><<<blocks, 128, kSmem, st>>>(mv, rotT, ldr, max_rows, n_rows_dev, sp, out, out_cols,
                                                              evals, (const k3::LogTable*)log_table);
    return 0;
}

// fast_rcp: every s_i + lambda of the search interval is inside rcp_fast's range (checked by the caller)
int JXB_CAT(k3_launch_solve_lane_p, JXB_P)(const k3::ModelView& mv, int sms, const float* rot, size_t ldc, int max_rows,
                                           const int32_t* n_rows_dev, const SolveParams& sp, double* out, int out_cols,
                                           int32_t* evals, const void* log_table, double* ssq, int32_t* queue,
                                           int fast_rcp, const k3::PrefixTables* prefix, int build_tables, cudaStream_t st) {
    constexpr int kSmem = 4 * (int)sizeof(k3::ThreadTile<JXB_P, JXB_K3L_TILE>);
    constexpr int kSmemPrefix = 4 * (int)sizeof(k3::PrefixTile<JXB_P>);
    constexpr int kPerSm = (JXB_P <= 4) ? JXB_K3T_MINB : 3;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k3::solve_lane_kernel<JXB_P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        cudaFuncSetAttribute(k3::solve_lane_kernel<JXB_P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        cudaFuncSetAttribute(k3::prefix_eval_kernel<JXB_P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemPrefix);
        cudaFuncSetAttribute(k3::prefix_eval_kernel<JXB_P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemPrefix);
        attr = true;
    }
    k3::row_ssq_kernel<<<sms * 8, 256, 0, st>>>(rot, ldc, mv.n, max_rows, n_rows_dev, ssq);
    cudaMemsetAsync(queue, 0, sizeof(int32_t), st);
    const double* slots = nullptr;
    if (prefix) {
        // tables of the SNP-independent abscissae, then the leading evaluations of every SNP (k3_solve.cuh)
        slots = prefix->slots;
        cudaMemsetAsync(prefix->slots, 0xFF, (size_t)max_rows * k3::kPrefixEvals * 6 * sizeof(double), st);   // NaN = empty
        const int pblocks = (max_rows + 127) / 128;
        if (fast_rcp) {
            if (build_tables) k3::prefix_table_kernel<JXB_P, true><<<1, 128, 0, st>>>(mv, sp, (const k3::LogTable*)log_table, *prefix);
            k3::prefix_eval_kernel<JXB_P, true><<<pblocks, 128, kSmemPrefix, st>>>(mv, *prefix, rot, ldc, max_rows, n_rows_dev, sp, ssq);
        } else {
            if (build_tables) k3::prefix_table_kernel<JXB_P, false><<<1, 128, 0, st>>>(mv, sp, (const k3::LogTable*)log_table, *prefix);
            k3::prefix_eval_kernel<JXB_P, false><<<pblocks, 128, kSmemPrefix, st>>>(mv, *prefix, rot, ldc, max_rows, n_rows_dev, sp, ssq);
        }
    }
    const int blocks = std::min((max_rows + 127) / 128, sms * kPerSm);   // persistent: lanes refill from the queue
    if (fast_rcp)
        k3::solve_lane_kernel<JXB_P, true><<<blocks, 128, kSmem, st>>>(mv, rot, ldc, max_rows, n_rows_dev, sp, out, out_cols, evals,
                                                                      (const k3::LogTable*)log_table, ssq, queue, slots);
    else
        k3::solve_lane_kernel<JXB_P, false><<<blocks, 128, kSmem, st>>>(mv, rot, ldc, max_rows, n_rows_dev, sp, out, out_cols, evals,
                                                                       (const k3::LogTable*)log_table, ssq, queue, slots);
    return 0;
}

int JXB_CAT(k3_prefix_table_doubles_p, JXB_P)() { return 4 * k3::PrefixDims<JXB_P>::NS; }
int JXB_CAT(k3_prefix_rec_doubles_p, JXB_P)() { return k3::PrefixDims<JXB_P>::RSF; }

// Streamed (co-resident) lane kernel: 3 CTAs per SM, 16-sample tiles.  `sync` = {queue, ready, abort} (zeroed by the caller).
int JXB_CAT(k3_launch_solve_lane_stream_p, JXB_P)(const k3::ModelView& mv, int sms, const float* rot, size_t ldc,
                                                  int max_rows, const SolveParams& sp, double* out, int out_cols,
                                                  int32_t* evals, const void* log_table, const double* ssq,
                                                  int32_t* sync, cudaStream_t st) {
    constexpr int kSmem = 4 * (int)sizeof(k3::ThreadTile<JXB_P, 16>);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k3::solve_lane_stream_kernel<JXB_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        // same (maximum) shared-memory carve-out as the rotation kernel it shares SMs with: an SM cannot change its
        // L1/shared split while CTAs of another kernel are resident
        cudaFuncSetAttribute(k3::solve_lane_stream_kernel<JXB_P>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        attr = true;
    }
    const int blocks = std::min((max_rows + 127) / 128, sms * 3);
    k3::solve_lane_stream_kernel<JXB_P><<<blocks, 128, kSmem, st>>>(mv, rot, ldc, max_rows, sp, out, out_cols, evals,
                                                                   (const k3::LogTable*)log_table, ssq, sync);
    return 0;
}

// registers per thread and shared memory per CTA (static + dynamic) of the streamed kernel, for the co-residency check
int JXB_CAT(k3_solve_lane_stream_res_p, JXB_P)(int* regs, int* smem) {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, k3::solve_lane_stream_kernel<JXB_P>) != cudaSuccess) return -1;
    *regs = fa.numRegs;
    *smem = (int)fa.sharedSizeBytes + 4 * (int)sizeof(k3::ThreadTile<JXB_P, 16>);
    return 0;
}

// fixed-lambda scan of large batches: lane per SNP (frec = interleaved f64 records written by fixed_prepare_kernel)
int JXB_CAT(k3_launch_fixed_lane_p, JXB_P)(const k3::ModelView& mv, int sms, const double* frec, const double* scal,
                                           const float* rot, size_t ldc, int max_rows, const int32_t* n_rows_dev,
                                           int has_nullml, double nullml, double* out, int out_cols, cudaStream_t st) {
    constexpr int kSmem = 4 * (int)sizeof(k3::FixedTile<JXB_P>);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k3::fixed_lane_kernel<JXB_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        attr = true;
    }
    const int blocks = std::max(1, std::min((max_rows + 127) / 128, sms * 4));
    k3::fixed_lane_kernel<JXB_P><<<blocks, 128, kSmem, st>>>(mv, frec, scal, rot, ldc, max_rows, n_rows_dev, has_nullml, nullml,
                                                            out, out_cols);
    return 0;
}

int JXB_CAT(k3_launch_null_p, JXB_P)(const k3::ModelView& mv, int kind, double low, double high, int max_iter,
                                     double tol, int has_init, double init, double* out_dev, cudaStream_t st) {
    constexpr int kSmem = JXB_K3_BUFS * k3::WarpDims<JXB_P, false>::SMEM_DOUBLES * (int)sizeof(double);
    k3::null_warp_kernel<JXB_P><<<1, 32, kSmem, st>>>(mv, kind, low, high, max_iter, tol, has_init, init, out_dev);
    return 0;
}

}  // namespace jxb
