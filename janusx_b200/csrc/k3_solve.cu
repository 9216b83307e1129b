// k3_solve.cu -- host launchers for the K3 kernels (device code: k3_solve.cuh; static covariate-count
// instantiations: k3_inst.cu compiled once per P; this unit holds the runtime-p and fixed-lambda kernels).
#include "k3_solve.cuh"

#include <cmath>
#include <cstring>

namespace jxb {

#define JXB_DECL_P(P)                                                                                          \
    int k3_launch_solve_p##P(const k3::ModelView&, int, const float*, size_t, int, const int32_t*,             \
                             const SolveParams&, double*, int, int32_t*, int32_t*, cudaStream_t);              \
    int k3_launch_null_p##P(const k3::ModelView&, int, double, double, int, double, int, double, double*,      \
                            cudaStream_t);                                                                     \
    int k3_solve_blocks_per_sm_p##P();                                                                         \
    int k3_launch_solve_thread_p##P(const k3::ModelView&, const float*, size_t, int, const int32_t*,           \
                                    const SolveParams&, double*, int, int32_t*, const void*, cudaStream_t);    \
    int k3_launch_solve_lane_p##P(const k3::ModelView&, int, const float*, size_t, int, const int32_t*,        \
                                  const SolveParams&, double*, int, int32_t*, const void*, double*, int32_t*,  \
                                  int, const k3::PrefixTables*, int, cudaStream_t);                            \
    int k3_prefix_table_doubles_p##P();                                                                        \
    int k3_prefix_rec_doubles_p##P();                                                                          \
    int k3_launch_solve_lane_stream_p##P(const k3::ModelView&, int, const float*, size_t, int, const SolveParams&, \
                                         double*, int, int32_t*, const void*, const double*, int32_t*, cudaStream_t); \
    int k3_solve_lane_stream_res_p##P(int*, int*);                                                             \
    int k3_launch_fixed_lane_p##P(const k3::ModelView&, int, const double*, const double*, const float*, size_t, int, \
                                  const int32_t*, int, double, double*, int, cudaStream_t);
JXB_DECL_P(1) JXB_DECL_P(2) JXB_DECL_P(3) JXB_DECL_P(4) JXB_DECL_P(5) JXB_DECL_P(6) JXB_DECL_P(7) JXB_DECL_P(8)
#undef JXB_DECL_P

size_t g_fixed_lane_min_rows = 4096;   // fixed-lambda batches at least this large use the lane-per-SNP kernel
int g_force_generic_divide = 0;   // tests: compare the rcp_fast kernels with the compiler-divide kernels
int g_prefix_evals = 1;           // lane-per-SNP solve: take the SNP-independent leading evaluations from shared tables (2 = any batch)
size_t g_prefix_min_rows = 2048;  // ... for batches at least this large (two extra launches)

namespace {

using namespace k3;

ModelView view_of(const Model& m) {
    ModelView v;
    v.s = m.s; v.y = m.y; v.xt = m.xt; v.ldn = m.ldn; v.n = (int)m.n; v.p = (int)m.p;
    v.rec = m.rec; v.rs = (int)m.rs;
    return v;
}

int sm_count(int device) {
    int v = 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
    return v;
}

}  // namespace

#define JXB_DISPATCH_P(P_, CALL_STATIC, CALL_DYN) \
    switch (P_) {                                 \
        case 1: { CALL_STATIC(1); break; }        \
        case 2: { CALL_STATIC(2); break; }        \
        case 3: { CALL_STATIC(3); break; }        \
        case 4: { CALL_STATIC(4); break; }        \
        case 5: { CALL_STATIC(5); break; }        \
        case 6: { CALL_STATIC(6); break; }        \
        case 7: { CALL_STATIC(7); break; }        \
        case 8: { CALL_STATIC(8); break; }        \
        default: { CALL_DYN(); break; }           \
    }

int launch_solve(const Model& m, const float* rot, size_t ldc, size_t max_rows, const int32_t* n_rows_dev,
                 const SolveParams& sp, double* out, int out_cols, int32_t* evals, int32_t* queue,
                 cudaStream_t st) {
    if (max_rows == 0) return 0;
    if (m.p < 1 || m.p > (size_t)kDynMaxCov) return fail(-2, "covariate columns must be in [1, 32]");
    const ModelView mv = view_of(m);
    if (m.p > 8) {
        // fallback: one thread per SNP
        const int blocks = (int)((max_rows + 63) / 64);
        solve_kernel<kDynMaxCov, true><<<blocks, 64, 0, st>>>(mv, rot, ldc, (int)max_rows, n_rows_dev, sp, out,
                                                              out_cols, evals);
        JXB_CUDA_OK(cudaGetLastError());
        return 0;
    }
    JXB_CUDA_OK(cudaMemsetAsync(queue, 0, sizeof(int32_t), st));
    int per_sm = 1;
#define B_STATIC(P) per_sm = k3_solve_blocks_per_sm_p##P()
#define B_DYN() per_sm = 1
    static int per_sm_cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (per_sm_cache[m.p] == 0) {
        JXB_DISPATCH_P((int)m.p, B_STATIC, B_DYN)
        per_sm_cache[m.p] = per_sm;
    }
    per_sm = per_sm_cache[m.p];
#undef B_STATIC
#undef B_DYN
    // persistent warps: fill every SM, never more warps than SNPs
    const int sms = sm_count(m.device);
    int blocks = (int)std::min<size_t>((size_t)sms * per_sm, (max_rows + 3) / 4);
    if (blocks < 1) blocks = 1;
#define S_STATIC(P) \
    k3_launch_solve_p##P(mv, blocks, rot, ldc, (int)max_rows, n_rows_dev, sp, out, out_cols, evals, queue, st)
#define S_DYN() (void)0
    JXB_DISPATCH_P((int)m.p, S_STATIC, S_DYN)
#undef S_STATIC
#undef S_DYN
    JXB_CUDA_OK(cudaGetLastError());
    return 0;
}

// 128-bucket table of the table-driven log (k3_solve.cuh table_log), built once per model
static int ensure_log_table(Model& m, cudaStream_t st) {
    if (!m.log_table) {
        // 128-bucket table: c = 1 + (i + 0.5)/128, 1/c rounded, ln c split hi/lo (long double on the host)
        LogTable h;
        for (int i = 0; i < 128; ++i) {
            const long double c = 1.0L + ((long double)i + 0.5L) / 128.0L;
            h.invc[i] = (double)(1.0L / c);
            // ln(c) must pair with the ROUNDED 1/c: m * invc - 1 = r  =>  ln m = ln(1+r) - ln(invc)
            const long double lc = -logl((long double)h.invc[i]);
            h.logc_hi[i] = (double)lc;
            h.logc_lo[i] = (double)(lc - (long double)h.logc_hi[i]);
        }
        JXB_CUDA_OK(cudaMalloc(&m.log_table, sizeof(LogTable)));
        JXB_CUDA_OK(cudaMemcpyAsync(m.log_table, &h, sizeof(LogTable), cudaMemcpyHostToDevice, st));
        JXB_CUDA_OK(cudaStreamSynchronize(st));
    }
    return 0;
}

// Large-batch variant: SNP-minor rotated block, one thread per SNP (p <= 8 only).
int launch_solve_thread(Model& m, const float* rotT, size_t ldr, size_t max_rows, const int32_t* n_rows_dev,
                        const SolveParams& sp, double* out, int out_cols, int32_t* evals, cudaStream_t st) {
    if (max_rows == 0) return 0;
    if (m.p < 1 || m.p > 8) return fail(-2, "thread-per-SNP solve supports 1..8 covariate columns");
    { int rc = ensure_log_table(m, st); if (rc) return rc; }
    const ModelView mv = view_of(m);
#define T_STATIC(P) k3_launch_solve_thread_p##P(mv, rotT, ldr, (int)max_rows, n_rows_dev, sp, out, out_cols, evals, m.log_table, st)
#define T_DYN() (void)0
    JXB_DISPATCH_P((int)m.p, T_STATIC, T_DYN)
#undef T_STATIC
#undef T_DYN
    JXB_CUDA_OK(cudaGetLastError());
    return 0;
}

// Large-batch variant, row-major rotated block: lane-per-SNP with refill (p <= 8 only).
int launch_solve_lane(Model& m, const float* rot, size_t ldc, size_t max_rows, const int32_t* n_rows_dev,
                      const SolveParams& sp, double* out, int out_cols, int32_t* evals, int32_t* queue, cudaStream_t st) {
    if (max_rows == 0) return 0;
    if (m.p < 1 || m.p > 8) return fail(-2, "lane-per-SNP solve supports 1..8 covariate columns");
    { int rc = ensure_log_table(m, st); if (rc) return rc; }
    if (m.ssq_cap < max_rows) {
        JXB_CUDA_OK(cudaStreamSynchronize(st));
        if (m.ssq) cudaFree(m.ssq);
        m.ssq = nullptr;
        JXB_CUDA_OK(cudaMalloc((void**)&m.ssq, std::max(max_rows, m.cap_rows) * sizeof(double)));
        m.ssq_cap = std::max(max_rows, m.cap_rows);
    }
    const ModelView mv = view_of(m);
    const int sms = sm_count(m.device);
    // rcp_fast (k3_solve.cuh) needs every s_i + lambda of the Brent interval to be positive and inside [1e-290, 1e290];
    // fl(s_i + lambda) is monotone in both, so the extreme eigenvalues and the interval ends decide
    int fast = 0;
    if (!m.s_host.empty() && !g_force_generic_divide) {
        const auto mm = std::minmax_element(m.s_host.begin(), m.s_host.end());
        const double lo = *mm.first + pow(10.0, std::min(sp.low, sp.high)), hi = *mm.second + pow(10.0, std::max(sp.low, sp.high));
        fast = std::isfinite(lo) && std::isfinite(hi) && lo >= 1e-290 && hi <= 1e290;
    }
    // shared-abscissa prefix (k3_solve.cuh prefix_eval_kernel): one allocation {xs[8] | sums | rec[n_pad][rsf] | slots}
    PrefixTables pt{};
    const PrefixTables* prefix = nullptr;
    int build_tables = 0;
    // (7 and 8 covariate columns: four abscissae x (p + 2) running sums no longer fit the register file; plain searches)
    if (m.p <= 6 && (g_prefix_evals == 2 || (g_prefix_evals && max_rows >= g_prefix_min_rows))) {
        const size_t n_pad = (m.n + 31) & ~(size_t)31;
        int sums_doubles = 0, rsf = 0;
#define PT_STATIC(P) sums_doubles = k3_prefix_table_doubles_p##P(), rsf = k3_prefix_rec_doubles_p##P()
#define PT_DYN() (void)0
        JXB_DISPATCH_P((int)m.p, PT_STATIC, PT_DYN)
#undef PT_STATIC
#undef PT_DYN
        const size_t head = 8 + (((size_t)sums_doubles + 7) & ~(size_t)7);
        const size_t rows_cap = std::max(max_rows, m.cap_rows);
        const size_t need = head + n_pad * rsf + rows_cap * kPrefixEvals * 6;
        if (m.prefix_cap < need) {
            JXB_CUDA_OK(cudaStreamSynchronize(st));
            if (m.prefix_buf) cudaFree(m.prefix_buf);
            m.prefix_buf = nullptr; m.prefix_cap = 0; m.prefix_valid = false;
            JXB_CUDA_OK(cudaMalloc((void**)&m.prefix_buf, need * sizeof(double)));
            m.prefix_cap = need;
        }
        pt.xs = m.prefix_buf;
        pt.sums = m.prefix_buf + 8;
        pt.rec = m.prefix_buf + head;
        pt.slots = pt.rec + n_pad * rsf;
        prefix = &pt;
        // the tables depend on the model and the search set-up only, not on the batch
        const double key[7] = {sp.low, sp.high, sp.tol, (double)sp.max_iter, (double)(sp.has_init != 0),
                               sp.has_init ? sp.init : 0.0, (double)fast};
        build_tables = !m.prefix_valid || memcmp(key, m.prefix_key, sizeof(key)) != 0;
        if (build_tables) { memcpy(m.prefix_key, key, sizeof(key)); m.prefix_valid = true; }
        note_launch(build_tables ? 2 : 1);   // [prefix_table_kernel +] prefix_eval_kernel
    }
#define L_STATIC(P) k3_launch_solve_lane_p##P(mv, sms, rot, ldc, (int)max_rows, n_rows_dev, sp, out, out_cols, evals, m.log_table, m.ssq, queue, fast, prefix, build_tables, st)
#define L_DYN() (void)0
    JXB_DISPATCH_P((int)m.p, L_STATIC, L_DYN)
#undef L_STATIC
#undef L_DYN
    JXB_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- streamed scan (cabi.cu scan_streamed): the solve runs while later row slabs are still being rotated ----------
int ensure_solve_lane_buffers(Model& m, size_t max_rows, cudaStream_t st) {
    { int rc = ensure_log_table(m, st); if (rc) return rc; }
    if (m.ssq_cap < max_rows) {
        JXB_CUDA_OK(cudaStreamSynchronize(st));
        if (m.ssq) cudaFree(m.ssq);
        m.ssq = nullptr;
        JXB_CUDA_OK(cudaMalloc((void**)&m.ssq, std::max(max_rows, m.cap_rows) * sizeof(double)));
        m.ssq_cap = std::max(max_rows, m.cap_rows);
    }
    return 0;
}

// sums of squares of rows [row0, row1) of the row-major rotated block, then publish `row1` rows as ready
int launch_row_ssq_publish(Model& m, const float* rot, size_t ldc, size_t row0, size_t row1, int32_t* sync, cudaStream_t st) {
    if (row1 > row0) {
        const int blocks = (int)std::min<size_t>((row1 - row0 + 7) / 8, (size_t)sm_count(m.device) * 8);
        row_ssq_kernel<<<blocks, 256, 0, st>>>(rot, ldc, (int)m.n, (int)row1, nullptr, m.ssq, (int)row0);
    }
    publish_ready_kernel<<<1, 1, 0, st>>>(sync, (int)row1);
    JXB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_solve_lane_stream(Model& m, const float* rot, size_t ldc, size_t rows, const SolveParams& sp, double* out,
                             int out_cols, int32_t* evals, int32_t* sync, cudaStream_t st) {
    if (rows == 0) return 0;
    if (m.p < 1 || m.p > 4) return fail(-2, "streamed lane solve supports 1..4 covariate columns");
    const ModelView mv = view_of(m);
    const int sms = sm_count(m.device);
    switch (m.p) {
        case 1: k3_launch_solve_lane_stream_p1(mv, sms, rot, ldc, (int)rows, sp, out, out_cols, evals, m.log_table, m.ssq, sync, st); break;
        case 2: k3_launch_solve_lane_stream_p2(mv, sms, rot, ldc, (int)rows, sp, out, out_cols, evals, m.log_table, m.ssq, sync, st); break;
        case 3: k3_launch_solve_lane_stream_p3(mv, sms, rot, ldc, (int)rows, sp, out, out_cols, evals, m.log_table, m.ssq, sync, st); break;
        default: k3_launch_solve_lane_stream_p4(mv, sms, rot, ldc, (int)rows, sp, out, out_cols, evals, m.log_table, m.ssq, sync, st); break;
    }
    JXB_CUDA_OK(cudaGetLastError());
    return 0;
}

int solve_lane_stream_resources(size_t p, int* regs, int* smem) {
    switch (p) {
        case 1: return k3_solve_lane_stream_res_p1(regs, smem);
        case 2: return k3_solve_lane_stream_res_p2(regs, smem);
        case 3: return k3_solve_lane_stream_res_p3(regs, smem);
        case 4: return k3_solve_lane_stream_res_p4(regs, smem);
        default: return -1;
    }
}

// rcp_fast against 1.0 / x on `count` pseudo-random doubles with exponents in [lo_exp, hi_exp]; returns mismatches
int rcp_selftest(size_t count, int lo_exp, int hi_exp, unsigned long long* mismatches_host) {
    unsigned long long* d = nullptr;
    JXB_CUDA_OK(cudaMalloc((void**)&d, sizeof(unsigned long long)));
    JXB_CUDA_OK(cudaMemset(d, 0, sizeof(unsigned long long)));
    const int threads = 256, blocks = 148 * 8, per = (int)std::max<size_t>(1, count / ((size_t)threads * blocks));
    rcp_selftest_kernel<<<blocks, threads>>>(0x1234567ull, per, (double)lo_exp, (double)hi_exp, d);
    cudaError_t e = cudaMemcpy(mismatches_host, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(-100, cudaGetErrorString(e));
    return 0;
}

int launch_null_fit(const Model& m, int kind, double low, double high, int max_iter, double tol, int has_init,
                    double init, double* out_dev, cudaStream_t st) {
    if (m.p < 1 || m.p > (size_t)kDynMaxCov) return fail(-2, "covariate columns must be in [1, 32]");
    const ModelView mv = view_of(m);
#define N_STATIC(P) k3_launch_null_p##P(mv, kind, low, high, max_iter, tol, has_init, init, out_dev, st)
#define N_DYN() null_kernel<kDynMaxCov, true><<<1, 32, 0, st>>>(mv, kind, low, high, max_iter, tol, has_init, init, out_dev)
    JXB_DISPATCH_P((int)m.p, N_STATIC, N_DYN)
#undef N_STATIC
#undef N_DYN
    JXB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_fixed_prepare(Model& m, double log10_lbd, cudaStream_t st) {
    if (m.p < 1 || m.p > (size_t)kDynMaxCov) return fail(-2, "covariate columns must be in [1, 32]");
    const ModelView mv = view_of(m);
    const double lbd = pow(10.0, log10_lbd);
    const int frs = (int)((m.p + 2 + 1) / 2 * 2);
    if (m.fx_rec) JXB_CUDA_OK(cudaMemsetAsync(m.fx_rec, 0, m.ldn * frs * sizeof(double), st));
    fixed_prepare_kernel<<<1, 32, 0, st>>>(mv, lbd, m.fx_w, m.fx_py, m.fx_wx, m.fx_scal, m.fx_rec, frs);
    JXB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_fixed_solve(const Model& m, const float* rot, size_t ldc, size_t max_rows, const int32_t* n_rows_dev,
                       int has_nullml, double nullml, double* out, int out_cols, cudaStream_t st) {
    if (max_rows == 0) return 0;
    if (m.p > 30) return fail(-2, "fixed-lambda scan supports at most 30 covariate columns");
    const ModelView mv = view_of(m);
    if (m.p <= 8 && m.fx_rec && max_rows >= g_fixed_lane_min_rows) {
        // large batches: lane per SNP, one HBM-bound pass over the rotated block (same ordered sums)
        const int sms = sm_count(m.device);
#define F_STATIC(P) k3_launch_fixed_lane_p##P(mv, sms, m.fx_rec, m.fx_scal, rot, ldc, (int)max_rows, n_rows_dev, has_nullml, nullml, out, out_cols, st)
#define F_DYN() (void)0
        JXB_DISPATCH_P((int)m.p, F_STATIC, F_DYN)
#undef F_STATIC
#undef F_DYN
        JXB_CUDA_OK(cudaGetLastError());
        return 0;
    }
    const int smem = 8 * 32 * 33 * (int)sizeof(double);
    static bool attr = false;
    if (!attr) {
        JXB_CUDA_OK(cudaFuncSetAttribute(fixed_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    const int sms = sm_count(m.device);
    int blocks = (int)std::min<size_t>((size_t)sms * 3, (max_rows + 7) / 8);
    if (blocks < 1) blocks = 1;
    fixed_solve_kernel<<<blocks, 256, smem, st>>>(mv, m.fx_w, m.fx_py, m.fx_wx, m.fx_scal, rot, ldc, (int)max_rows,
                                                  n_rows_dev, has_nullml, nullml, out, out_cols);
    JXB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace jxb
