// jxb_common.cuh -- shared declarations for the B200 (sm_100a) exact-LMM scan kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace jxb {

// thread-local last error message surfaced through jxb_last_error()
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
void note_launch(int k);  // counts kernel launches for jxb_launch_count()

#define JXB_CUDA_OK(expr)                                                                     \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            return ::jxb::fail(-100, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
        }                                                                                     \
    } while (0)

constexpr int kMaxCov = 16;        // covariate columns (incl. intercept) with a register-resident kernel
constexpr int kRotBM = 128;        // rotation CTA tile (SNP rows)
constexpr int kRotBN = 128;        // rotation CTA tile (eigen-directions)
constexpr int kRotBK = 16;         // doubles per k-slab = one 128-byte swizzle row
constexpr int kRotStages = 6;

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Null model resident in HBM (one per device).  Layouts (DESIGN.md "Data layout"):
//   s[n], y[n]            f64
//   xt[p][ldn]            f64, covariate-major (SoA) so a warp reads 32 consecutive samples
//   ut[n_pad][ldk]        f64, row k = k-th eigenvector (U^T row-major), zero padded:
//                         ldk = round_up(n,16) (TMA row pitch), n_pad = round_up(n,128)
struct Model {
    int device = 0;
    size_t n = 0, p = 0;
    size_t ldn = 0, ldk = 0, n_pad = 0;
    double* s = nullptr;
    double* y = nullptr;
    double* xt = nullptr;
    double* rec = nullptr;     // [round_up(n,32)][rs] interleaved per-sample records for K3 (s, y, x..)
    size_t rs = 0;
    double* ut = nullptr;      // may be null for rotated-input-only use
    bool owns = true;
    std::vector<double> s_host;  // host copy of S (records are rebuilt when Xcov / y change)
    // scan workspace (grown on demand)
    size_t cap_rows = 0;
    double* g64 = nullptr;     // [cap_rows_pad][ldk] decoded+centred genotypes (GEMM A operand)
    float* rot = nullptr;      // rotated block, f32 (reference storage type): row-major [cap_rows][ldc] for the warp
                               // solve kernel, or viewed SNP-minor [round_up(n,32)][cap_rows] for the thread kernel
    size_t ldc = 0;
    double* out = nullptr;     // [cap_rows][8]
    int32_t* evals = nullptr;  // [cap_rows]
    uint8_t* packed = nullptr; // [cap_rows][bps_cap]
    size_t bps_cap = 0;
    int32_t* counts = nullptr; // [cap_rows][4] missing, het, hom_alt, keep
    float* af = nullptr;       // [cap_rows]
    int32_t* src_row = nullptr;// [cap_rows] compacted -> source row
    int32_t* n_kept = nullptr; // device scalar (+ work counters): [0]=rows kept, [1]=solve queue, [2]=tile queue
    int64_t* sample_idx = nullptr; size_t n_sel_cap = 0;
    float* stage_f32 = nullptr; size_t stage_f32_cap = 0; // H2D staging for f32 inputs
    void* tmap_ut = nullptr;   // CUtensorMap for ut (host copy, 128 B)
    void* tmap_g = nullptr;    // CUtensorMap for g64
    cudaStream_t stream = nullptr;
    // int8-sliced rotation (k2_int8.cu)
    int8_t* q8 = nullptr; size_t ld8 = 0, q8_rows = 0;   // 7 digit planes [q8_rows][ld8]
    double* q8_inv_scale = nullptr; double* q8_rk = nullptr;
    int8_t* a8 = nullptr; size_t a8_rows = 0;            // 3 operand planes [a8_rows][ld8]: dosage, hom, missing
    double* coef = nullptr;                              // [a8_rows][4]
    int32_t* flags8 = nullptr;                           // [0] = any missing call in the batch
    int32_t* c32 = nullptr; size_t c32_elems = 0;        // int32 slice results (library variant)
    void* lt_ws = nullptr; size_t lt_ws_bytes = 0;
    double* corr64 = nullptr; size_t corr_rows = 0, ld_corr = 0;   // f64 correction terms (tcgen05 variant)
    void* tmap_a8 = nullptr; size_t tmap_a8_rows = 0; void* tmap_q8_7 = nullptr; void* tmap_q8_3 = nullptr;
    void* log_table = nullptr;                           // k3::LogTable (thread-per-SNP solve)
    double* ssq = nullptr; size_t ssq_cap = 0;           // per-row sum of squares (lane-per-SNP solve)
    double* prefix_buf = nullptr; size_t prefix_cap = 0; // tables + per-SNP slots of the shared-abscissa evaluations (doubles)
    bool prefix_valid = false; double prefix_key[7] = {0, 0, 0, 0, 0, 0, 0};   // tables in prefix_buf are current for this {low, high, tol, max_iter, has_init, init, divide}
    // fixed-lambda cache (A14)
    float* fx_w = nullptr; float* fx_py = nullptr; float* fx_wx = nullptr; double* fx_scal = nullptr;
    double* fx_rec = nullptr;                            // [ldn][round_up(p+2,2)] interleaved f64 records (p <= 8)
    double fx_log10_lbd = 0.0; bool fx_valid = false;
};

struct SolveParams {
    double low, high, tol;
    int max_iter;
    int has_init; double init;        // REML start (seed_with_init_guess)
    int has_nullml; double nullml;    // LMM: adds plrt column.  LMM2: required.
    int mode;                         // 0 = LMM (3|4 cols), 1 = LMM2 (6 cols)
};

// ---- kernels' host launchers (each file documents the reference lines it replaces) ----
int launch_count_qc(const Model& m, const uint8_t* packed, size_t bps, size_t rows, size_t n_full,
                    const int64_t* sample_idx, size_t n_sel, float maf_thr, float miss_thr, float het_thr,
                    int32_t* counts, float* af, float* miss_rate, cudaStream_t st);
int launch_compact(const int32_t* counts, size_t rows, int32_t* src_row, int32_t* n_kept, cudaStream_t st);
int launch_decode_center(const uint8_t* packed, size_t bps, const int32_t* src_row, const int32_t* n_kept,
                         size_t max_rows, size_t n_full, const int64_t* sample_idx, size_t n,
                         const float* af_by_src, const int32_t* counts_by_src, int model_code,
                         double* g64, size_t ldk, float* g32, size_t ld32, cudaStream_t st,
                         const float* row_lut = nullptr /* [rows][4] by source row and PLINK code: decode through it, no centring */);
int launch_widen_f32(const float* src, size_t ld_src, size_t rows, size_t n, double* dst, size_t ldk,
                     cudaStream_t st);
// out: transposed ? rotT-style [n][ld] : row-major [rows][ld]
int launch_rotate(Model& m, size_t max_rows, const int32_t* n_rows_dev, float* out, size_t ld, int transposed,
                  cudaStream_t st, int variant);
int launch_rotate_xy(const Model& m, const float* ut_f32, const double* x, size_t q, const double* y,
                     double* x_rot, double* y_rot, cudaStream_t st);
int launch_solve(const Model& m, const float* g_rot, size_t ldc, size_t max_rows, const int32_t* n_rows_dev,
                 const SolveParams& sp, double* out, int out_cols, int32_t* evals, int32_t* queue,
                 cudaStream_t st);
int launch_solve_thread(Model& m, const float* rotT, size_t ldr, size_t max_rows, const int32_t* n_rows_dev,
                        const SolveParams& sp, double* out, int out_cols, int32_t* evals, cudaStream_t st);
int launch_solve_lane(Model& m, const float* rot, size_t ldc, size_t max_rows, const int32_t* n_rows_dev,
                      const SolveParams& sp, double* out, int out_cols, int32_t* evals, int32_t* queue, cudaStream_t st);
int rcp_selftest(size_t count, int lo_exp, int hi_exp, unsigned long long* mismatches_host);
extern int g_force_generic_divide;
extern int g_prefix_evals;
extern size_t g_fixed_lane_min_rows;
// streamed scan: solve while later row slabs are still being rotated (k3_solve.cu / cabi.cu scan_streamed)
int ensure_solve_lane_buffers(Model& m, size_t max_rows, cudaStream_t st);
int launch_row_ssq_publish(Model& m, const float* rot, size_t ldc, size_t row0, size_t row1, int32_t* sync, cudaStream_t st);
int launch_solve_lane_stream(Model& m, const float* rot, size_t ldc, size_t rows, const SolveParams& sp, double* out,
                             int out_cols, int32_t* evals, int32_t* sync, cudaStream_t st);
int solve_lane_stream_resources(size_t p, int* regs_per_thread, int* smem_per_cta);
int launch_null_fit(const Model& m, int kind /*0 reml-null(3 out), 1 ml-null brent(2 out), 2 ml at x(1 out)*/,
                    double low, double high, int max_iter, double tol, int has_init, double init,
                    double* out_dev, cudaStream_t st);
int launch_fixed_prepare(Model& m, double log10_lbd, cudaStream_t st);
int launch_fixed_solve(const Model& m, const float* g_rot, size_t ldc, size_t max_rows,
                       const int32_t* n_rows_dev, int has_nullml, double nullml, double* out, int out_cols,
                       cudaStream_t st);
int make_tensor_maps(Model& m);
// int8-sliced exact rotation (k2_int8.cu)
int prepare_int8_slices(Model& m, cudaStream_t st);
int ensure_int8_workspace(Model& m, size_t rows_cap);
int launch_decode_int8(Model& m, const uint8_t* packed, size_t bps, const int32_t* src_row, const int32_t* n_kept,
                       size_t max_rows, size_t n_full, const int64_t* sample_idx, const float* af_by_src,
                       const int32_t* counts_by_src, int model_code, cudaStream_t st);
int launch_rotate_int8_lib(Model& m, size_t rows, bool has_missing, cudaStream_t st);
int launch_rotate_int8_tc(Model& m, size_t rows, bool has_missing, bool transposed_out, cudaStream_t st);   // k2_i8mma.cu
int prepare_rotate_tc(Model& m, size_t corr_rows);
int launch_rotate_int8_tc_slab(Model& m, size_t row0, size_t row1, bool has_missing, bool slim, cudaStream_t st);
int rotate_slim_resources(int* regs_per_thread, int* smem_per_cta);

}  // namespace jxb
