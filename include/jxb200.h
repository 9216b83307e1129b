/*
 * jxb200.h -- C ABI of libjxb200.so: the B200 (sm_100a) exact-LMM association scan.
 *
 * This is the drop-in boundary for JanusX's exact LMM path.  Each entry point replaces one PyO3
 * function of the reference extension module `janusx.janusx` (registered at src/lib.rs:911-941);
 * the reference symbol and its file:line are cited on every declaration.  INTEGRATION.md shows the
 * binding a JanusX maintainer would add (a Rust `extern "C"` block behind the existing #[pyfunction]
 * signatures, and the ctypes binding this repository ships in janusx_b200/_cabi.py).
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types.  All matrices row-major, dense unless an `ld`
 *     argument says otherwise.
 *   - return 0 on success, <0 on failure; jxb_last_error() returns a thread-local message.
 *     Per-SNP numerical failure is NOT an error: the row is NaN, NaN, 1.0[, ...] like the reference
 *     (src/stats/lmm.rs:74-91).
 *   - `*_host` arguments are host pointers (pageable or pinned); the library stages them.  Functions
 *     ending in `_dev` take device pointers on the model's device and do not synchronise the host
 *     unless stated.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef JXB200_H
#define JXB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jxb_model jxb_model; /* opaque: the null model resident in HBM + scan workspace */

/* genetic model: src/decode/decode.rs:99-146 */
enum { JXB_MODEL_ADD = 0, JXB_MODEL_DOM = 1, JXB_MODEL_REC = 2, JXB_MODEL_HET = 3 };

/* per-SNP optimiser settings: brent_minimize_with_init, src/math/brent.rs:16-136 */
typedef struct {
    double low, high;      /* bounds on log10(lambda) */
    double tol;            /* reference default 1e-2 */
    int32_t max_iter;      /* reference default 30 (CLI) / 50 (API) */
    int32_t has_init;      /* 1 => every SNP starts at init_log10_lbd (seed_with_init_guess) */
    double init_log10_lbd;
    int32_t has_nullml;    /* LMM: append plrt column.  LMM2: must be 1 (or fit with jxb_ml_null). */
    double nullml;
} jxb_solve_cfg;

/* QC thresholds of the unified BED scan: src/stats/lmm.rs:1262-1323 */
typedef struct {
    float maf_thr, miss_thr, het_thr;
    int32_t genetic_model; /* JXB_MODEL_* */
} jxb_qc_cfg;

const char* jxb_last_error(void);
int jxb_device_count(void);
/* Returns the library build tag ("jxb200 sm_100a ..."); used by the loader to prove the native path. */
const char* jxb_build_info(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
uint64_t jxb_launch_count(void);

/* ---- null model ---------------------------------------------------------------------------------
 * Holds S[n], Xcov[n,p], y_rot[n] (f64) and, when u_t is given, U^T[n,n] (f32 values, widened to a
 * K-padded f64 copy in HBM: 8*n*round_up(n,16) bytes).  Replaces the borrowed numpy views every
 * reference LMM function takes (s, xcov, y_rot, u_t), e.g. src/stats/lmm.rs:1479-1495. */
int jxb_model_create(int device, size_t n, size_t p, const double* s_host, const double* xcov_host,
                     const double* y_rot_host, const float* u_t_host /* nullable */, jxb_model** out);
/* Same, from device-resident buffers (e.g. after an NCCL broadcast done by the caller). u_t_dev is
 * f32[n,n] row-major, nullable.  The buffers are copied; the caller keeps ownership. */
int jxb_model_create_dev(int device, size_t n, size_t p, const double* s_dev, const double* xcov_dev,
                         const double* y_rot_dev, const float* u_t_dev, jxb_model** out);
void jxb_model_destroy(jxb_model* m);
/* Replace Xcov / y_rot (same n, p) without re-uploading U^T: a new trait on the same GRM. */
int jxb_model_set_xy(jxb_model* m, const double* xcov_host, const double* y_rot_host);
int jxb_model_sync(jxb_model* m);

/* lmm_rotate_x_y_with_ut_f64 -- src/stats/reml.rs:107-198.  x_host f64[n,q], y_host f64[n] ->
 * x_rot_host f64[n,q], y_rot_host f64[n].  Needs a model created with U^T (S/Xcov/y may be dummies). */
int jxb_rotate_xy(jxb_model* m, const double* x_host, size_t q, const double* y_host, double* x_rot_host,
                  double* y_rot_host);

/* lmm_reml_null_f32 -- src/stats/reml.rs:570-616.  out3 = (lambda, ml, reml). */
int jxb_reml_null(jxb_model* m, double low, double high, int max_iter, double tol, double out3[3]);
/* ml_loglike_null_f32 -- src/stats/reml.rs:618-646. */
int jxb_ml_loglike_null(jxb_model* m, double log10_lbd, double* ml);
/* null ML by Brent inside lmm_reml_lmm2_assoc_bed_to_tsv_f32 -- src/stats/lmm.rs:2901-2924.
 * out2 = (log10 lambda_ml, ml0). */
int jxb_ml_null(jxb_model* m, double low, double high, int max_iter, double tol, int has_init, double init,
                double out2[2]);

/* lmm_reml_chunk_f32 -- src/stats/lmm.rs:333-518.  g_rot_host f32[m,n] already rotated.
 * out_host f64[m, has_nullml ? 4 : 3] = beta, se, pwald[, plrt].  evals_host (nullable) i32[m]. */
int jxb_lmm_reml_chunk_f32(jxb_model* m, const float* g_rot_host, size_t rows, const jxb_solve_cfg* cfg,
                           double* out_host, int32_t* evals_host);
/* lmm_reml_chunk_from_snp_f32 -- src/stats/lmm.rs:1479-1630.  snp_host f32[m,n] centred genotypes. */
int jxb_lmm_reml_chunk_from_snp_f32(jxb_model* m, const float* snp_host, size_t rows, const jxb_solve_cfg* cfg,
                                    double* out_host, int32_t* evals_host);
/* lmm_reml_lmm2_chunk_from_snp_f32 -- src/stats/lmm.rs:1632-1780.  out_host f64[m,6] =
 * beta, se, pwald, lambda_reml, ml_alt, plrt.  `rotated` != 0 => snp_host is already rotated. */
int jxb_lmm2_chunk_f32(jxb_model* m, const float* snp_host, size_t rows, int rotated, const jxb_solve_cfg* cfg,
                       double* out_host, int32_t* evals_host);
/* lmm_assoc_chunk_f32 / lmm_assoc_chunk_from_snp_f32 / fvlmm_assoc_* (fixed lambda) --
 * src/stats/lmm.rs:2010-2238, src/stats/fvlmm.rs:1484-1563, 1691-1805.
 * out_host f64[m, nullml ? 4 : 3].  meta3 (nullable) = (ypy, log_det_v, df). */
int jxb_lmm_fixed_chunk_f32(jxb_model* m, const float* snp_host, size_t rows, int rotated, double log10_lbd,
                            const double* nullml /* nullable */, double* out_host, double meta3[3]);

/* Rotation only (rotate_snp_block_with_ut_blas, src/stats/lmm.rs:728-783): snp_host f32[m,n] ->
 * rot_host f32[m,n].  variant 0 = DMMA/TMA kernel, 1 = CUDA-core cross-check kernel. */
int jxb_rotate_block_f32(jxb_model* m, const float* snp_host, size_t rows, float* rot_host, int variant);

/* Packed scan: the producer+consumer body of run_unified_bed_scan_to_tsv_common
 * (src/stats/lmm.rs:1140-1432) for one batch of packed SNP rows:
 * count -> QC -> decode/impute/centre -> rotate -> per-SNP solve.
 *   packed_host u8[rows, bytes_per_snp] (PLINK SNP-major payload, 4 samples per byte)
 *   sample_idx_host nullable i64[n] (positions of the model's samples in FAM order); NULL = identity
 *   pre_keep_host nullable u8[rows]: extra host-side row mask (e.g. snps_only allele filter)
 *   mode 0 = LMM (3|4 cols), 1 = LMM2 (6 cols), 2 = fixed lambda at cfg->init_log10_lbd (3|4 cols)
 * Outputs (all host, indexed by SOURCE row unless stated):
 *   keep_host u8[rows], af_host f32[rows], missing_host i32[rows]
 *   out_host f64[n_kept, out_cols] compacted in SNP order; *n_kept_host = rows kept
 *   evals_host nullable i32[n_kept] */
int jxb_scan_packed(jxb_model* m, const uint8_t* packed_host, size_t bytes_per_snp, size_t rows, size_t n_full,
                    const int64_t* sample_idx_host, const uint8_t* pre_keep_host, const jxb_qc_cfg* qc,
                    const jxb_solve_cfg* cfg, int mode, uint8_t* keep_host, float* af_host, int32_t* missing_host,
                    double* out_host, int32_t* evals_host, size_t* n_kept_host);
/* Same with prepared row metadata (src/stats/lmm.rs:1237-1262, 3040-3187): row_af_host f32[rows] (nullable) replaces the
 * allele frequency counted on the device -- it is the imputation mean 2*row_maf the reference decodes with
 * (src/decode/decode.rs:213-219), whatever samples it was computed over, and the value returned in af_host;
 * row_flip_host u8[rows] (nullable) reverses the code LUT to [2, mean, 1, 0] (decode.rs:163-178).  The QC thresholds
 * in `qc` still apply to the device counts; callers with trusted metadata pass maf 0, miss 1, het 0. */
int jxb_scan_packed_prepared(jxb_model* m, const uint8_t* packed_host, size_t bytes_per_snp, size_t rows, size_t n_full,
                             const int64_t* sample_idx_host, const uint8_t* pre_keep_host, const float* row_af_host,
                             const uint8_t* row_flip_host, const jxb_qc_cfg* qc, const jxb_solve_cfg* cfg, int mode,
                             uint8_t* keep_host, float* af_host, int32_t* missing_host, double* out_host,
                             int32_t* evals_host, size_t* n_kept_host);
/* Double-buffered input staging -- the reference's producer/consumer double buffer (src/io/pipeline.rs:47-92) on the
 * device side: the library keeps two packed-row buffers in HBM and a copy stream.
 *   jxb_stage_packed       asynchronous H2D of the NEXT batch (pinned host memory for a truly asynchronous copy) into
 *                          the idle buffer; returns at once.  One batch may be staged at a time.
 *   jxb_scan_staged_begin  the staged batch becomes the current one (the compute stream waits for its copy; the buffers
 *                          swap), which frees the idle buffer for the next jxb_stage_packed.
 *   jxb_scan_staged        jxb_scan_packed_prepared on the current batch (no H2D of packed rows inside).
 * Loop: stage(0); for i: begin(); stage(i+1); scan_staged(i) -- batch i+1 goes up while batch i computes.
 *   jxb_stage_cancel       drops a staged / current batch (error paths). */
int jxb_stage_packed(jxb_model* m, const uint8_t* packed_host, size_t bytes_per_snp, size_t rows);
int jxb_scan_staged_begin(jxb_model* m);
int jxb_scan_staged(jxb_model* m, size_t n_full, const int64_t* sample_idx_host, const uint8_t* pre_keep_host,
                    const float* row_af_host, const uint8_t* row_flip_host, const jxb_qc_cfg* qc, const jxb_solve_cfg* cfg,
                    int mode, uint8_t* keep_host, float* af_host, int32_t* missing_host, double* out_host,
                    int32_t* evals_host, size_t* n_kept_host);
void jxb_stage_cancel(jxb_model* m);
/* Same with the packed batch already in HBM (bench.py `value`; multi-GPU shards).  Results stay on the
 * device in the model workspace; fetch with jxb_scan_fetch.  Asynchronous on the model's stream. */
int jxb_scan_packed_dev(jxb_model* m, const uint8_t* packed_dev, size_t bytes_per_snp, size_t rows, size_t n_full,
                        const int64_t* sample_idx_dev, const jxb_qc_cfg* qc, const jxb_solve_cfg* cfg, int mode);
int jxb_scan_fetch(jxb_model* m, size_t rows, int out_cols, uint8_t* keep_host, float* af_host,
                   int32_t* missing_host, double* out_host, int32_t* evals_host, size_t* n_kept_host);

/* Device-to-device variant for callers that keep results in HBM (multi-GPU gather over NCCL): out_dst_dev
 * f64[n_kept, out_cols] compacted, af_dst_dev f32[rows] and counts_dst_dev i32[rows, 4] = missing, het, hom_alt, keep
 * by source row (any may be NULL).  Synchronises the model's stream; *n_kept_host = rows kept. */
int jxb_scan_fetch_dev(jxb_model* m, size_t rows, int out_cols, double* out_dst_dev, float* af_dst_dev,
                       int32_t* counts_dst_dev, size_t* n_kept_host);

/* K1 alone, for parity tests: counts + QC + decode/centre of one packed batch.
 *   g_host nullable f32[n_kept, n] (compacted), counts_host i32[rows,4] = missing, het, hom_alt, keep */
int jxb_decode_packed(jxb_model* m, const uint8_t* packed_host, size_t bytes_per_snp, size_t rows, size_t n_full,
                      const int64_t* sample_idx_host, const jxb_qc_cfg* qc, int32_t* counts_host, float* af_host,
                      float* miss_rate_host, float* g_host, size_t* n_kept_host);
/* Decode with caller-supplied row decisions: keep_host u8[rows] selects the rows, af_host f32[rows] is the allele
 * frequency whose double is the imputed dosage.  Serves BedChunkReader.next_chunk_prepared (route B,
 * src/io/gfreader.rs:3580-3700), whose QC arithmetic (f64 rates, src/io/gfcore.rs:405-480) differs from the unified
 * scan's f32 expressions and is therefore evaluated by the caller from the exact integer counts of jxb_decode_packed.
 * g_host f32[n_kept, n] centred rows in source order. */
int jxb_decode_packed_prepared(jxb_model* m, const uint8_t* packed_host, size_t bps, size_t rows, size_t n_full,
                               const int64_t* sample_idx_host, const uint8_t* keep_host, const float* af_host,
                               int genetic_model, float* g_host, size_t* n_kept);

/* Decode through a caller-supplied value LUT per row (row_lut_host f32[rows, 4], indexed by the PLINK code 00 / 01 =
 * missing / 10 / 11) with NO centring; every supplied row is decoded.  g_host f32[rows, n].  Serves
 * BedChunkReaderFromMeta.next_chunk_prepared (src/io/gfreader.rs:7623-7730: [(0 - mean), 0, (1 - mean), (2 - mean)], mean =
 * 2 * ALT frequency of the shared per-trait metadata) and the raw BedChunkReader.next_chunk (src/io/gfreader.rs:3319-3440,
 * process_snp_row src/io/gfcore.rs:405-480: [0, imputed, 1, 2], reversed when the row is flipped to the minor allele). */
int jxb_decode_packed_lut(jxb_model* m, const uint8_t* packed_host, size_t bps, size_t rows, size_t n_full,
                          const int64_t* sample_idx_host, const float* row_lut_host, float* g_host);

/* Debug: rows [row0, row0 + rows) of the rotated block of the last scan, INCLUDING the zero padding of every row
 * (rot_host f32[rows, round_up(n, 32)], compacted row order).  Parity tests compare it with the oracle's rotation. */
int jxb_debug_fetch_rot(jxb_model* m, size_t row0, size_t rows, float* rot_host);

/* Per-stage device timers of the last jxb_scan_packed* call, milliseconds:
 * [0]=count+qc+compact [1]=decode [2]=rotate [3]=solve [4]=h2d [5]=d2h.  (The reference's
 * JX_LMM_UNIFIED_STAGE_TIMING, src/stats/lmm.rs:2711-2742.)  Enabled by jxb_set_timing(1). */
void jxb_set_timing(int on);
int jxb_last_stage_ms(jxb_model* m, float ms6[6]);
/* Measured FP64 CUDA-core issue rate of `device`, in 1e12 lane-instructions per second: tops3 = {DFMA, DADD, DMUL}
 * (register-resident chains; tools/fp64_probe.cu is the stand-alone version).  The solve kernels are compiled without
 * FMA contraction (reference rounding), so their ceiling is the DADD/DMUL rate; bench.py reports against it. */
int jxb_fp64_probe(int device, double tops3[3]);
/* raw stream handle (cudaStream_t) the model launches on, for callers timing with CUDA events */
void* jxb_model_stream(jxb_model* m);
/* rotation kernel for subsequent packed additive scans: 3 = hand-written tcgen05 int8-sliced exact rotation
 * (default), 2 = same arithmetic with cuBLASLt slice GEMMs, 0 = FP64 DMMA/TMA GEMM, 1 = CUDA-core cross-check.
 * Chunk entry points taking arbitrary f32 genotypes always use the FP64 GEMM. */
void jxb_set_rotate_variant(int variant);
/* batches with at least this many kept SNPs use the one-thread-per-SNP solve kernel (default 32768);
 * smaller ones the warp-per-SNP kernel.  Both reproduce the reference summation order. */
void jxb_set_thread_solve_min_rows(size_t rows);
/* Kernel for those large batches: 0 (default) = lane-per-SNP with refill on the row-major block, 1 = the earlier
 * thread-per-SNP kernel on an SNP-minor block (tcgen05 rotation only).  Same per-SNP arithmetic, identical results. */
void jxb_set_big_solve_kernel(int variant);
/* The large-batch solve kernels take their per-sample 1/(s_i + lambda) from a branch-free copy of the compiler's own f64
 * reciprocal sequence whenever every s_i + lambda of the search interval lies in [1e-290, 1e290] (bit-identical
 * results; no divide slow path in the sample loops).  jxb_set_generic_divide(1) forces the compiler-generated divide
 * (tests compare the two); jxb_selftest_rcp compares `count` pseudo-random values with binary exponents in
 * [lo_exp, hi_exp] (extreme mantissas included) and returns the number of differing bit patterns. */
void jxb_set_generic_divide(int on);
/* The first three abscissae of every per-SNP REML search (src/math/brent.rs:16-136 started at the same point of the same
 * interval: x0, the golden-section step, then one of two golden-section steps) do not depend on the SNP.  The lane-per-SNP
 * solve therefore evaluates them for the whole batch ahead of the searches, taking 1/(s_i + lambda), the covariate block of
 * Z'V^-1 Z, Z'V^-1 y and sum ln v from per-batch tables; only the SNP column's sums are formed per SNP.  Same operations
 * in the same order, bit-identical results and evaluation counts (up to 6 covariate columns; 7 and 8 run plain
 * searches).  1 (default) = batches of at least 2048 kept SNPs,
 * 2 = every batch the lane-per-SNP kernel handles, 0 = off (tests compare). */
void jxb_set_prefix_evals(int on);
/* fixed-lambda batches with at least this many rows use the lane-per-SNP kernel (one HBM-bound pass over the rotated
 * block; default 4096); smaller ones the warp-per-SNP kernel.  Same ordered sums, identical results. */
void jxb_set_fixed_lane_min_rows(size_t rows);
int jxb_selftest_rcp(size_t count, int lo_exp, int hi_exp, uint64_t* mismatches);
/* Streamed scan (default on): large additive LMM / LMM2 batches are rotated in slabs of `slab_rows` rows (0 = keep the
 * current value, default 8192) while ONE persistent solve kernel consumes the rows already rotated -- the tensor pipe
 * (rotation) and the FP64 pipe (solve) of every SM work at the same time.  Same kernels' arithmetic, identical results.
 * on = 0 restores rotate-then-solve. */
void jxb_set_stream_overlap(int on, size_t slab_rows);
/* jxb_last_stage_ms plus [6] = duration of the solve kernel on its own stream, [7] = 1 when the last scan was
 * streamed (then [2] is the rotation of all slabs and [3] the part of the solve left after the last slab). */
int jxb_last_stage_ms8(jxb_model* m, float ms8[8]);

/* ---- file level ------------------------------------------------------------------------------------
 * lmm_reml_assoc_bed_to_tsv_f32 / lmm_reml_lmm2_assoc_bed_to_tsv_f32 / fvlmm_assoc_bed_to_tsv_f32 --
 * src/stats/lmm.rs:2488-2750, 2753-3038; src/stats/fvlmm.rs:2482-2527.  Streams prefix.bed/.bim/.fam,
 * writes the reference TSV schema (src/io/assoc2tsv.rs:45-57, 430-517), SNP order preserved.
 * Warm start across SNPs (schedule-dependent in the reference, src/stats/lmm.rs:134-161) is never
 * used: rows equal the reference run with JX_LMM_UNIFIED_NO_WARM_START=1. */
typedef int (*jxb_progress_cb)(size_t done, size_t total, void* user); /* nonzero return aborts */
typedef struct {
    const char* bed_prefix;
    const char* out_tsv;
    jxb_qc_cfg qc;
    jxb_solve_cfg solve;
    int32_t mode;              /* 0 LMM, 1 LMM2, 2 fixed lambda (solve.init_log10_lbd) */
    int32_t snps_only;
    const char* const* sample_ids; /* nullable; n entries (IIDs) */
    size_t n_sample_ids;
    size_t batch_rows;         /* SNP rows per device batch (reference: rotate_block_rows) */
    size_t snp_begin, snp_end; /* scan [begin,end) of BED rows; end==0 => all (multi-GPU shards) */
    int32_t write_header;
    size_t progress_every;
    /* prepared row metadata (src/stats/lmm.rs:2576-2612): nullable ascending BED row indices.  When given, only
     * these rows are scanned and the QC thresholds are NOT re-applied (the list is trusted, like the reference);
     * allele frequency and missingness are recomputed from the packed rows on the device. */
    const int64_t* row_indices;
    size_t n_row_indices;
    /* the rest of the prepared metadata, each nullable, n_row_indices entries parallel to row_indices: row_maf = the
     * imputation frequency and the TSV af column; row_flip reverses the code LUT; row_missing = missing RATE, turned
     * into a count by round(rate * n) and back into the TSV miss column (src/stats/lmm.rs:1934-1950, 2576-2612) */
    const float* row_maf;
    const uint8_t* row_flip;
    const float* row_missing;
    /* host staging (src/io/pipeline.rs:47-92, src/io/gload.rs:523-800): 0 = default.  The BED payload is copied in
     * windows of at most this many MiB through a pinned double buffer, batch i+1 going up while batch i computes. */
    size_t mmap_window_mb;
} jxb_bed_scan_cfg;
int jxb_scan_bed_to_tsv(jxb_model* m, const jxb_bed_scan_cfg* cfg, size_t* rows_written, jxb_progress_cb cb,
                        void* user);

/* TSV row formatter (append_assoc_row_from_fields, src/io/assoc2tsv.rs:430-517); exported for tests.
 * Returns bytes written.  Never writes past buf[cap-1]: when cap is below the worst case for these strings
 * (2*strlen(chrom) + strlen(snp) + strlen(a0) + strlen(a1) + 3088) nothing useful is written and the return value is
 * that size (> cap) -- call again with a buffer at least that large.  (Rust's String has no row length limit.) */
size_t jxb_format_row(char* buf, size_t cap, const char* chrom, int64_t pos, const char* snp, const char* a0,
                      const char* a1, float af, float miss_rate, const double* row, int out_cols);

/* Block formatter behind GwasAssocTsvWriter.write_chunk (src/io/assoc2tsv.rs:789-860, append_assoc_row_text
 * :364-428): `rows` result rows; every string column is one blob of `rows` NUL-terminated strings; genetic_model
 * 0..3 = add/dom/rec/het applies transform_alleles_by_model (:117-137).  Returns the bytes the block needs; nothing
 * beyond `cap` is written, so a return value > cap means "call again with a larger buffer".  0 = bad arguments. */
size_t jxb_format_block(char* buf, size_t cap, size_t rows, const char* chrom, const int64_t* pos, const char* snp,
                        const char* a0, const char* a1, const float* af, const float* miss_rate, const double* res,
                        int out_cols, int genetic_model);

/* Position-dependent checksum of a host buffer, computed at memory bandwidth on up to 8 threads (cache key of the
 * resident U^T in the Python front end). */
void jxb_host_checksum(const void* data, size_t bytes, uint64_t out2[2]);

/* The writer's number formatters (exact fast path for `{:.N}` / `{:.Ne}`) against their printf route on `count` values
 * chosen to stress them (ties, near-ties, powers of ten, carries); returns the number of differing strings. */
size_t jxb_selftest_format(size_t count, uint64_t seed, int prec, char* first_bad, size_t bad_cap);

/* Header line for 3 / 4 / 6 result columns (AssocResultCols::header, src/io/assoc2tsv.rs:45-57); NULL otherwise. */
const char* jxb_tsv_header(int out_cols);

/* ---- SURVEY 8(f) "next" rows: the two steps in front of the scan ------------------------------------------------
 * N1  GRM: grm_packed_f32 / grm_packed_f64, method 1 = centred additive (src/stats/grm.rs:204-608, 3053-3623;
 *     decode_additive_grm_block_f32, src/decode/decode.rs:728-900).  K = Z Z^T / sum_s 2p(1-p), z = code LUT
 *     {0-mu, 0 (missing), 1-mu, 2-mu}, mu = 2*clamp(row_maf,0,1).  The contraction runs as exact int8 digit-plane
 *     MMAs (mu on a 2^-21 grid; csrc/grm.cu), so K does not depend on the batch split or the summation order. */
typedef struct jxb_grm jxb_grm; /* opaque: the n x n accumulator resident in HBM + batch workspace */
/* sample_idx_host: nullable (all n_full samples, FAM order); otherwise n_sel positions into the packed rows. */
int jxb_grm_create(int device, size_t n_full, const int64_t* sample_idx_host, size_t n_sel, int method, jxb_grm** out);
/* Add `rows` packed SNP rows ([rows][bps], bps = ceil(n_full/4)).  row_maf_host f32[rows] is the prepared allele
 * frequency the reference takes; NULL => computed on the device over the selected samples (A3 formula).  Every
 * supplied row enters the GRM unless `qc` says otherwise.  May be called repeatedly; rows are batched internally. */
int jxb_grm_update(jxb_grm* g, const uint8_t* packed_host, size_t bps, size_t rows, const float* row_maf_host,
                   const jxb_qc_cfg* qc /* nullable; only with row_maf_host == NULL: rows failing the A3 thresholds
                                           (src/stats/lmm.rs:1262-1323) are left out */);
/* Same with device-resident inputs on the handle's device (the caller orders its own producer before the call;
 * the call synchronises the handle's stream before returning). */
int jxb_grm_update_dev(jxb_grm* g, const uint8_t* packed_dev, size_t bps, size_t rows, const float* row_maf_dev,
                       const jxb_qc_cfg* qc);
size_t jxb_grm_rows_used(jxb_grm* g);      /* SNP rows that entered the GRM so far */
/* Scale by 1/sum 2p(1-p) and mirror (grm_scale_and_symmetrize_raw_f64, grm.rs:2771-2786).  k_host f64[n,n]
 * nullable (leave the matrix on the device for jxb_eigh_dev); varsum_out nullable. */
int jxb_grm_finish(jxb_grm* g, double* k_host, double* varsum_out);
double* jxb_grm_device_matrix(jxb_grm* g); /* f64[n,n] on the handle's device; valid until jxb_grm_destroy */
void* jxb_grm_stream(jxb_grm* g);          /* the cudaStream_t the handle's work is ordered on */
void jxb_grm_destroy(jxb_grm* g);

/* N3  VCF(.gz) -> PLINK BED/BIM/FAM, the one-off conversion in front of a `-vcf` scan (VcfSnpIter::next_snp_raw,
 *     src/io/gfcore.rs:2875-2980; plink2bits_from_g_f32, src/io/gfreader.rs:2630-2641).  GT strings 0/0 0|0 -> 00,
 *     0/1 1/0 0|1 1|0 -> 10, 1/1 1|1 -> 11, anything else -> 01 (missing); dosage counts ALT, BIM col 5 = REF, col 6 =
 *     ALT; ID "." or empty -> chrom_pos.  snps_only drops sites whose REF/ALT are not single A/C/G/T.  Host only. */
int jxb_vcf_to_plink(const char* vcf_path, const char* out_prefix, int snps_only, size_t* n_samples, size_t* n_sites);

/* N2  eigendecomposition: rust_eigh_from_array_f64[_inplace] (src/math/eigh.rs:1621-1705, 1883-1990).  One
 *     cuSOLVER call (cusolverDnXsyevd, dlopen).  a f64[n,n] symmetric; diag_shift is added to the diagonal first (the
 *     reference's 1e-6 ridge, workflow_model_stream.py:902).  Eigenvalues ascending; the matrix is returned as U^T
 *     row-major (row k = k-th eigenvector = numpy.linalg.eigh(a)[1].T).  ut_f32_* receive the f32-rounded U^T the
 *     scan consumes (jxb_model_create[_dev]). */
/* Devices for the eigendecomposition: 0 / 1 (default) = one cusolverDnXsyevd call on the caller's device; k > 1 =
 * cusolverMgSyevd over the first k visible devices (1-D block-cyclic column panels, peer copies over NVLink).
 * n > 46,340 always takes the cusolverMg path (cusolverDnXsyevd rejects n*n >= 2^31), on one device unless k says more. */
void jxb_set_eigh_devices(int n_devices);
int jxb_eigh(int device, size_t n, const double* a_host, double diag_shift, double* evals_host,
             double* ut_host /* nullable */, float* ut_f32_host /* nullable */);
/* In place on the device: a_dev is overwritten by U^T.  Synchronises `stream` once (convergence flag); the
 * ut_f32_dev conversion is enqueued on `stream` after that. */
int jxb_eigh_dev(int device, size_t n, double* a_dev, double diag_shift, double* evals_dev,
                 float* ut_f32_dev /* nullable */, void* stream);
/* N1 -> N2 chained on the device (the f64 matrix never visits the host): jxb_grm_finish, then K + diag_shift*I is
 * decomposed in place (the handle's matrix becomes U^T; no further updates).  evals_host f64[n], ut_f32_host f32[n,n]. */
int jxb_grm_eigh(jxb_grm* g, double diag_shift, double* evals_host, float* ut_f32_host);

#ifdef __cplusplus
}
#endif
#endif
