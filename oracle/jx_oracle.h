/*
 * jx_oracle.h -- CPU restatement of the JanusX exact-LMM scan (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle for the B200 path. It is NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it. The product (janusx_b200/) never links, imports or executes it.
 *
 * PARITY UNPINNED: the reference (a Rust cdylib) cannot be built in this image and its
 * own tests hold no expected beta/se/p/lambda values for this path.
 * The only reference-held known answers are checked in tests/test_oracle_kat.py
 * (decode codes, value LUT, chi-square tails, sanitize rules, toy-data determinism).
 *
 * Every function cites the reference file:line (relative to the JanusX tree) it restates.
 * Arithmetic follows Rust semantics: no fused multiply-add (build with -ffp-contract=off),
 * left-to-right evaluation, sequential i-order f64 sums.
 */
#ifndef JX_ORACLE_H
#define JX_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* genetic model codes: src/decode/decode.rs:99-146 */
enum { JXO_MODEL_ADD = 0, JXO_MODEL_DOM = 1, JXO_MODEL_REC = 2, JXO_MODEL_HET = 3 };

/* src/io/gfreader.rs:1378-1395 (full row) and :1453-1470 (selected samples).
 * sample_idx == NULL means identity over n_full samples. */
void jxo_count_row(const uint8_t *row, size_t n_full, const int64_t *sample_idx, size_t n_sel,
                   int64_t *missing, int64_t *het, int64_t *hom_alt);

/* src/stats/lmm.rs:1262-1323: QC closure. Returns 1 if kept. n = number of analysed samples. */
int jxo_qc_row(int64_t missing, int64_t het, int64_t hom_alt, size_t n,
               float maf_thr, float miss_thr, float het_thr,
               float *af, float *miss_rate);

/* counts + QC over a block of packed rows (OpenMP over rows). keep[r] in {0,1}. */
void jxo_count_qc_block(const uint8_t *packed, size_t bytes_per_snp, size_t rows, size_t n_full,
                        const int64_t *sample_idx, size_t n_sel,
                        float maf_thr, float miss_thr, float het_thr,
                        uint8_t *keep, float *af, float *miss_rate, int64_t *missing);

/* src/decode/decode.rs:163-271: decode + impute(2*af) + genetic model + mean-centre, f32.
 * row_indices (nullable) selects source rows inside `packed`. flip may be NULL (= all false). */
void jxo_decode_centered_block(const uint8_t *packed, size_t bytes_per_snp,
                               const int64_t *row_indices, size_t rows,
                               size_t n_full, const int64_t *sample_idx, size_t n,
                               const uint8_t *flip, const float *maf, int model,
                               float *out);

/* src/stats/lmm.rs:728-783 with the rotation arithmetic pinned:
 *   mode 0 = ROT_F64_F32STORE: f32-valued inputs, sequential-j f64 accumulate, round to f32.
 *   mode 1 = f32 multiply/accumulate, sequential j (a deterministic stand-in for SGEMM). */
void jxo_rotate_block(const float *g, size_t rows, size_t n, const float *ut, float *out, int mode);

/* src/stats/reml.rs:109-198 */
void jxo_rotate_xy(const float *ut, size_t n, const double *x, size_t q, const double *y,
                   double *x_rot, double *y_rot);

/* src/stats/reml.rs:255-362 / :364-470 / :472-568; snp may be NULL for the null model. */
double jxo_reml_loglike(double log10_lbd, const double *s, const double *xcov, const double *y,
                        const double *snp, size_t n, size_t p_cov);
double jxo_ml_loglike(double log10_lbd, const double *s, const double *xcov, const double *y,
                      const double *snp, size_t n, size_t p_cov);
void jxo_final_beta_se(double log10_lbd, const double *s, const double *xcov, const double *y,
                       const double *snp, size_t n, size_t p_cov, double out3[3]);

/* src/math/brent.rs:16-136. has_init=0 => midpoint start. f returns the cost (minimised). */
typedef double (*jxo_cost_fn)(double x, void *ctx);
void jxo_brent(jxo_cost_fn f, void *ctx, double low, double high, double tol, size_t max_iter,
               int has_init, double init_x, double *best_x, double *best_f, size_t *n_eval);

/* src/math/linalg.rs:2-17 */
double jxo_normal_sf(double z);
double jxo_chi2_sf_df1(double stat);

/* src/stats/reml.rs:570-616 -> (lambda, ml, reml) */
void jxo_reml_null(const double *s, const double *xcov, const double *y, size_t n, size_t p_cov,
                   double low, double high, size_t max_iter, double tol, double out3[3]);
/* src/stats/lmm.rs:2901-2924: null ML by Brent -> (log10 lambda, ml0). */
void jxo_ml_null(const double *s, const double *xcov, const double *y, size_t n, size_t p_cov,
                 double low, double high, size_t max_iter, double tol, int has_init, double init_x,
                 double out2[2]);

/* src/stats/lmm.rs:94-199: per-SNP REML scan on a rotated f32 block (no warm start carry).
 * has_nullml => out_cols = 4 else 3. has_init => every SNP seeded with init_log10_lbd.
 * n_eval (nullable): objective evaluations per SNP (REML Brent + 1 final [+1 ML]). */
void jxo_lmm_reml_block(const float *g_rot, size_t rows, size_t n,
                        const double *s, const double *xcov, const double *y, size_t p_cov,
                        double low, double high, double tol, size_t max_iter,
                        int has_init, double init_log10_lbd,
                        int has_nullml, double nullml,
                        double *out, int32_t *n_eval, int threads);

/* src/stats/lmm.rs:202-331: LMM2 rows = beta, se, pwald, lambda_reml, ml_alt, plrt. */
void jxo_lmm2_block(const float *g_rot, size_t rows, size_t n,
                    const double *s, const double *xcov, const double *y, size_t p_cov,
                    double low, double high, double tol, size_t max_iter,
                    int has_init_reml, double init_reml, int has_init_ml, double init_ml,
                    double nullml, double *out, int32_t *n_eval, int threads);

/* src/stats/fvlmm.rs:1484-1563 (cache) + :1691-1805 (block), with the two f32 GEMVs pinned to
 * sequential-i f64 accumulation rounded to f32. Returns 0 on success, <0 on the reference's
 * error conditions. */
typedef struct {
    size_t n, p;
    float *w;        /* n */
    float *py_tilde; /* n */
    float *wx_tilde; /* n*p */
    double *a_chol;  /* p*p */
    double ypy, log_det_v;
    int df;
} jxo_fixed_cache;
int jxo_fixed_cache_prepare(const double *s, const double *xcov, const double *y, size_t n, size_t p,
                            double lbd, jxo_fixed_cache *c);
void jxo_fixed_cache_free(jxo_fixed_cache *c);
void jxo_fixed_lambda_block(const float *g_rot, size_t rows, const jxo_fixed_cache *c,
                            int has_nullml, double nullml, double *out, int threads);

/* src/io/assoc2tsv.rs:430-517 + src/math/linalg.rs:99-108, 288-312: one TSV row.
 * miss is the missing RATE (f32). out_cols in {3,4,6}. Returns bytes written (excluding NUL). */
size_t jxo_format_row(char *buf, size_t cap, const char *chrom, int64_t pos, const char *snp,
                      const char *a0, const char *a1, float af, float miss_rate,
                      const double *row, int out_cols);

/* Rust `{:.Ne}` / `{:.N}` float formatting (used by jxo_format_row; exported for tests). */
size_t jxo_fmt_exp(char *buf, size_t cap, double v, int prec);
size_t jxo_fmt_fixed(char *buf, size_t cap, double v, int prec);

int jxo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
