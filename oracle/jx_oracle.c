/*
 * jx_oracle.c -- CPU restatement of the JanusX exact-LMM scan.  TEST INFRASTRUCTURE ONLY
 * (see jx_oracle.h).  PARITY UNPINNED: the reference holds no expected values for this path.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared (see oracle/Makefile).
 * -ffp-contract=off matters: Rust never fuses a*b+c, so neither may this file.
 */
#include "jx_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define JXO_MAX_DIM 64

int jxo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * A2: genotype counts.  PLINK 2-bit codes (src/math/bedmath.rs:20-27): 00 -> 0, 01 -> missing,
 * 10 -> het, 11 -> hom alt; sample j sits in bits 2*(j&3) of byte j>>2.
 * ---------------------------------------------------------------------------------------- */
void jxo_count_row(const uint8_t *row, size_t n_full, const int64_t *sample_idx, size_t n_sel,
                   int64_t *missing, int64_t *het, int64_t *hom_alt) {
    int64_t m = 0, h = 0, a = 0;
    if (sample_idx == NULL) {
        /* src/io/gfreader.rs:1378-1395 (the popcount kernels count the same three codes) */
        for (size_t j = 0; j < n_full; ++j) {
            unsigned code = (row[j >> 2] >> ((j & 3) * 2)) & 3u;
            if (code == 1u) ++m;
            else if (code == 2u) ++h;
            else if (code == 3u) ++a;
        }
    } else {
        /* src/io/gfreader.rs:1453-1470 */
        for (size_t k = 0; k < n_sel; ++k) {
            size_t sid = (size_t)sample_idx[k];
            unsigned code = (row[sid >> 2] >> ((sid & 3) * 2)) & 3u;
            if (code == 1u) ++m;
            else if (code == 2u) ++h;
            else if (code == 3u) ++a;
        }
    }
    *missing = m;
    *het = h;
    *hom_alt = a;
}

/* A3: src/stats/lmm.rs:1262-1323 */
int jxo_qc_row(int64_t missing, int64_t het, int64_t hom_alt, size_t n,
               float maf_thr, float miss_thr, float het_thr, float *af, float *miss_rate) {
    int64_t non_missing = (int64_t)n - missing;
    if (non_missing < 0) non_missing = 0;
    float mr = (n > 0) ? ((float)missing / (float)n) : 1.0f;
    *miss_rate = mr;
    *af = 0.0f;
    if (mr > miss_thr) return 0;
    if (non_missing == 0) {
        return (maf_thr > 0.0f) ? 0 : 1;
    }
    if (het_thr > 0.0f) {
        float het_rate = (float)het / (float)non_missing;
        if (het_rate > het_thr) return 0;
    }
    int64_t alt_sum = het + 2 * hom_alt;
    float alt_freq = (float)alt_sum / (2.0f * (float)non_missing);
    float other = 1.0f - alt_freq;
    float maf_v = (alt_freq < other) ? alt_freq : other; /* f32::min */
    if (maf_v < maf_thr) return 0;
    *af = alt_freq;
    return 1;
}

void jxo_count_qc_block(const uint8_t *packed, size_t bytes_per_snp, size_t rows, size_t n_full,
                        const int64_t *sample_idx, size_t n_sel,
                        float maf_thr, float miss_thr, float het_thr,
                        uint8_t *keep, float *af, float *miss_rate, int64_t *missing) {
    size_t n = sample_idx ? n_sel : n_full;
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)rows; ++r) {
        int64_t m, h, a;
        jxo_count_row(packed + (size_t)r * bytes_per_snp, n_full, sample_idx, n_sel, &m, &h, &a);
        missing[r] = m;
        keep[r] = (uint8_t)jxo_qc_row(m, h, a, n, maf_thr, miss_thr, het_thr, &af[r], &miss_rate[r]);
    }
}

/* src/decode/decode.rs:121-145 */
static double model_apply(int model, double g) {
    switch (model) {
    case JXO_MODEL_DOM: return (g > 0.0) ? 1.0 : 0.0;
    case JXO_MODEL_REC: return (fabs(g - 2.0) < 1e-6) ? 1.0 : 0.0;
    case JXO_MODEL_HET: return (fabs(g - 1.0) < 1e-6) ? 1.0 : 0.0;
    default: return g;
    }
}

/* A4: src/decode/decode.rs:163-271 (identity, dense-subset and gather-plan paths all yield the
 * same values: LUT by code, then subtract the f32 mean of the selected samples). */
void jxo_decode_centered_block(const uint8_t *packed, size_t bytes_per_snp,
                               const int64_t *row_indices, size_t rows,
                               size_t n_full, const int64_t *sample_idx, size_t n,
                               const uint8_t *flip, const float *maf, int model, float *out) {
    (void)n_full;
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)rows; ++r) {
        size_t src = row_indices ? (size_t)row_indices[r] : (size_t)r;
        const uint8_t *row = packed + src * bytes_per_snp;
        /* decode.rs:218: mean_g = (2.0 * maf as f64).max(0.0) as f32 */
        double mg64 = 2.0 * (double)maf[r];
        if (!(mg64 > 0.0)) mg64 = 0.0; /* f64::max(NaN,0)=0 */
        float mean_g = (float)mg64;
        float raw[4];
        if (flip && flip[r]) {
            raw[0] = 2.0f; raw[1] = mean_g; raw[2] = 1.0f; raw[3] = 0.0f;
        } else {
            raw[0] = 0.0f; raw[1] = mean_g; raw[2] = 1.0f; raw[3] = 2.0f;
        }
        float lut[4];
        for (int c = 0; c < 4; ++c) lut[c] = (float)model_apply(model, (double)raw[c]);
        float *dst = out + (size_t)r * n;
        double sum = 0.0;
        for (size_t j = 0; j < n; ++j) {
            size_t sid = sample_idx ? (size_t)sample_idx[j] : j;
            unsigned code = (row[sid >> 2] >> ((sid & 3) * 2)) & 3u;
            float v = lut[code];
            dst[j] = v;
            sum += (double)v; /* decode.rs:185 sequential f64 sum */
        }
        if (n > 0) {
            float mean = (float)(sum / (double)n);
            for (size_t j = 0; j < n; ++j) dst[j] = dst[j] - mean;
        }
    }
}

/* A5 */
void jxo_rotate_block(const float *g, size_t rows, size_t n, const float *ut, float *out, int mode) {
#pragma omp parallel for schedule(dynamic, 4)
    for (long r = 0; r < (long)rows; ++r) {
        const float *gr = g + (size_t)r * n;
        float *o = out + (size_t)r * n;
        for (size_t k = 0; k < n; ++k) {
            const float *u = ut + k * n;
            if (mode == 0) {
                double acc = 0.0;
                for (size_t j = 0; j < n; ++j) acc += (double)gr[j] * (double)u[j];
                o[k] = (float)acc;
            } else {
                float acc = 0.0f;
                for (size_t j = 0; j < n; ++j) acc += gr[j] * u[j];
                o[k] = acc;
            }
        }
    }
}

/* A6: src/stats/reml.rs:158-171 */
void jxo_rotate_xy(const float *ut, size_t n, const double *x, size_t q, const double *y,
                   double *x_rot, double *y_rot) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; ++i) {
        const float *u = ut + (size_t)i * n;
        for (size_t c = 0; c < q; ++c) {
            double acc = 0.0;
            for (size_t j = 0; j < n; ++j) acc += ((double)u[j]) * x[j * q + c];
            x_rot[(size_t)i * q + c] = acc;
        }
        double accy = 0.0;
        for (size_t j = 0; j < n; ++j) accy += ((double)u[j]) * y[j];
        y_rot[i] = accy;
    }
}

/* ------------------------------------------------------------------------------------------
 * A13: src/math/linalg.rs:314-367, src/stats/reml.rs:46-66
 * ---------------------------------------------------------------------------------------- */
static int cholesky_inplace(double *a, size_t dim) {
    for (size_t i = 0; i < dim; ++i) {
        for (size_t j = 0; j <= i; ++j) {
            double sum = a[i * dim + j];
            for (size_t k = 0; k < j; ++k) sum -= a[i * dim + k] * a[j * dim + k];
            if (i == j) {
                if (sum <= 1e-18) return 0;
                a[i * dim + j] = sqrt(sum);
            } else {
                a[i * dim + j] = sum / a[j * dim + j];
            }
        }
        for (size_t j = i + 1; j < dim; ++j) a[i * dim + j] = 0.0;
    }
    return 1;
}

static void cholesky_solve(const double *a, size_t dim, const double *b, double *x) {
    double yv[JXO_MAX_DIM];
    for (size_t i = 0; i < dim; ++i) {
        double sum = b[i];
        for (size_t k = 0; k < i; ++k) sum -= a[i * dim + k] * yv[k];
        yv[i] = sum / a[i * dim + i];
    }
    for (size_t ii = 0; ii < dim; ++ii) {
        size_t i = dim - 1 - ii;
        double sum = yv[i];
        for (size_t k = i + 1; k < dim; ++k) sum -= a[k * dim + i] * x[k];
        x[i] = sum / a[i * dim + i];
    }
}

static double cholesky_logdet(const double *l, size_t dim) {
    double s = 0.0;
    for (size_t i = 0; i < dim; ++i) s += log(l[i * dim + i]);
    return 2.0 * s;
}

double jxo_normal_sf(double z) { return 0.5 * erfc(z / 1.4142135623730951); }

static double clamp_p(double p) {
    /* f64::clamp(MIN_POSITIVE, 1.0); NaN passes through */
    if (p < DBL_MIN) return DBL_MIN;
    if (p > 1.0) return 1.0;
    return p;
}

double jxo_chi2_sf_df1(double stat) {
    if (!isfinite(stat) || stat <= 0.0) return 1.0;
    double p = erfc(sqrt(0.5 * stat));
    if (isfinite(p)) return clamp_p(p);
    return 1.0;
}

/* ------------------------------------------------------------------------------------------
 * A11: shared normal-equation pass of reml_loglike / ml_loglike / final_beta_se.
 * Returns 0 when the reference would bail out before the residual pass.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    size_t dim;
    double a[JXO_MAX_DIM * JXO_MAX_DIM]; /* cholesky factor (lower) */
    double beta[JXO_MAX_DIM];
} normal_eq;

static int build_and_solve(double lbd, const double *s, const double *xcov, const double *y,
                           const double *snp, size_t n, size_t p_cov, double *vinv, normal_eq *ne) {
    size_t dim = p_cov + (snp ? 1 : 0);
    ne->dim = dim;
    for (size_t i = 0; i < n; ++i) {
        double vv = s[i] + lbd;
        if (vv <= 0.0) return 0;
        vinv[i] = 1.0 / vv;
    }
    double *A = ne->a;
    double b[JXO_MAX_DIM];
    for (size_t k = 0; k < dim * dim; ++k) A[k] = 0.0;
    for (size_t k = 0; k < dim; ++k) b[k] = 0.0;
    for (size_t i = 0; i < n; ++i) {
        double vi = vinv[i];
        double yi = y[i];
        for (size_t r = 0; r < dim; ++r) {
            double xir = (r < p_cov) ? xcov[i * p_cov + r] : snp[i];
            b[r] += vi * xir * yi;
            for (size_t c = 0; c <= r; ++c) {
                double xic = (c < p_cov) ? xcov[i * p_cov + c] : snp[i];
                A[r * dim + c] += vi * xir * xic;
            }
        }
    }
    const double ridge = 1e-6;
    for (size_t r = 0; r < dim; ++r) {
        A[r * dim + r] += ridge;
        for (size_t c = 0; c < r; ++c) A[c * dim + r] = A[r * dim + c];
    }
    if (!cholesky_inplace(A, dim)) return 0;
    cholesky_solve(A, dim, b, ne->beta);
    return 1;
}

static double residual_quadratic(const double *vinv, const double *xcov, const double *y,
                                 const double *snp, size_t n, size_t p_cov, const normal_eq *ne) {
    size_t dim = ne->dim;
    double rtv = 0.0;
    for (size_t i = 0; i < n; ++i) {
        double xb = 0.0;
        for (size_t r = 0; r < dim; ++r) {
            double xir = (r < p_cov) ? xcov[i * p_cov + r] : snp[i];
            xb += xir * ne->beta[r];
        }
        double ri = y[i] - xb;
        rtv += vinv[i] * ri * ri;
    }
    return rtv;
}

/* src/stats/reml.rs:255-362 */
double jxo_reml_loglike(double log10_lbd, const double *s, const double *xcov, const double *y,
                        const double *snp, size_t n, size_t p_cov) {
    double lbd = pow(10.0, log10_lbd);
    if (!isfinite(lbd) || lbd <= 0.0) return -1e8;
    size_t p = p_cov + (snp ? 1 : 0);
    if (n <= p || p > JXO_MAX_DIM) return -1e8;
    double *vinv = (double *)malloc(n * sizeof(double));
    normal_eq ne;
    double out = -1e8;
    if (build_and_solve(lbd, s, xcov, y, snp, n, p_cov, vinv, &ne)) {
        double rtv = residual_quadratic(vinv, xcov, y, snp, n, p_cov, &ne);
        double log_det_v = 0.0;
        for (size_t i = 0; i < n; ++i) log_det_v += log(s[i] + lbd);
        double log_det_xtv = cholesky_logdet(ne.a, ne.dim);
        double n_f = (double)n, p_f = (double)p;
        double total_log = (n_f - p_f) * log(rtv) + log_det_v + log_det_xtv;
        double c = (n_f - p_f) * (log(n_f - p_f) - 1.0 - log(2.0 * M_PI)) / 2.0;
        double reml = c - 0.5 * total_log;
        out = isfinite(reml) ? reml : -1e8;
    }
    free(vinv);
    return out;
}

/* src/stats/reml.rs:364-470 */
double jxo_ml_loglike(double log10_lbd, const double *s, const double *xcov, const double *y,
                      const double *snp, size_t n, size_t p_cov) {
    double lbd = pow(10.0, log10_lbd);
    if (!isfinite(lbd) || lbd <= 0.0) return -1e8;
    size_t p = p_cov + (snp ? 1 : 0);
    if (n <= p || p > JXO_MAX_DIM) return -1e8;
    double *vinv = (double *)malloc(n * sizeof(double));
    normal_eq ne;
    double out = -1e8;
    if (build_and_solve(lbd, s, xcov, y, snp, n, p_cov, vinv, &ne)) {
        double rtv = residual_quadratic(vinv, xcov, y, snp, n, p_cov, &ne);
        if (isfinite(rtv) && rtv > 0.0) {
            double log_det_v = 0.0;
            for (size_t i = 0; i < n; ++i) log_det_v += log(s[i] + lbd);
            double n_f = (double)n;
            double total_log = n_f * log(rtv) + log_det_v;
            double c = n_f * (log(n_f) - 1.0 - log(2.0 * M_PI)) / 2.0;
            double ml = c - 0.5 * total_log;
            out = isfinite(ml) ? ml : -1e8;
        }
    }
    free(vinv);
    return out;
}

/* src/stats/reml.rs:472-568 -> (beta_snp, se, lambda) */
void jxo_final_beta_se(double log10_lbd, const double *s, const double *xcov, const double *y,
                       const double *snp, size_t n, size_t p_cov, double out3[3]) {
    double lbd = pow(10.0, log10_lbd);
    out3[0] = NAN; out3[1] = NAN; out3[2] = NAN;
    if (!isfinite(lbd) || lbd <= 0.0) return;
    out3[2] = lbd;
    size_t p = p_cov + 1;
    if (n <= p || p > JXO_MAX_DIM) return;
    double *vinv = (double *)malloc(n * sizeof(double));
    normal_eq ne;
    if (build_and_solve(lbd, s, xcov, y, snp, n, p_cov, vinv, &ne)) {
        double rtv = residual_quadratic(vinv, xcov, y, snp, n, p_cov, &ne);
        double n_f = (double)n, p_f = (double)p;
        double sigma2 = rtv / (n_f - p_f);
        size_t dim = ne.dim, k = dim - 1;
        double e[JXO_MAX_DIM], x[JXO_MAX_DIM];
        for (size_t i = 0; i < dim; ++i) e[i] = 0.0;
        e[k] = 1.0;
        cholesky_solve(ne.a, dim, e, x);
        double var_beta_k = sigma2 * x[k];
        if (!(var_beta_k <= 0.0) && isfinite(var_beta_k)) {
            out3[0] = ne.beta[k];
            out3[1] = sqrt(var_beta_k);
        }
    }
    free(vinv);
}

/* ------------------------------------------------------------------------------------------
 * A12: src/math/brent.rs:16-136 -- note `e` is only refreshed on golden-section steps.
 * ---------------------------------------------------------------------------------------- */
void jxo_brent(jxo_cost_fn f, void *ctx, double low, double high, double tol, size_t max_iter,
               int has_init, double init_x, double *best_x, double *best_f, size_t *n_eval) {
    double a = low, c = high;
    if (!(a < c)) { double t = a; a = c; c = t; }
    const double eps = DBL_EPSILON;
    tol = fabs(tol);
    if (!(tol > 1e-12)) tol = 1e-12; /* f64::max(NaN, 1e-12) = 1e-12 */

    double x;
    if (has_init && isfinite(init_x) && init_x >= a && init_x <= c) x = init_x;
    else x = 0.5 * (a + c);
    double w = x, v = x;
    size_t evals = 0;
    double fx = f(x, ctx); ++evals;
    double fw = fx, fv = fx;
    double d = 0.0, e = 0.0;

    for (size_t it = 0; it < max_iter; ++it) {
        double m = 0.5 * (a + c);
        double tol1 = tol * fabs(x) + eps;
        double tol2 = 2.0 * tol1;
        if (fabs(x - m) <= tol2 - 0.5 * (c - a)) break;

        double u;
        int use_parabolic = 0;
        if (fabs(e) > tol1) {
            double p = (x - v) * ((x - w) * (fx - fv)) - (x - w) * ((x - v) * (fx - fw));
            double q = 2.0 * (((x - v) * (fx - fw)) - ((x - w) * (fx - fv)));
            if (q > 0.0) p = -p; else q = -q;
            int ok = 0;
            if (fabs(q) > eps) {
                double sstep = p / q;
                u = x + sstep;
                if ((u - a) >= tol2 && (c - u) >= tol2 && fabs(sstep) < 0.5 * fabs(e)) ok = 1;
            }
            if (ok) {
                d = p / q;
                u = x + d;
                if ((u - a) < tol2 || (c - u) < tol2) d = (x < m) ? tol1 : -tol1;
                use_parabolic = 1;
            }
        }
        if (!use_parabolic) {
            e = (x < m) ? (c - x) : (a - x);
            d = 0.3819660 * e;
        }
        if (fabs(d) < tol1) d = (d >= 0.0) ? tol1 : -tol1;

        u = x + d;
        double fu = f(u, ctx); ++evals;

        if (fu <= fx) {
            if (u >= x) a = x; else c = x;
            v = w; fv = fw;
            w = x; fw = fx;
            x = u; fx = fu;
        } else {
            if (u >= x) c = u; else a = u;
            if (fu <= fw || w == x) {
                v = w; fv = fw;
                w = u; fw = fu;
            } else if (fu <= fv || v == x || v == w) {
                v = u; fv = fu;
            }
        }
    }
    *best_x = x;
    *best_f = fx;
    if (n_eval) *n_eval = evals;
}

typedef struct {
    const double *s, *xcov, *y, *snp;
    size_t n, p_cov;
} ll_ctx;

static double neg_reml(double x, void *c) {
    ll_ctx *k = (ll_ctx *)c;
    return -jxo_reml_loglike(x, k->s, k->xcov, k->y, k->snp, k->n, k->p_cov);
}
static double neg_ml(double x, void *c) {
    ll_ctx *k = (ll_ctx *)c;
    return -jxo_ml_loglike(x, k->s, k->xcov, k->y, k->snp, k->n, k->p_cov);
}

/* A7: src/stats/reml.rs:594-615 */
void jxo_reml_null(const double *s, const double *xcov, const double *y, size_t n, size_t p_cov,
                   double low, double high, size_t max_iter, double tol, double out3[3]) {
    ll_ctx k = {s, xcov, y, NULL, n, p_cov};
    double bx, bf;
    jxo_brent(neg_reml, &k, low, high, tol, max_iter, 0, 0.0, &bx, &bf, NULL);
    double ml = jxo_ml_loglike(bx, s, xcov, y, NULL, n, p_cov);
    out3[0] = pow(10.0, bx);
    out3[1] = ml;
    out3[2] = -bf;
}

/* src/stats/lmm.rs:2901-2924 */
void jxo_ml_null(const double *s, const double *xcov, const double *y, size_t n, size_t p_cov,
                 double low, double high, size_t max_iter, double tol, int has_init, double init_x,
                 double out2[2]) {
    ll_ctx k = {s, xcov, y, NULL, n, p_cov};
    double bx, bf;
    jxo_brent(neg_ml, &k, low, high, tol, max_iter, has_init, init_x, &bx, &bf, NULL);
    double ml0 = -bf;
    if (!isfinite(ml0)) ml0 = jxo_ml_loglike(bx, s, xcov, y, NULL, n, p_cov);
    out2[0] = bx;
    out2[1] = ml0;
}

/* src/stats/lmm.rs:63-71 */
static double widen_row(const float *src, double *dst, size_t n) {
    double ssq = 0.0;
    for (size_t i = 0; i < n; ++i) {
        double v = (double)src[i];
        dst[i] = v;
        ssq += v * v;
    }
    return ssq;
}

/* A9: src/stats/lmm.rs:94-199 (carry_warm_start = false) */
void jxo_lmm_reml_block(const float *g_rot, size_t rows, size_t n,
                        const double *s, const double *xcov, const double *y, size_t p_cov,
                        double low, double high, double tol, size_t max_iter,
                        int has_init, double init_log10_lbd,
                        int has_nullml, double nullml,
                        double *out, int32_t *n_eval, int threads) {
    int out_cols = has_nullml ? 4 : 3;
#ifdef _OPENMP
    int nt = threads > 0 ? threads : omp_get_max_threads();
#else
    int nt = 1; (void)threads;
#endif
#pragma omp parallel num_threads(nt)
    {
        double *snp = (double *)malloc(n * sizeof(double));
#pragma omp for schedule(dynamic, 1)
        for (long r = 0; r < (long)rows; ++r) {
            double *o = out + (size_t)r * out_cols;
            size_t evals = 0;
            double ssq = widen_row(g_rot + (size_t)r * n, snp, n);
            int valid = 1;
            if (!isfinite(ssq) || ssq <= 1e-12) valid = 0;
            double bx = 0.0, bf = 0.0;
            double fin[3];
            double pwald = 1.0;
            if (valid) {
                ll_ctx k = {s, xcov, y, snp, n, p_cov};
                jxo_brent(neg_reml, &k, low, high, tol, max_iter, has_init, init_log10_lbd, &bx, &bf, &evals);
                jxo_final_beta_se(bx, s, xcov, y, snp, n, p_cov, fin);
                ++evals;
                if (isfinite(fin[0]) && isfinite(fin[1]) && fin[1] > 0.0) {
                    double z = fin[0] / fin[1];
                    pwald = clamp_p(2.0 * jxo_normal_sf(fabs(z)));
                } else {
                    valid = 0;
                }
            }
            if (!valid) {
                o[0] = NAN; o[1] = NAN; o[2] = 1.0;
                if (has_nullml) o[3] = 1.0;
            } else {
                o[0] = fin[0];
                o[1] = fin[1];
                o[2] = isfinite(pwald) ? pwald : 1.0;
                if (has_nullml) {
                    double ml = jxo_ml_loglike(bx, s, xcov, y, snp, n, p_cov);
                    ++evals;
                    if (isfinite(ml)) {
                        double stat = 2.0 * (ml - nullml);
                        if (!isfinite(stat) || stat < 0.0) stat = 0.0;
                        o[3] = jxo_chi2_sf_df1(stat);
                    } else {
                        o[3] = 1.0;
                    }
                }
            }
            if (n_eval) n_eval[r] = (int32_t)evals;
        }
        free(snp);
    }
}

/* A10: src/stats/lmm.rs:202-331 (use_warm_start = false) */
void jxo_lmm2_block(const float *g_rot, size_t rows, size_t n,
                    const double *s, const double *xcov, const double *y, size_t p_cov,
                    double low, double high, double tol, size_t max_iter,
                    int has_init_reml, double init_reml, int has_init_ml, double init_ml,
                    double nullml, double *out, int32_t *n_eval, int threads) {
#ifdef _OPENMP
    int nt = threads > 0 ? threads : omp_get_max_threads();
#else
    int nt = 1; (void)threads;
#endif
    /* lmm.rs:237-244: reml init = init_reml.or(init_ml) */
    int has_r0 = has_init_reml || has_init_ml;
    double r0 = has_init_reml ? init_reml : init_ml;
#pragma omp parallel num_threads(nt)
    {
        double *snp = (double *)malloc(n * sizeof(double));
#pragma omp for schedule(dynamic, 1)
        for (long r = 0; r < (long)rows; ++r) {
            double *o = out + (size_t)r * 6;
            size_t evals = 0, e1 = 0;
            double ssq = widen_row(g_rot + (size_t)r * n, snp, n);
            int valid = 1;
            if (!isfinite(ssq) || ssq <= 1e-12) valid = 0;
            double fin[3] = {NAN, NAN, NAN};
            double pwald = 1.0, bx = 0.0, bf = 0.0;
            ll_ctx k = {s, xcov, y, snp, n, p_cov};
            if (valid) {
                jxo_brent(neg_reml, &k, low, high, tol, max_iter, has_r0, r0, &bx, &bf, &e1);
                evals += e1;
                jxo_final_beta_se(bx, s, xcov, y, snp, n, p_cov, fin);
                ++evals;
                if (isfinite(fin[0]) && isfinite(fin[1]) && fin[1] > 0.0) {
                    double z = fin[0] / fin[1];
                    pwald = clamp_p(2.0 * jxo_normal_sf(fabs(z)));
                } else {
                    valid = 0;
                }
            }
            if (!valid) {
                o[0] = NAN; o[1] = NAN; o[2] = 1.0; o[3] = NAN; o[4] = NAN; o[5] = 1.0;
            } else {
                /* lmm.rs:285-296: ml init = Some(best_reml) */
                double mx, mf;
                jxo_brent(neg_ml, &k, low, high, tol, max_iter, 1, bx, &mx, &mf, &e1);
                evals += e1;
                double ml_alt = -mf;
                if (!isfinite(ml_alt)) {
                    ml_alt = jxo_ml_loglike(mx, s, xcov, y, snp, n, p_cov);
                    ++evals;
                }
                double stat = isfinite(ml_alt) ? 2.0 * (ml_alt - nullml) : 0.0;
                if (!isfinite(stat) || stat < 0.0) stat = 0.0;
                double plrt = jxo_chi2_sf_df1(stat);
                o[0] = fin[0];
                o[1] = fin[1];
                o[2] = isfinite(pwald) ? pwald : 1.0;
                o[3] = fin[2];
                o[4] = ml_alt;
                o[5] = isfinite(plrt) ? plrt : 1.0;
            }
            if (n_eval) n_eval[r] = (int32_t)evals;
        }
        free(snp);
    }
}

/* ------------------------------------------------------------------------------------------
 * A14: fixed-lambda (-fvlmm).  src/stats/fvlmm.rs:1484-1563, 1691-1805
 * ---------------------------------------------------------------------------------------- */
int jxo_fixed_cache_prepare(const double *s, const double *xcov, const double *y, size_t n, size_t p,
                            double lbd, jxo_fixed_cache *c) {
    memset(c, 0, sizeof(*c));
    if (p > JXO_MAX_DIM) return -4;
    c->n = n; c->p = p;
    c->w = (float *)malloc(n * sizeof(float));
    c->py_tilde = (float *)malloc(n * sizeof(float));
    c->wx_tilde = (float *)malloc(n * (p ? p : 1) * sizeof(float));
    c->a_chol = (double *)calloc((p ? p * p : 1), sizeof(double));
    double log_det_v = 0.0;
    for (size_t i = 0; i < n; ++i) {
        double vv = s[i] + lbd;
        if (!(isfinite(vv) && vv > 0.0)) return -1;
        c->w[i] = (float)(1.0 / vv);
        log_det_v += log(vv);
    }
    double *a = c->a_chol;
    double b[JXO_MAX_DIM];
    for (size_t r = 0; r < p; ++r) b[r] = 0.0;
    double ywy = 0.0;
    for (size_t i = 0; i < n; ++i) {
        double wi = (double)c->w[i];
        double yi = y[i];
        ywy += wi * yi * yi;
        size_t base = i * p;
        for (size_t r = 0; r < p; ++r) {
            double xir = xcov[base + r];
            b[r] += wi * xir * yi;
            for (size_t cc = 0; cc <= r; ++cc) a[r * p + cc] += wi * xir * xcov[base + cc];
        }
    }
    for (size_t r = 0; r < p; ++r) {
        a[r * p + r] += 1e-6;
        for (size_t cc = 0; cc < r; ++cc) a[cc * p + r] = a[r * p + cc];
    }
    if (!cholesky_inplace(a, p)) return -2;
    double a_inv_b[JXO_MAX_DIM];
    cholesky_solve(a, p, b, a_inv_b);
    double bd = 0.0;
    for (size_t r = 0; r < p; ++r) bd += b[r] * a_inv_b[r];
    double ypy = ywy - bd;
    if (!(ypy > 0.0)) ypy = 0.0;
    for (size_t i = 0; i < n; ++i) {
        double wi = (double)c->w[i];
        size_t base = i * p;
        double x_aib = 0.0;
        for (size_t r = 0; r < p; ++r) {
            double xir = xcov[base + r];
            c->wx_tilde[base + r] = (float)(wi * xir);
            x_aib += xir * a_inv_b[r];
        }
        c->py_tilde[i] = (float)(wi * (y[i] - x_aib));
    }
    c->ypy = ypy;
    c->log_det_v = log_det_v;
    c->df = (int)n - (int)p - 1;
    if (c->df <= 0) return -3;
    return 0;
}

void jxo_fixed_cache_free(jxo_fixed_cache *c) {
    free(c->w); free(c->py_tilde); free(c->wx_tilde); free(c->a_chol);
    memset(c, 0, sizeof(*c));
}

void jxo_fixed_lambda_block(const float *g_rot, size_t rows, const jxo_fixed_cache *c,
                            int has_nullml, double nullml, double *out, int threads) {
    size_t n = c->n, p = c->p;
    int out_cols = has_nullml ? 4 : 3;
    double n_f = (double)n;
    double c_ml = n_f * (log(n_f) - 1.0 - log(2.0 * M_PI)) / 2.0;
#ifdef _OPENMP
    int nt = threads > 0 ? threads : omp_get_max_threads();
#else
    int nt = 1; (void)threads;
#endif
#pragma omp parallel for schedule(static) num_threads(nt)
    for (long r = 0; r < (long)rows; ++r) {
        const float *row = g_rot + (size_t)r * n;
        double *o = out + (size_t)r * out_cols;
        /* the two SGEMMs (fvlmm.rs:1711-1730), pinned: f64 accumulate, f32 store */
        double acc = 0.0;
        for (size_t i = 0; i < n; ++i) acc += (double)row[i] * (double)c->py_tilde[i];
        float num_f = (float)acc;
        double cv[JXO_MAX_DIM], a_inv_c[JXO_MAX_DIM];
        for (size_t k = 0; k < p; ++k) {
            double a2 = 0.0;
            for (size_t i = 0; i < n; ++i) a2 += (double)row[i] * (double)c->wx_tilde[i * p + k];
            cv[k] = (double)(float)a2;
        }
        double d = 0.0;
        for (size_t i = 0; i < n; ++i) {
            double gi = (double)row[i];
            d += ((double)c->w[i]) * gi * gi;
        }
        cholesky_solve(c->a_chol, p, cv, a_inv_c);
        double ct = 0.0;
        for (size_t k = 0; k < p; ++k) ct += cv[k] * a_inv_c[k];
        double schur = d - ct;
        if (schur <= 1e-12 || !isfinite(schur)) {
            o[0] = NAN; o[1] = NAN; o[2] = NAN;
            if (has_nullml) o[3] = 1.0;
            continue;
        }
        double num = (double)num_f;
        double beta_g = num / schur;
        double rwr = c->ypy - (num * num) / schur;
        if (!(rwr > 0.0)) rwr = 0.0;
        double sigma2 = rwr / (double)c->df;
        double se_g = sqrt(sigma2 / schur);
        double pval = 1.0;
        if (isfinite(se_g) && se_g > 0.0 && isfinite(beta_g)) {
            double z = fabs(beta_g / se_g);
            pval = clamp_p(2.0 * jxo_normal_sf(z));
        }
        o[0] = beta_g; o[1] = se_g; o[2] = pval;
        if (has_nullml) {
            double ml = NAN;
            if (rwr > 0.0 && isfinite(rwr)) {
                double total_log = n_f * log(rwr) + c->log_det_v;
                ml = c_ml - 0.5 * total_log;
            }
            double stat = isfinite(ml) ? 2.0 * (ml - nullml) : 0.0;
            if (!isfinite(stat) || stat < 0.0) stat = 0.0;
            o[3] = jxo_chi2_sf_df1(stat);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * A16: TSV row formatting.  Rust `{:.4}` == C "%.4f" except NaN/inf spelling; Rust `{:.4e}`
 * prints the exponent without sign padding ("1.2345e-5", "0.0000e0").
 * ---------------------------------------------------------------------------------------- */
size_t jxo_fmt_fixed(char *buf, size_t cap, double v, int prec) {
    if (isnan(v)) return (size_t)snprintf(buf, cap, "NaN");
    if (isinf(v)) return (size_t)snprintf(buf, cap, v > 0 ? "inf" : "-inf");
    return (size_t)snprintf(buf, cap, "%.*f", prec, v);
}

size_t jxo_fmt_exp(char *buf, size_t cap, double v, int prec) {
    if (isnan(v)) return (size_t)snprintf(buf, cap, "NaN");
    if (isinf(v)) return (size_t)snprintf(buf, cap, v > 0 ? "inf" : "-inf");
    char tmp[64];
    snprintf(tmp, sizeof tmp, "%.*e", prec, v);
    char *e = strchr(tmp, 'e');
    int ex = atoi(e + 1);
    *e = '\0';
    return (size_t)snprintf(buf, cap, "%se%d", tmp, ex);
}

/* src/math/linalg.rs:99-108 */
static double sanitize_p(double beta, double se, double p) {
    if (!(isfinite(beta) && isfinite(se) && se > 0.0)) return 1.0;
    if (isfinite(p)) return clamp_p(p);
    return 1.0;
}

size_t jxo_format_row(char *buf, size_t cap, const char *chrom, int64_t pos, const char *snp,
                      const char *a0, const char *a1, float af, float miss_rate,
                      const double *row, int out_cols) {
    char f_af[48], f_ms[48], f_b[48], f_se[48], f_chi[48], f_p[48];
    double beta = row[0], se = row[1];
    double pw = sanitize_p(beta, se, row[2]);
    double chisq = NAN;
    if (isfinite(beta) && isfinite(se) && se > 0.0) { double z = beta / se; chisq = z * z; }
    jxo_fmt_fixed(f_af, sizeof f_af, (double)af, 4);
    jxo_fmt_fixed(f_ms, sizeof f_ms, (double)miss_rate, 4);
    jxo_fmt_fixed(f_b, sizeof f_b, beta, 4);
    jxo_fmt_fixed(f_se, sizeof f_se, se, 4);
    jxo_fmt_exp(f_chi, sizeof f_chi, chisq, 4);
    jxo_fmt_exp(f_p, sizeof f_p, pw, 4);
    /* snp name: src/stats/lmm.rs:1952-1958 */
    char namebuf[512];
    const char *name = snp;
    if (snp[0] == '\0' || strcmp(snp, ".") == 0) {
        snprintf(namebuf, sizeof namebuf, "%s_%lld", chrom, (long long)pos);
        name = namebuf;
    }
    int w = snprintf(buf, cap, "%s\t%lld\t%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s", chrom, (long long)pos, name,
                     a0, a1, f_af, f_ms, f_b, f_se, f_chi, f_p);
    size_t off = (size_t)w;
    if (out_cols == 4) {
        char f3[48];
        jxo_fmt_exp(f3, sizeof f3, row[3], 4);
        off += (size_t)snprintf(buf + off, cap - off, "\t%s", f3);
    } else if (out_cols == 6) {
        char f3[48], f4[48], f5[48];
        jxo_fmt_exp(f3, sizeof f3, row[3], 6);
        jxo_fmt_exp(f4, sizeof f4, row[4], 6);
        jxo_fmt_exp(f5, sizeof f5, row[5], 4);
        off += (size_t)snprintf(buf + off, cap - off, "\t%s\t%s\t%s", f3, f4, f5);
    }
    off += (size_t)snprintf(buf + off, cap - off, "\n");
    return off;
}
