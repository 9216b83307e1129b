// cabi.cu -- C ABI of libjxb200.so (see include/jxb200.h): model residency, staging, batch pipeline.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/jxb200.h"
#include "jxb_common.cuh"

struct jxb_model {
    jxb::Model m;
    bool timing_ready = false;
    cudaEvent_t ev[10];
    bool ev_rec[10] = {false, false, false, false, false, false, false, false, false, false};
    float stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float* missr = nullptr;       // [cap_rows] device
    uint8_t* mask = nullptr;      // [cap_rows] device (pre-keep mask)
    float* prep_af = nullptr;     // [cap_rows] device: caller-supplied allele frequency per source row (prepared metadata)
    uint8_t* prep_flip = nullptr; // [cap_rows] device: caller-supplied row_flip
    // streamed scan: the solve runs on stream2 while m.stream still rotates later row slabs
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_slab0 = nullptr, ev_solve = nullptr;
    // double-buffered input staging (jxb_stage_packed / jxb_scan_staged): batch i+1 goes up on copy_stream while batch i computes
    uint8_t* packed2 = nullptr;
    size_t packed2_bytes = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy = nullptr;
    size_t staged_rows = 0, staged_bps = 0, cur_rows = 0, cur_bps = 0;
    bool staged = false, cur_valid = false;
    bool last_streamed = false;
    size_t last_nk = 0;           // kept rows of the last scan (known on the host after the decode stage)
    double* scal = nullptr;       // small device scratch (null fit outputs)
    size_t last_rows = 0;
    bool rot_dirty = false;       // rot holds a transposed block: re-zero before the row-major kernels read padding
    int last_out_cols = 0;
};

namespace jxb {

static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};
static int g_timing = 0;
// rotation for packed additive scans: 3 = hand-written tcgen05 int8-sliced exact rotation (default),
// 2 = same arithmetic with cuBLASLt slice GEMMs, 0 = FP64 DMMA GEMM, 1 = CUDA-core cross-check.
// Chunk entry points that take arbitrary f32 genotypes always use the FP64 DMMA GEMM (or 1).
static int g_rotate_variant = 3;
static size_t g_thread_solve_min_rows = 32768;   // batches at least this large use the large-batch solve kernels
static int g_big_solve_kernel = 0;               // 0 = lane-per-SNP with refill (row-major block), 1 = thread-per-SNP (SNP-minor)
static int g_stream_overlap = 0;                 // 1 = large additive LMM/LMM2 batches rotate and solve concurrently (scan_streamed);
                                                 // off by default: both kernels are bound by the shared-memory pipe (profiles/README.md)
static size_t g_stream_slab_rows = 8192;         // rows per rotation slab of the streamed scan (multiple of 256)

void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
void note_launch(int k) { g_launches.fetch_add((uint64_t)k, std::memory_order_relaxed); }

namespace {

__global__ void widen_ut_kernel(const float* __restrict__ src, size_t n, size_t row0, size_t rows,
                                double* __restrict__ dst, size_t ldk) {
    for (size_t r = blockIdx.y; r < rows; r += gridDim.y) {
        const float* s = src + r * n;
        double* d = dst + (row0 + r) * ldk;
        for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x)
            d[j] = (double)s[j];
    }
}

__global__ void apply_mask_kernel(int32_t* counts, const uint8_t* mask, int rows) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows && mask[r] == 0) counts[4 * r + 3] = 0;
}

// Prepared row metadata (src/stats/lmm.rs:1237-1262; decode.rs:213-219): the caller's allele frequency replaces the
// one counted on the device (it is the imputation mean the reference uses, whatever samples it was computed over) and
// row_flip travels in bit 1 of the keep word to the decode kernels.
__global__ void apply_prepared_kernel(int32_t* counts, float* af, const float* row_af, const uint8_t* row_flip, int rows) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    if (row_af) af[r] = row_af[r];
    if (row_flip && row_flip[r] && counts[4 * r + 3] != 0) counts[4 * r + 3] |= 2;
}

template <class T>
int dev_alloc(T** p, size_t count) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (count == 0) return 0;
    JXB_CUDA_OK(cudaMalloc((void**)p, count * sizeof(T)));
    return 0;
}

int upload_small(Model& m, const double* s, const double* xcov, const double* y) {
    const size_t n = m.n, p = m.p;
    std::vector<double> xt(p * m.ldn, 0.0);
    for (size_t i = 0; i < n; ++i)
        for (size_t r = 0; r < p; ++r) xt[r * m.ldn + i] = xcov[i * p + r];
    if (s) {
        JXB_CUDA_OK(cudaMemcpyAsync(m.s, s, n * sizeof(double), cudaMemcpyHostToDevice, m.stream));
        m.s_host.assign(s, s + n);
    }
    // interleaved K3 records; padding samples: s = 1 (so v > 0), everything else 0
    std::vector<double> rec(m.ldn * m.rs, 0.0);
    for (size_t i = 0; i < m.ldn; ++i) {
        double* r = rec.data() + i * m.rs;
        if (i < n) {
            r[0] = m.s_host[i];
            r[1] = y[i];
            for (size_t c = 0; c < p; ++c) r[2 + c] = xcov[i * p + c];
        } else {
            r[0] = 1.0;
        }
    }
    JXB_CUDA_OK(cudaMemcpyAsync(m.rec, rec.data(), rec.size() * sizeof(double), cudaMemcpyHostToDevice, m.stream));
    JXB_CUDA_OK(cudaMemcpyAsync(m.y, y, n * sizeof(double), cudaMemcpyHostToDevice, m.stream));
    JXB_CUDA_OK(cudaMemcpyAsync(m.xt, xt.data(), xt.size() * sizeof(double), cudaMemcpyHostToDevice, m.stream));
    JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
    m.fx_valid = false;
    m.prefix_valid = false;
    return 0;
}

int widen_ut(Model& m, const float* u_t, bool on_device) {
    const size_t n = m.n;
    JXB_CUDA_OK(cudaMemsetAsync(m.ut, 0, m.n_pad * m.ldk * sizeof(double), m.stream));
    // stage in row slabs of <= 256 MiB so the temporary never doubles the footprint
    const size_t slab_rows = std::max<size_t>(1, std::min<size_t>(n, (256u << 20) / (n * sizeof(float) + 1)));
    float* tmp = nullptr;
    if (!on_device) JXB_CUDA_OK(cudaMalloc((void**)&tmp, slab_rows * n * sizeof(float)));
    for (size_t r0 = 0; r0 < n; r0 += slab_rows) {
        const size_t rows = std::min(slab_rows, n - r0);
        const float* src = u_t + r0 * n;
        if (!on_device) {
            JXB_CUDA_OK(cudaMemcpyAsync(tmp, src, rows * n * sizeof(float), cudaMemcpyHostToDevice, m.stream));
            src = tmp;
        }
        dim3 grid((unsigned)std::min<size_t>((n + 255) / 256, 64), (unsigned)std::min<size_t>(rows, 2048));
        widen_ut_kernel<<<grid, 256, 0, m.stream>>>(src, n, r0, rows, m.ut, m.ldk);
        note_launch(1);
        JXB_CUDA_OK(cudaGetLastError());
        if (!on_device) JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
    }
    JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
    if (tmp) cudaFree(tmp);
    return 0;
}

int model_alloc(int device, size_t n, size_t p, bool with_ut, jxb_model** out) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(-1, "no CUDA device is visible: libjxb200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(-2, "device index out of range");
    if (n == 0) return fail(-2, "y must not be empty");
    if (p < 1 || p > 32) return fail(-2, "Xcov must have between 1 and 32 columns");
    if (n > (size_t)0x7fffffff) return fail(-2, "n too large");
    JXB_CUDA_OK(cudaSetDevice(device));
    jxb_model* h = new jxb_model();
    Model& m = h->m;
    m.device = device;
    m.n = n;
    m.p = p;
    m.ldn = round_up(n, 32);
    m.ldk = round_up(n, kRotBK);
    m.n_pad = round_up(n, kRotBN);
    m.ldc = round_up(n, 32);
    JXB_CUDA_OK(cudaStreamCreateWithFlags(&m.stream, cudaStreamNonBlocking));
    int rc = 0;
    rc |= dev_alloc(&m.s, n);
    rc |= dev_alloc(&m.y, n);
    rc |= dev_alloc(&m.xt, p * m.ldn);
    m.rs = (p + 2 + 1) / 2 * 2;
    rc |= dev_alloc(&m.rec, m.ldn * m.rs);
    rc |= dev_alloc(&m.n_kept, 8);
    rc |= dev_alloc(&h->scal, 16);
    rc |= dev_alloc(&m.fx_w, m.ldn);
    rc |= dev_alloc(&m.fx_py, m.ldn);
    rc |= dev_alloc(&m.fx_wx, p * m.ldn);
    rc |= dev_alloc(&m.fx_scal, 8 + 32 * 32);
    if (p <= 8) rc |= dev_alloc(&m.fx_rec, m.ldn * ((p + 2 + 1) / 2 * 2));
    if (with_ut) rc |= dev_alloc(&m.ut, m.n_pad * m.ldk);
    if (rc) { jxb_model_destroy(h); return rc; }
    *out = h;
    return 0;
}

int ensure_capacity(jxb_model* h, size_t rows, size_t bps, bool need_g) {
    Model& m = h->m;
    JXB_CUDA_OK(cudaSetDevice(m.device));
    if (rows > m.cap_rows) {
        JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
        const size_t cap = round_up(std::max<size_t>(rows, 256), kRotBM);
        int rc = 0;
        rc |= dev_alloc(&m.rot, cap * m.ldc);
        if (!rc) cudaMemset(m.rot, 0, cap * m.ldc * sizeof(float));   // K3 reads the zero padding past n
        rc |= dev_alloc(&m.out, cap * 8);
        rc |= dev_alloc(&m.evals, cap);
        rc |= dev_alloc(&m.counts, cap * 4);
        rc |= dev_alloc(&m.af, cap);
        rc |= dev_alloc(&h->missr, cap);
        rc |= dev_alloc(&h->mask, cap);
        rc |= dev_alloc(&h->prep_af, cap);
        rc |= dev_alloc(&h->prep_flip, cap);
        rc |= dev_alloc(&m.src_row, cap);
        if (m.g64) { cudaFree(m.g64); m.g64 = nullptr; }
        if (m.packed) { cudaFree(m.packed); m.packed = nullptr; m.bps_cap = 0; }
        if (rc) return rc;
        m.cap_rows = cap;
    }
    if (need_g && !m.g64) {
        JXB_CUDA_OK(cudaMalloc((void**)&m.g64, m.cap_rows * m.ldk * sizeof(double)));
        JXB_CUDA_OK(cudaMemsetAsync(m.g64, 0, m.cap_rows * m.ldk * sizeof(double), m.stream));
        if (m.ut) {
            int rc = make_tensor_maps(m);
            if (rc) return rc;
        }
    }
    if (bps > 0 && (bps > m.bps_cap || !m.packed)) {
        JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
        if (m.packed) cudaFree(m.packed);
        m.packed = nullptr;
        JXB_CUDA_OK(cudaMalloc((void**)&m.packed, m.cap_rows * bps));
        m.bps_cap = bps;
    }
    return 0;
}

// rot may hold a transposed (SNP-minor) block from the large-batch path: restore the zero padding the
// row-major kernels rely on before anything row-major is written or read
int clean_rot(jxb_model* h) {
    if (h->rot_dirty && h->m.rot) {
        JXB_CUDA_OK(cudaMemsetAsync(h->m.rot, 0, h->m.cap_rows * h->m.ldc * sizeof(float), h->m.stream));
        h->rot_dirty = false;
    }
    return 0;
}

int ensure_stage_f32(Model& m, size_t count) {
    if (count > m.stage_f32_cap) {
        JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
        if (m.stage_f32) cudaFree(m.stage_f32);
        m.stage_f32 = nullptr;
        JXB_CUDA_OK(cudaMalloc((void**)&m.stage_f32, count * sizeof(float)));
        m.stage_f32_cap = count;
    }
    return 0;
}

int ensure_sample_idx(Model& m, const int64_t* idx, size_t n_sel, bool on_device, size_t n_full) {
    if (!on_device) {
        // the kernels index packed rows with these values unchecked: one host pass over n entries
        for (size_t k = 0; k < n_sel; ++k)
            if (idx[k] < 0 || (size_t)idx[k] >= n_full)
                return fail(-2, "sample_idx[" + std::to_string(k) + "]=" + std::to_string((long long)idx[k]) +
                                    " is outside [0, n_full=" + std::to_string(n_full) + ")");
    }
    if (n_sel > m.n_sel_cap) {
        JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
        if (m.sample_idx) cudaFree(m.sample_idx);
        m.sample_idx = nullptr;
        JXB_CUDA_OK(cudaMalloc((void**)&m.sample_idx, n_sel * sizeof(int64_t)));
        m.n_sel_cap = n_sel;
    }
    JXB_CUDA_OK(cudaMemcpyAsync(m.sample_idx, idx, n_sel * sizeof(int64_t),
                                on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, m.stream));
    return 0;
}

void tick(jxb_model* h, int i, cudaStream_t st = nullptr) {
    if (!g_timing) return;
    if (!h->timing_ready) {
        for (auto& e : h->ev) cudaEventCreate(&e);
        h->timing_ready = true;
    }
    cudaEventRecord(h->ev[i], st ? st : h->m.stream);
    h->ev_rec[i] = true;
}

SolveParams to_params(const jxb_solve_cfg* c, int mode);
int out_cols_of(const jxb_solve_cfg* c, int mode);

// ---- streamed scan -------------------------------------------------------------------------------------------------
// Rotation (tensor pipe) and per-SNP solve (FP64 pipe) of ONE batch at the same time on the same SMs: the decoded
// batch is rotated in slabs of g_stream_slab_rows rows on m.stream; after the first slab a single persistent
// solve_lane_stream_kernel starts on stream2 and consumes rows as the rotation publishes them (sync[1]).  Three solve
// CTAs (128 threads x 136 registers, 25 KB) and one SLIM i8_rotate_kernel CTA (192 x 64 registers, 121 KB) fit one SM
// together; streamed_feasible() checks that from the compiled kernels' own attributes before the mode is used, and the
// solve kernel carries a watchdog, so a rotation that cannot become resident is an error, never a hang.
bool streamed_feasible(Model& m) {
    static int cache[5] = {0, 0, 0, 0, 0};   // per covariate count: 0 unknown, 1 yes, -1 no
    if (m.p < 1 || m.p > 4) return false;
    if (cache[m.p] == 0) {
        int r_regs = 0, r_smem = 0, s_regs = 0, s_smem = 0;
        bool ok = rotate_slim_resources(&r_regs, &r_smem) == 0 && solve_lane_stream_resources(m.p, &s_regs, &s_smem) == 0;
        if (ok) {
            cudaDeviceProp prop;
            ok = cudaGetDeviceProperties(&prop, m.device) == cudaSuccess;
            if (ok) {
                // Registers live in four per-scheduler files of regsPerMultiprocessor / 4 each; a CTA's warps are dealt
                // round-robin to the schedulers and a warp is allocated in units of 256 registers.  Worst scheduler:
                // 3 solve CTAs x 1 warp + 2 of the rotation CTA's 6 warps.  1 KB of shared memory is reserved per CTA.
                auto warp_regs = [](int r) { return (r * 32 + 255) / 256 * 256; };
                const int per_sched = prop.regsPerMultiprocessor / 4;
                const int regs = 3 * warp_regs(s_regs) + 2 * warp_regs(r_regs);
                const size_t smem = 3 * (size_t)(s_smem + 1024) + (size_t)(r_smem + 1024);
                // ... and a 4th solve CTA must not be able to take the rotation's slot
                ok = regs <= per_sched && smem <= prop.sharedMemPerMultiprocessor && 4 * warp_regs(s_regs) > per_sched;
            }
        }
        cache[m.p] = ok ? 1 : -1;
    }
    return cache[m.p] == 1;
}

int scan_streamed(jxb_model* h, size_t nk, bool has_missing, const jxb_solve_cfg* cfg, int mode) {
    Model& m = h->m;
    if (!h->stream2) {
        JXB_CUDA_OK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
        JXB_CUDA_OK(cudaEventCreateWithFlags(&h->ev_slab0, cudaEventDisableTiming));
        JXB_CUDA_OK(cudaEventCreateWithFlags(&h->ev_solve, cudaEventDisableTiming));
    }
    const size_t slab = std::max<size_t>(256, g_stream_slab_rows / 256 * 256);
    int rc = prepare_rotate_tc(m, std::min(round_up(slab, 128), m.a8_rows));
    if (rc) return rc;
    rc = ensure_solve_lane_buffers(m, nk, m.stream);
    if (rc) return rc;
    const int oc = out_cols_of(cfg, mode);
    h->last_out_cols = oc;
    int32_t* sync = m.n_kept + 4;   // {queue, ready, abort}
    JXB_CUDA_OK(cudaMemsetAsync(sync, 0, 3 * sizeof(int32_t), m.stream));
    for (size_t r0 = 0; r0 < nk; r0 += slab) {
        const size_t r1 = std::min(nk, r0 + slab);
        rc = launch_rotate_int8_tc_slab(m, r0, r1, has_missing, true, m.stream);
        if (rc) return rc;
        rc = launch_row_ssq_publish(m, m.rot, m.ldc, r0, r1, sync, m.stream);
        note_launch(2);
        if (rc) return rc;
        // g_stream_overlap == 2 (diagnostic): the same slim kernels, but the solve starts after the LAST slab
        if (g_stream_overlap == 2 ? r1 == nk : r0 == 0) {
            JXB_CUDA_OK(cudaEventRecord(h->ev_slab0, m.stream));
            JXB_CUDA_OK(cudaStreamWaitEvent(h->stream2, h->ev_slab0, 0));
            tick(h, 8, h->stream2);
            rc = launch_solve_lane_stream(m, m.rot, m.ldc, nk, to_params(cfg, mode), m.out, oc, m.evals, sync, h->stream2);
            note_launch(1);
            if (rc) return rc;
            tick(h, 9, h->stream2);
            JXB_CUDA_OK(cudaEventRecord(h->ev_solve, h->stream2));
        }
    }
    tick(h, 4);                                                 // rotation done (solve still running)
    JXB_CUDA_OK(cudaStreamWaitEvent(m.stream, h->ev_solve, 0)); // join: later work on m.stream sees the results
    h->last_streamed = true;
    return 0;
}

SolveParams to_params(const jxb_solve_cfg* c, int mode) {
    SolveParams sp;
    sp.low = c->low; sp.high = c->high; sp.tol = c->tol; sp.max_iter = c->max_iter;
    sp.has_init = c->has_init; sp.init = c->init_log10_lbd;
    sp.has_nullml = c->has_nullml; sp.nullml = c->nullml;
    sp.mode = mode;
    return sp;
}

int check_cfg(const jxb_solve_cfg* c, int mode) {
    if (!c) return fail(-2, "solve cfg is null");
    if (mode != 2) {
        if (!(c->low < c->high)) return fail(-2, "low must be < high");
        if (!(std::isfinite(c->tol) && c->tol > 0.0)) return fail(-2, "tol must be positive and finite");
        if (c->max_iter < 0) return fail(-2, "max_iter must be >= 0");
    }
    if (mode == 1 && !c->has_nullml) return fail(-2, "LMM2 needs nullml (fit it with jxb_ml_null)");
    if (mode == 1 && !std::isfinite(c->nullml)) return fail(-2, "nullml must be finite when provided");
    return 0;
}

int out_cols_of(const jxb_solve_cfg* c, int mode) { return mode == 1 ? 6 : (c->has_nullml ? 4 : 3); }

// rotated f32 block already in m.rot -> out
int run_solve(jxb_model* h, size_t rows, const int32_t* n_rows_dev, const jxb_solve_cfg* cfg, int mode) {
    Model& m = h->m;
    const int oc = out_cols_of(cfg, mode);
    h->last_out_cols = oc;
    if (mode == 2) {
        if (!m.fx_valid || m.fx_log10_lbd != cfg->init_log10_lbd) {
            int rc = launch_fixed_prepare(m, cfg->init_log10_lbd, m.stream);
            note_launch(1);
            if (rc) return rc;
            double st[4];
            JXB_CUDA_OK(cudaMemcpyAsync(st, m.fx_scal, sizeof st, cudaMemcpyDeviceToHost, m.stream));
            JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
            if (st[3] == -1.0) return fail(-10, "non-positive s[i]+lbd");
            if (st[3] == -2.0) return fail(-11, "X'WX not SPD");
            if (st[3] == -3.0) return fail(-12, "df <= 0");
            m.fx_valid = true;
            m.fx_log10_lbd = cfg->init_log10_lbd;
        }
        note_launch(1);
        return launch_fixed_solve(m, m.rot, m.ldc, rows, n_rows_dev, cfg->has_nullml, cfg->nullml, m.out, oc, m.stream);
    }
    note_launch(1);
    return launch_solve(m, m.rot, m.ldc, rows, n_rows_dev, to_params(cfg, mode), m.out, oc, m.evals, m.n_kept + 1,
                        m.stream);
}

// f32 host block -> (optionally rotate) -> solve -> host
int chunk_common(jxb_model* h, const float* host, size_t rows, bool rotated, const jxb_solve_cfg* cfg, int mode,
                 double* out_host, int32_t* evals_host) {
    if (!h) return fail(-2, "model is null");
    int rc = check_cfg(cfg, mode);
    if (rc) return rc;
    Model& m = h->m;
    if (rows == 0) return 0;
    if (!rotated && !m.ut) return fail(-3, "u_t must be (n, n) and row-major U^T");
    rc = ensure_capacity(h, rows, 0, !rotated);
    if (rc) return rc;
    rc = clean_rot(h);
    if (rc) return rc;
    const size_t n = m.n;
    if (rotated) {
        JXB_CUDA_OK(cudaMemcpy2DAsync(m.rot, m.ldc * sizeof(float), host, n * sizeof(float), n * sizeof(float), rows,
                                      cudaMemcpyHostToDevice, m.stream));
    } else {
        rc = ensure_stage_f32(m, rows * n);
        if (rc) return rc;
        JXB_CUDA_OK(cudaMemcpyAsync(m.stage_f32, host, rows * n * sizeof(float), cudaMemcpyHostToDevice, m.stream));
        rc = launch_widen_f32(m.stage_f32, n, rows, n, m.g64, m.ldk, m.stream);
        note_launch(1);
        if (rc) return rc;
        rc = launch_rotate(m, rows, nullptr, m.rot, m.ldc, 0, m.stream, g_rotate_variant);
        note_launch(1);
        if (rc) return rc;
    }
    rc = run_solve(h, rows, nullptr, cfg, mode);
    if (rc) return rc;
    const int oc = h->last_out_cols;
    JXB_CUDA_OK(cudaMemcpyAsync(out_host, m.out, rows * oc * sizeof(double), cudaMemcpyDeviceToHost, m.stream));
    if (evals_host && mode != 2)
        JXB_CUDA_OK(cudaMemcpyAsync(evals_host, m.evals, rows * sizeof(int32_t), cudaMemcpyDeviceToHost, m.stream));
    JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
    return 0;
}

int scan_device_stages(jxb_model* h, const uint8_t* packed_dev, size_t bps, size_t rows, size_t n_full,
                       const int64_t* sidx_dev, size_t n_sel, bool have_mask, const jxb_qc_cfg* qc,
                       const jxb_solve_cfg* cfg, int mode, const float* prep_af_dev = nullptr,
                       const uint8_t* prep_flip_dev = nullptr) {
    Model& m = h->m;
    h->last_nk = (size_t)-1;    // unknown on the host until the decode stage of the int8 path reads it back
    h->last_streamed = false;
    int rc = clean_rot(h);
    if (rc) return rc;
    tick(h, 1);
    rc = launch_count_qc(m, packed_dev, bps, rows, n_full, sidx_dev, n_sel, qc->maf_thr, qc->miss_thr, qc->het_thr,
                         m.counts, m.af, h->missr, m.stream);
    note_launch(1);
    if (rc) return rc;
    if (have_mask) {
        apply_mask_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, m.stream>>>(m.counts, h->mask, (int)rows);
        note_launch(1);
    }
    if (prep_af_dev || prep_flip_dev) {
        apply_prepared_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, m.stream>>>(m.counts, m.af, prep_af_dev, prep_flip_dev,
                                                                                  (int)rows);
        note_launch(1);
    }
    rc = launch_compact(m.counts, rows, m.src_row, m.n_kept, m.stream);
    note_launch(1);
    if (rc) return rc;
    tick(h, 2);
    if ((g_rotate_variant == 2 || g_rotate_variant == 3) && qc->genetic_model == JXB_MODEL_ADD) {
        // int8-sliced exact rotation (k2_int8.cu): int8 operands instead of the f64 block (additive coding
        // only: for dom/rec/het the T_2 coefficient is O(1) and the DMMA path is used)
        rc = prepare_int8_slices(m, m.stream);
        if (rc) return rc;
        rc = ensure_int8_workspace(m, m.cap_rows);
        if (rc) return rc;
        rc = launch_decode_int8(m, packed_dev, bps, m.src_row, m.n_kept, rows, n_full, sidx_dev, m.af, m.counts,
                                qc->genetic_model, m.stream);
        if (rc) return rc;
        tick(h, 3);
        int32_t nk = 0, anym = 0;
        JXB_CUDA_OK(cudaMemcpyAsync(&nk, m.n_kept, sizeof nk, cudaMemcpyDeviceToHost, m.stream));
        JXB_CUDA_OK(cudaMemcpyAsync(&anym, m.flags8, sizeof anym, cudaMemcpyDeviceToHost, m.stream));
        JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
        // large batches: SNP-minor output + one-thread-per-SNP solve (no shared-memory transposition)
        const bool big = mode != 2 && m.p <= 8 && (size_t)nk >= g_thread_solve_min_rows;
        h->last_streamed = false;
        h->last_nk = (size_t)nk;
        if (big && g_big_solve_kernel == 0 && g_rotate_variant == 3 && g_stream_overlap && streamed_feasible(m)) {
            rc = clean_rot(h);
            if (rc) return rc;
            rc = scan_streamed(h, (size_t)nk, anym != 0, cfg, mode);
            if (rc) return rc;
            tick(h, 5);
            h->last_rows = rows;
            return 0;
        }
        const bool lane_solve = big && g_big_solve_kernel == 0;
        const bool thread_solve = big && !lane_solve && g_rotate_variant == 3;
        if (lane_solve) {
            rc = clean_rot(h);
            if (rc) return rc;
        }
        if (thread_solve) {
            // SNP-minor view rotT[round_up(n,32)][cap_rows]: the sample rows past n must read as zeros
            const size_t n32 = round_up(m.n, 32);
            if (n32 > m.n)
                JXB_CUDA_OK(cudaMemsetAsync(m.rot + m.n * m.cap_rows, 0, (n32 - m.n) * m.cap_rows * sizeof(float), m.stream));
        }
        rc = g_rotate_variant == 3 ? launch_rotate_int8_tc(m, (size_t)nk, anym != 0, thread_solve, m.stream)
                                   : launch_rotate_int8_lib(m, (size_t)nk, anym != 0, m.stream);
        if (rc) return rc;
        tick(h, 4);
        if (lane_solve) {
            const int oc = out_cols_of(cfg, mode);
            h->last_out_cols = oc;
            rc = launch_solve_lane(m, m.rot, m.ldc, (size_t)nk, nullptr, to_params(cfg, mode), m.out, oc, m.evals,
                                   m.n_kept + 1, m.stream);
            note_launch(2);
        } else if (thread_solve) {
            const int oc = out_cols_of(cfg, mode);
            h->last_out_cols = oc;
            rc = launch_solve_thread(m, m.rot, m.cap_rows, (size_t)nk, nullptr, to_params(cfg, mode), m.out, oc, m.evals,
                                     m.stream);
            note_launch(1);
        } else {
            rc = run_solve(h, rows, m.n_kept, cfg, mode);
        }
        if (rc) return rc;
        tick(h, 5);
        h->last_rows = rows;
        // the transposed epilogue may have written beyond column n of the row-major view: restore the zero
        // padding the warp kernel relies on the next time it runs
        if (thread_solve) h->rot_dirty = true;
        return 0;
    } else {
        rc = ensure_capacity(h, rows, 0, true);   // the FP64 path needs the f64 operand block
        if (rc) return rc;
        rc = launch_decode_center(packed_dev, bps, m.src_row, m.n_kept, rows, n_full, sidx_dev, m.n, m.af, m.counts,
                                  qc->genetic_model, m.g64, m.ldk, nullptr, 0, m.stream);
        note_launch(1);
        if (rc) return rc;
        tick(h, 3);
        rc = launch_rotate(m, rows, m.n_kept, m.rot, m.ldc, 0, m.stream, g_rotate_variant >= 2 ? 0 : g_rotate_variant);
        note_launch(1);
        if (rc) return rc;
    }
    tick(h, 4);
    rc = run_solve(h, rows, m.n_kept, cfg, mode);
    if (rc) return rc;
    tick(h, 5);
    h->last_rows = rows;
    return 0;
}

// stage timers of the last scan from the recorded events (call after the stream has been synchronised)
void collect_stage_ms(jxb_model* h) {
    h->stage_ms[7] = h->last_streamed ? 1.f : 0.f;     // reported with or without timers
    if (g_timing && h->timing_ready) {
    for (int i = 0; i < 8; ++i) h->stage_ms[i] = 0.f;
    auto span = [&](int a, int b) -> float {
        float t = 0.f;
        if (h->ev_rec[a] && h->ev_rec[b] && cudaEventElapsedTime(&t, h->ev[a], h->ev[b]) == cudaSuccess) return t;
        (void)cudaGetLastError();  // an unrecorded pair must not poison later launch checks
        return 0.f;
    };
    h->stage_ms[0] = span(1, 2);
    h->stage_ms[1] = span(2, 3);
    h->stage_ms[2] = span(3, 4);
    h->stage_ms[3] = span(4, 5);
    h->stage_ms[4] = span(0, 1);
    h->stage_ms[5] = span(5, 6);
    // streamed scan: [2] = rotation of all slabs (the solve runs underneath from the second slab on), [3] = what is
    // left of the solve after the last slab, [6] = the solve kernel's own duration on its stream, [7] = 1
    h->stage_ms[6] = h->last_streamed ? span(8, 9) : h->stage_ms[3];
    h->stage_ms[7] = h->last_streamed ? 1.f : 0.f;
    for (bool& r : h->ev_rec) r = false;
}
}

int check_scan_args(jxb_model* h, size_t bps, size_t n_full, const int64_t* sidx, const jxb_qc_cfg* qc) {
    if (!h) return fail(-2, "model is null");
    if (!qc) return fail(-2, "qc cfg is null");
    if (!h->m.ut) return fail(-3, "u_t must be (n, n) row-major U^T");
    if (n_full == 0) return fail(-2, "no samples in PLINK FAM");
    if (bps != (n_full + 3) / 4) return fail(-2, "bytes_per_snp must equal ceil(n_full/4)");
    if (!sidx && n_full != h->m.n) return fail(-2, "sample_ids length != expected sample count");
    if (qc->genetic_model < 0 || qc->genetic_model > 3) return fail(-2, "model must be one of: add, dom, rec, het");
    return 0;
}

}  // namespace
}  // namespace jxb

using namespace jxb;

extern "C" {

const char* jxb_last_error(void) { return g_err.c_str(); }

int jxb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* jxb_build_info(void) { return "jxb200 sm_100a cuda-kernels: k1_decode k2_rotate(dmma884+tma) k3_solve"; }

uint64_t jxb_launch_count(void) { return g_launches.load(); }

void jxb_set_timing(int on) { g_timing = on; }
void jxb_set_rotate_variant(int variant) { g_rotate_variant = variant; }
void jxb_set_thread_solve_min_rows(size_t rows) { g_thread_solve_min_rows = rows; }
void jxb_set_big_solve_kernel(int variant) { g_big_solve_kernel = variant == 1 ? 1 : 0; }
void jxb_set_generic_divide(int on) { g_force_generic_divide = on ? 1 : 0; }
void jxb_set_prefix_evals(int on) { g_prefix_evals = on == 2 ? 2 : (on ? 1 : 0); }
void jxb_set_fixed_lane_min_rows(size_t rows) { g_fixed_lane_min_rows = rows; }

int jxb_selftest_rcp(size_t count, int lo_exp, int hi_exp, uint64_t* mismatches) {
    if (!mismatches) return fail(-2, "null argument");
    unsigned long long v = 0;
    int rc = rcp_selftest(count, lo_exp, hi_exp, &v);
    *mismatches = (uint64_t)v;
    return rc;
}

void jxb_set_stream_overlap(int on, size_t slab_rows) {
    g_stream_overlap = on == 2 ? 2 : (on ? 1 : 0);
    if (slab_rows) g_stream_slab_rows = slab_rows;
}

int jxb_model_create(int device, size_t n, size_t p, const double* s, const double* xcov, const double* y,
                     const float* u_t, jxb_model** out) {
    if (!out || !s || !xcov || !y) return fail(-2, "null argument");
    jxb_model* h = nullptr;
    int rc = model_alloc(device, n, p, u_t != nullptr, &h);
    if (rc) return rc;
    rc = upload_small(h->m, s, xcov, y);
    if (!rc && u_t) rc = widen_ut(h->m, u_t, false);
    if (rc) { jxb_model_destroy(h); return rc; }
    *out = h;
    return 0;
}

int jxb_model_create_dev(int device, size_t n, size_t p, const double* s_dev, const double* xcov_dev,
                         const double* y_dev, const float* u_t_dev, jxb_model** out) {
    if (!out || !s_dev || !xcov_dev || !y_dev) return fail(-2, "null argument");
    jxb_model* h = nullptr;
    int rc = model_alloc(device, n, p, u_t_dev != nullptr, &h);
    if (rc) return rc;
    std::vector<double> s(n), x(n * p), y(n);
    cudaError_t e = cudaMemcpy(s.data(), s_dev, n * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(x.data(), xcov_dev, n * p * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(y.data(), y_dev, n * sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { jxb_model_destroy(h); return fail(-100, cudaGetErrorString(e)); }
    rc = upload_small(h->m, s.data(), x.data(), y.data());
    if (!rc && u_t_dev) rc = widen_ut(h->m, u_t_dev, true);
    if (rc) { jxb_model_destroy(h); return rc; }
    *out = h;
    return 0;
}

void jxb_model_destroy(jxb_model* h) {
    if (!h) return;
    Model& m = h->m;
    cudaSetDevice(m.device);
    if (m.stream) cudaStreamSynchronize(m.stream);
    void* ptrs[] = {m.s, m.y, m.xt, m.rec, m.ut, m.g64, m.rot, m.log_table, m.ssq, m.prefix_buf, m.out, m.evals, m.packed, m.counts, m.af, m.src_row,
                    m.n_kept, m.sample_idx, m.stage_f32, m.fx_w, m.fx_py, m.fx_wx, m.fx_scal, m.fx_rec, h->missr, h->mask,
                    h->prep_af, h->prep_flip,
                    h->scal, m.q8, m.q8_inv_scale, m.q8_rk, m.a8, m.coef, m.flags8, m.c32, m.lt_ws, m.corr64};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (h->timing_ready)
        for (auto& e : h->ev) cudaEventDestroy(e);
    if (m.stream) cudaStreamDestroy(m.stream);
    if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    if (h->ev_copy) cudaEventDestroy(h->ev_copy);
    if (h->packed2) cudaFree(h->packed2);
    if (h->ev_slab0) cudaEventDestroy(h->ev_slab0);
    if (h->ev_solve) cudaEventDestroy(h->ev_solve);
    for (void* t : {m.tmap_a8, m.tmap_q8_7, m.tmap_q8_3})
        if (t) free(t);
    if (m.tmap_ut) free(m.tmap_ut);
    if (m.tmap_g) free(m.tmap_g);
    delete h;
}

int jxb_model_set_xy(jxb_model* h, const double* xcov, const double* y) {
    if (!h || !xcov || !y) return fail(-2, "null argument");
    JXB_CUDA_OK(cudaSetDevice(h->m.device));
    return upload_small(h->m, nullptr, xcov, y);
}

int jxb_model_sync(jxb_model* h) {
    if (!h) return fail(-2, "model is null");
    JXB_CUDA_OK(cudaSetDevice(h->m.device));
    JXB_CUDA_OK(cudaStreamSynchronize(h->m.stream));
    return 0;
}

void* jxb_model_stream(jxb_model* h) { return h ? (void*)h->m.stream : nullptr; }

int jxb_rotate_xy(jxb_model* h, const double* x, size_t q, const double* y, double* x_rot, double* y_rot) {
    if (!h || !x || !y || !x_rot || !y_rot) return fail(-2, "null argument");
    Model& m = h->m;
    if (!m.ut) return fail(-3, "u_t must be shape (n, n) and row-major U^T");
    JXB_CUDA_OK(cudaSetDevice(m.device));
    const size_t n = m.n;
    double *dx = nullptr, *dy = nullptr, *dxr = nullptr, *dyr = nullptr;
    JXB_CUDA_OK(cudaMalloc((void**)&dx, std::max<size_t>(1, n * q) * sizeof(double)));
    JXB_CUDA_OK(cudaMalloc((void**)&dy, n * sizeof(double)));
    JXB_CUDA_OK(cudaMalloc((void**)&dxr, std::max<size_t>(1, n * q) * sizeof(double)));
    JXB_CUDA_OK(cudaMalloc((void**)&dyr, n * sizeof(double)));
    int rc = 0;
    cudaMemcpyAsync(dx, x, n * q * sizeof(double), cudaMemcpyHostToDevice, m.stream);
    cudaMemcpyAsync(dy, y, n * sizeof(double), cudaMemcpyHostToDevice, m.stream);
    rc = launch_rotate_xy(m, nullptr, dx, q, dy, dxr, dyr, m.stream);
    note_launch(1);
    if (!rc) {
        cudaMemcpyAsync(x_rot, dxr, n * q * sizeof(double), cudaMemcpyDeviceToHost, m.stream);
        cudaMemcpyAsync(y_rot, dyr, n * sizeof(double), cudaMemcpyDeviceToHost, m.stream);
        cudaError_t e = cudaStreamSynchronize(m.stream);
        if (e != cudaSuccess) rc = fail(-100, cudaGetErrorString(e));
    }
    cudaFree(dx); cudaFree(dy); cudaFree(dxr); cudaFree(dyr);
    return rc;
}

static int null_common(jxb_model* h, int kind, double low, double high, int max_iter, double tol, int has_init,
                       double init, double* out, int n_out) {
    if (!h || !out) return fail(-2, "null argument");
    Model& m = h->m;
    JXB_CUDA_OK(cudaSetDevice(m.device));
    int rc = launch_null_fit(m, kind, low, high, max_iter, tol, has_init, init, h->scal, m.stream);
    note_launch(1);
    if (rc) return rc;
    JXB_CUDA_OK(cudaMemcpyAsync(out, h->scal, n_out * sizeof(double), cudaMemcpyDeviceToHost, m.stream));
    JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
    return 0;
}

int jxb_reml_null(jxb_model* h, double low, double high, int max_iter, double tol, double out3[3]) {
    if (!(low < high)) return fail(-2, "low must be < high");
    return null_common(h, 0, low, high, max_iter, tol, 0, 0.0, out3, 3);
}

int jxb_ml_loglike_null(jxb_model* h, double log10_lbd, double* ml) {
    return null_common(h, 2, 0, 0, 0, 0, 1, log10_lbd, ml, 1);
}

int jxb_ml_null(jxb_model* h, double low, double high, int max_iter, double tol, int has_init, double init,
                double out2[2]) {
    if (!(low < high)) return fail(-2, "low must be < high");
    return null_common(h, 1, low, high, max_iter, tol, has_init, init, out2, 2);
}

int jxb_lmm_reml_chunk_f32(jxb_model* h, const float* g_rot, size_t rows, const jxb_solve_cfg* cfg, double* out,
                           int32_t* evals) {
    return chunk_common(h, g_rot, rows, true, cfg, 0, out, evals);
}

int jxb_lmm_reml_chunk_from_snp_f32(jxb_model* h, const float* snp, size_t rows, const jxb_solve_cfg* cfg,
                                    double* out, int32_t* evals) {
    return chunk_common(h, snp, rows, false, cfg, 0, out, evals);
}

int jxb_lmm2_chunk_f32(jxb_model* h, const float* snp, size_t rows, int rotated, const jxb_solve_cfg* cfg,
                       double* out, int32_t* evals) {
    return chunk_common(h, snp, rows, rotated != 0, cfg, 1, out, evals);
}

int jxb_lmm_fixed_chunk_f32(jxb_model* h, const float* snp, size_t rows, int rotated, double log10_lbd,
                            const double* nullml, double* out, double meta3[3]) {
    if (!h) return fail(-2, "model is null");
    if (h->m.n <= h->m.p + 1) return fail(-2, "n must be > p_cov+1");
    jxb_solve_cfg cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.has_init = 1;
    cfg.init_log10_lbd = log10_lbd;
    cfg.has_nullml = nullml ? 1 : 0;
    cfg.nullml = nullml ? *nullml : 0.0;
    int rc = chunk_common(h, snp, rows, rotated != 0, &cfg, 2, out, nullptr);
    if (rc) return rc;
    if (meta3 && rows > 0) {
        JXB_CUDA_OK(cudaMemcpy(meta3, h->m.fx_scal, 3 * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return 0;
}

int jxb_rotate_block_f32(jxb_model* h, const float* snp, size_t rows, float* rot_host, int variant) {
    if (!h || !snp || !rot_host) return fail(-2, "null argument");
    Model& m = h->m;
    if (!m.ut) return fail(-3, "u_t must be (n, n) and row-major U^T");
    if (rows == 0) return 0;
    int rc = ensure_capacity(h, rows, 0, true);
    if (rc) return rc;
    rc = clean_rot(h);
    if (rc) return rc;
    const size_t n = m.n;
    rc = ensure_stage_f32(m, rows * n);
    if (rc) return rc;
    JXB_CUDA_OK(cudaMemcpyAsync(m.stage_f32, snp, rows * n * sizeof(float), cudaMemcpyHostToDevice, m.stream));
    rc = launch_widen_f32(m.stage_f32, n, rows, n, m.g64, m.ldk, m.stream);
    note_launch(1);
    if (rc) return rc;
    rc = launch_rotate(m, rows, nullptr, m.rot, m.ldc, 0, m.stream, variant);
    note_launch(1);
    if (rc) return rc;
    JXB_CUDA_OK(cudaMemcpy2DAsync(rot_host, n * sizeof(float), m.rot, m.ldc * sizeof(float), n * sizeof(float), rows,
                                  cudaMemcpyDeviceToHost, m.stream));
    JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
    return 0;
}

int jxb_scan_packed_dev(jxb_model* h, const uint8_t* packed_dev, size_t bps, size_t rows, size_t n_full,
                        const int64_t* sidx_dev, const jxb_qc_cfg* qc, const jxb_solve_cfg* cfg, int mode) {
    int rc = check_scan_args(h, bps, n_full, sidx_dev, qc);
    if (rc) return rc;
    rc = check_cfg(cfg, mode);
    if (rc) return rc;
    if (rows == 0) { h->last_rows = 0; return 0; }
    rc = ensure_capacity(h, rows, 0, false);
    if (rc) return rc;
    return scan_device_stages(h, packed_dev, bps, rows, n_full, sidx_dev, h->m.n, false, qc, cfg, mode);
}

int jxb_scan_fetch(jxb_model* h, size_t rows, int out_cols, uint8_t* keep_host, float* af_host, int32_t* missing_host,
                   double* out_host, int32_t* evals_host, size_t* n_kept_host) {
    if (!h) return fail(-2, "model is null");
    Model& m = h->m;
    JXB_CUDA_OK(cudaSetDevice(m.device));
    if (rows == 0) { if (n_kept_host) *n_kept_host = 0; return 0; }
    if (rows > m.cap_rows) return fail(-2, "rows exceeds the last scan");
    std::vector<int32_t> counts(rows * 4);
    int32_t scal[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    JXB_CUDA_OK(cudaMemcpyAsync(counts.data(), m.counts, rows * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, m.stream));
    JXB_CUDA_OK(cudaMemcpyAsync(scal, m.n_kept, sizeof scal, cudaMemcpyDeviceToHost, m.stream));
    if (af_host) JXB_CUDA_OK(cudaMemcpyAsync(af_host, m.af, rows * sizeof(float), cudaMemcpyDeviceToHost, m.stream));
    JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
    const int32_t nk = scal[0];
    if (h->last_streamed && scal[6] != 0)
        return fail(-50, "streamed scan: the solve kernel gave up waiting for rotated rows (rotation kernel not co-resident); "
                         "disable the overlap with jxb_set_stream_overlap(0)");
    if (out_host && nk > 0)
        JXB_CUDA_OK(cudaMemcpyAsync(out_host, m.out, (size_t)nk * out_cols * sizeof(double), cudaMemcpyDeviceToHost, m.stream));
    if (evals_host && nk > 0)
        JXB_CUDA_OK(cudaMemcpyAsync(evals_host, m.evals, (size_t)nk * sizeof(int32_t), cudaMemcpyDeviceToHost, m.stream));
    tick(h, 6);
    JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
    for (size_t r = 0; r < rows; ++r) {
        if (keep_host) keep_host[r] = (uint8_t)(counts[4 * r + 3] != 0);
        if (missing_host) missing_host[r] = counts[4 * r + 0];
    }
    if (n_kept_host) *n_kept_host = (size_t)nk;
    collect_stage_ms(h);
    return 0;
}

int jxb_scan_fetch_dev(jxb_model* h, size_t rows, int out_cols, double* out_dst_dev, float* af_dst_dev,
                       int32_t* counts_dst_dev, size_t* n_kept_host) {
    if (!h) return fail(-2, "model is null");
    Model& m = h->m;
    JXB_CUDA_OK(cudaSetDevice(m.device));
    if (rows == 0) { if (n_kept_host) *n_kept_host = 0; return 0; }
    if (rows > m.cap_rows) return fail(-2, "rows exceeds the last scan");
    int32_t scal[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    JXB_CUDA_OK(cudaMemcpyAsync(scal, m.n_kept, sizeof scal, cudaMemcpyDeviceToHost, m.stream));
    if (h->last_nk == (size_t)-1) {
        JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
        h->last_nk = (size_t)scal[0];
    }
    const size_t nk_known = h->last_nk;
    if (out_dst_dev && nk_known > 0)
        JXB_CUDA_OK(cudaMemcpyAsync(out_dst_dev, m.out, nk_known * out_cols * sizeof(double), cudaMemcpyDeviceToDevice, m.stream));
    if (af_dst_dev) JXB_CUDA_OK(cudaMemcpyAsync(af_dst_dev, m.af, rows * sizeof(float), cudaMemcpyDeviceToDevice, m.stream));
    if (counts_dst_dev)
        JXB_CUDA_OK(cudaMemcpyAsync(counts_dst_dev, m.counts, rows * 4 * sizeof(int32_t), cudaMemcpyDeviceToDevice, m.stream));
    tick(h, 6);
    JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
    if (h->last_streamed && scal[6] != 0)
        return fail(-50, "streamed scan: the solve kernel gave up waiting for rotated rows (rotation kernel not co-resident); "
                         "disable the overlap with jxb_set_stream_overlap(0)");
    if ((size_t)scal[0] != nk_known) return fail(-51, "internal error: kept-row count changed between scan and fetch");
    if (n_kept_host) *n_kept_host = nk_known;
    collect_stage_ms(h);
    return 0;
}

int jxb_scan_packed_prepared(jxb_model* h, const uint8_t* packed, size_t bps, size_t rows, size_t n_full,
                             const int64_t* sidx, const uint8_t* pre_keep, const float* row_af, const uint8_t* row_flip,
                             const jxb_qc_cfg* qc, const jxb_solve_cfg* cfg, int mode, uint8_t* keep_host, float* af_host,
                             int32_t* missing_host, double* out_host, int32_t* evals_host, size_t* n_kept_host) {
    int rc = check_scan_args(h, bps, n_full, sidx, qc);
    if (rc) return rc;
    rc = check_cfg(cfg, mode);
    if (rc) return rc;
    if (rows == 0) { if (n_kept_host) *n_kept_host = 0; return 0; }
    if (!packed) return fail(-2, "packed is null");
    Model& m = h->m;
    rc = ensure_capacity(h, rows, bps, false);
    if (rc) return rc;
    tick(h, 0);
    JXB_CUDA_OK(cudaMemcpyAsync(m.packed, packed, rows * bps, cudaMemcpyHostToDevice, m.stream));
    const int64_t* sidx_dev = nullptr;
    if (sidx) {
        rc = ensure_sample_idx(m, sidx, m.n, false, n_full);
        if (rc) return rc;
        sidx_dev = m.sample_idx;
    }
    if (pre_keep) JXB_CUDA_OK(cudaMemcpyAsync(h->mask, pre_keep, rows, cudaMemcpyHostToDevice, m.stream));
    if (row_af) JXB_CUDA_OK(cudaMemcpyAsync(h->prep_af, row_af, rows * sizeof(float), cudaMemcpyHostToDevice, m.stream));
    if (row_flip) JXB_CUDA_OK(cudaMemcpyAsync(h->prep_flip, row_flip, rows, cudaMemcpyHostToDevice, m.stream));
    rc = scan_device_stages(h, m.packed, bps, rows, n_full, sidx_dev, m.n, pre_keep != nullptr, qc, cfg, mode,
                            row_af ? h->prep_af : nullptr, row_flip ? h->prep_flip : nullptr);
    if (rc) return rc;
    return jxb_scan_fetch(h, rows, h->last_out_cols, keep_host, af_host, missing_host, out_host, evals_host,
                          n_kept_host);
}

int jxb_stage_packed(jxb_model* h, const uint8_t* packed_host, size_t bps, size_t rows) {
    if (!h || !packed_host) return fail(-2, "null argument");
    if (rows == 0) return fail(-2, "cannot stage an empty batch");
    Model& m = h->m;
    if (h->staged) return fail(-2, "a staged batch is already waiting: run jxb_scan_staged_begin first");
    if (h->cur_valid && (rows > m.cap_rows || bps > m.bps_cap))
        return fail(-2, "cannot stage a batch larger than the workspace while another batch is current: stage the largest batch first");
    int rc = ensure_capacity(h, rows, bps, false);
    if (rc) return rc;
    if (!h->copy_stream) {
        JXB_CUDA_OK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        JXB_CUDA_OK(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
    }
    const size_t need = m.cap_rows * m.bps_cap;
    if (h->packed2_bytes < need) {
        if (h->packed2) cudaFree(h->packed2);
        h->packed2 = nullptr;
        h->packed2_bytes = 0;
        JXB_CUDA_OK(cudaMalloc((void**)&h->packed2, need));
        h->packed2_bytes = need;
    }
    JXB_CUDA_OK(cudaMemcpyAsync(h->packed2, packed_host, rows * bps, cudaMemcpyHostToDevice, h->copy_stream));
    JXB_CUDA_OK(cudaEventRecord(h->ev_copy, h->copy_stream));
    h->staged = true;
    h->staged_rows = rows;
    h->staged_bps = bps;
    return 0;
}

int jxb_scan_staged_begin(jxb_model* h) {
    if (!h) return fail(-2, "model is null");
    if (!h->staged) return fail(-2, "no staged batch: call jxb_stage_packed first");
    Model& m = h->m;
    JXB_CUDA_OK(cudaSetDevice(m.device));
    // the staged copy becomes the working buffer; the previous working buffer is the next staging target
    JXB_CUDA_OK(cudaStreamWaitEvent(m.stream, h->ev_copy, 0));
    std::swap(m.packed, h->packed2);
    h->cur_rows = h->staged_rows;
    h->cur_bps = h->staged_bps;
    h->staged = false;
    h->cur_valid = true;
    return 0;
}

void jxb_stage_cancel(jxb_model* h) {
    if (!h) return;
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    h->staged = false;
    h->cur_valid = false;
}

int jxb_scan_staged(jxb_model* h, size_t n_full, const int64_t* sidx, const uint8_t* pre_keep, const float* row_af,
                    const uint8_t* row_flip, const jxb_qc_cfg* qc, const jxb_solve_cfg* cfg, int mode, uint8_t* keep_host,
                    float* af_host, int32_t* missing_host, double* out_host, int32_t* evals_host, size_t* n_kept_host) {
    if (!h) return fail(-2, "model is null");
    if (!h->cur_valid) return fail(-2, "no current batch: call jxb_stage_packed and jxb_scan_staged_begin first");
    const size_t rows = h->cur_rows, bps = h->cur_bps;
    h->cur_valid = false;
    int rc = check_scan_args(h, bps, n_full, sidx, qc);
    if (!rc) rc = check_cfg(cfg, mode);
    if (rc) return rc;
    Model& m = h->m;
    if (rows > m.cap_rows || bps > m.bps_cap) return fail(-2, "internal error: staged batch exceeds the workspace");
    tick(h, 0);
    const int64_t* sidx_dev = nullptr;
    if (sidx) {
        rc = ensure_sample_idx(m, sidx, m.n, false, n_full);
        if (rc) return rc;
        sidx_dev = m.sample_idx;
    }
    if (pre_keep) JXB_CUDA_OK(cudaMemcpyAsync(h->mask, pre_keep, rows, cudaMemcpyHostToDevice, m.stream));
    if (row_af) JXB_CUDA_OK(cudaMemcpyAsync(h->prep_af, row_af, rows * sizeof(float), cudaMemcpyHostToDevice, m.stream));
    if (row_flip) JXB_CUDA_OK(cudaMemcpyAsync(h->prep_flip, row_flip, rows, cudaMemcpyHostToDevice, m.stream));
    rc = scan_device_stages(h, m.packed, bps, rows, n_full, sidx_dev, m.n, pre_keep != nullptr, qc, cfg, mode,
                            row_af ? h->prep_af : nullptr, row_flip ? h->prep_flip : nullptr);
    if (rc) return rc;
    return jxb_scan_fetch(h, rows, h->last_out_cols, keep_host, af_host, missing_host, out_host, evals_host, n_kept_host);
}

int jxb_scan_packed(jxb_model* h, const uint8_t* packed, size_t bps, size_t rows, size_t n_full,
                    const int64_t* sidx, const uint8_t* pre_keep, const jxb_qc_cfg* qc, const jxb_solve_cfg* cfg,
                    int mode, uint8_t* keep_host, float* af_host, int32_t* missing_host, double* out_host,
                    int32_t* evals_host, size_t* n_kept_host) {
    return jxb_scan_packed_prepared(h, packed, bps, rows, n_full, sidx, pre_keep, nullptr, nullptr, qc, cfg, mode, keep_host,
                                    af_host, missing_host, out_host, evals_host, n_kept_host);
}

int jxb_decode_packed(jxb_model* h, const uint8_t* packed, size_t bps, size_t rows, size_t n_full,
                      const int64_t* sidx, const jxb_qc_cfg* qc, int32_t* counts_host, float* af_host,
                      float* miss_rate_host, float* g_host, size_t* n_kept_host) {
    if (!h || !qc || !packed) return fail(-2, "null argument");
    if (bps != (n_full + 3) / 4) return fail(-2, "bytes_per_snp must equal ceil(n_full/4)");
    if (!sidx && n_full != h->m.n) return fail(-2, "sample_ids length != expected sample count");
    if (rows == 0) { if (n_kept_host) *n_kept_host = 0; return 0; }
    Model& m = h->m;
    int rc = ensure_capacity(h, rows, bps, false);
    if (rc) return rc;
    const size_t n = m.n;
    JXB_CUDA_OK(cudaMemcpyAsync(m.packed, packed, rows * bps, cudaMemcpyHostToDevice, m.stream));
    const int64_t* sidx_dev = nullptr;
    if (sidx) {
        rc = ensure_sample_idx(m, sidx, n, false, n_full);
        if (rc) return rc;
        sidx_dev = m.sample_idx;
    }
    rc = launch_count_qc(m, m.packed, bps, rows, n_full, sidx_dev, n, qc->maf_thr, qc->miss_thr, qc->het_thr, m.counts,
                         m.af, h->missr, m.stream);
    note_launch(1);
    if (rc) return rc;
    rc = launch_compact(m.counts, rows, m.src_row, m.n_kept, m.stream);
    note_launch(1);
    if (rc) return rc;
    int32_t nk = 0;
    JXB_CUDA_OK(cudaMemcpyAsync(&nk, m.n_kept, sizeof nk, cudaMemcpyDeviceToHost, m.stream));
    if (counts_host)
        JXB_CUDA_OK(cudaMemcpyAsync(counts_host, m.counts, rows * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, m.stream));
    if (af_host) JXB_CUDA_OK(cudaMemcpyAsync(af_host, m.af, rows * sizeof(float), cudaMemcpyDeviceToHost, m.stream));
    if (miss_rate_host)
        JXB_CUDA_OK(cudaMemcpyAsync(miss_rate_host, h->missr, rows * sizeof(float), cudaMemcpyDeviceToHost, m.stream));
    JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
    if (n_kept_host) *n_kept_host = (size_t)nk;
    if (g_host && nk > 0) {
        rc = ensure_stage_f32(m, (size_t)nk * n);
        if (rc) return rc;
        rc = launch_decode_center(m.packed, bps, m.src_row, m.n_kept, rows, n_full, sidx_dev, n, m.af, m.counts,
                                  qc->genetic_model, nullptr, 0, m.stage_f32, n, m.stream);
        note_launch(1);
        if (rc) return rc;
        JXB_CUDA_OK(cudaMemcpyAsync(g_host, m.stage_f32, (size_t)nk * n * sizeof(float), cudaMemcpyDeviceToHost, m.stream));
        JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
    }
    return 0;
}

int jxb_decode_packed_prepared(jxb_model* h, const uint8_t* packed, size_t bps, size_t rows, size_t n_full,
                               const int64_t* sidx, const uint8_t* keep_host, const float* af_host, int genetic_model,
                               float* g_host, size_t* n_kept_host) {
    if (!h || !packed || !keep_host || !af_host || !g_host) return fail(-2, "null argument");
    if (bps != (n_full + 3) / 4) return fail(-2, "bytes_per_snp must equal ceil(n_full/4)");
    if (!sidx && n_full != h->m.n) return fail(-2, "sample_ids length != expected sample count");
    if (genetic_model < 0 || genetic_model > 3) return fail(-2, "model must be one of: add, dom, rec, het");
    if (rows == 0) { if (n_kept_host) *n_kept_host = 0; return 0; }
    Model& m = h->m;
    int rc = ensure_capacity(h, rows, bps, false);
    if (rc) return rc;
    const size_t n = m.n;
    JXB_CUDA_OK(cudaMemcpyAsync(m.packed, packed, rows * bps, cudaMemcpyHostToDevice, m.stream));
    const int64_t* sidx_dev = nullptr;
    if (sidx) {
        rc = ensure_sample_idx(m, sidx, n, false, n_full);
        if (rc) return rc;
        sidx_dev = m.sample_idx;
    }
    // genotype counts feed the closed-form row mean; the caller's keep flags and allele frequencies replace the QC
    rc = launch_count_qc(m, m.packed, bps, rows, n_full, sidx_dev, n, 0.0f, 1.0f, 0.0f, m.counts, m.af, h->missr, m.stream);
    note_launch(1);
    if (rc) return rc;
    JXB_CUDA_OK(cudaMemcpyAsync(h->mask, keep_host, rows, cudaMemcpyHostToDevice, m.stream));
    JXB_CUDA_OK(cudaMemcpyAsync(m.af, af_host, rows * sizeof(float), cudaMemcpyHostToDevice, m.stream));
    apply_mask_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, m.stream>>>(m.counts, h->mask, (int)rows);
    note_launch(1);
    rc = launch_compact(m.counts, rows, m.src_row, m.n_kept, m.stream);
    note_launch(1);
    if (rc) return rc;
    int32_t nk = 0;
    JXB_CUDA_OK(cudaMemcpyAsync(&nk, m.n_kept, sizeof nk, cudaMemcpyDeviceToHost, m.stream));
    JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
    if (n_kept_host) *n_kept_host = (size_t)nk;
    if (nk > 0) {
        rc = ensure_stage_f32(m, (size_t)nk * n);
        if (rc) return rc;
        rc = launch_decode_center(m.packed, bps, m.src_row, m.n_kept, rows, n_full, sidx_dev, n, m.af, m.counts,
                                  genetic_model, nullptr, 0, m.stage_f32, n, m.stream);
        note_launch(1);
        if (rc) return rc;
        JXB_CUDA_OK(cudaMemcpyAsync(g_host, m.stage_f32, (size_t)nk * n * sizeof(float), cudaMemcpyDeviceToHost, m.stream));
        JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
    }
    return 0;
}

int jxb_decode_packed_lut(jxb_model* h, const uint8_t* packed, size_t bps, size_t rows, size_t n_full, const int64_t* sidx,
                          const float* row_lut_host, float* g_host) {
    if (!h || !packed || !row_lut_host || !g_host) return fail(-2, "null argument");
    if (bps != (n_full + 3) / 4) return fail(-2, "bytes_per_snp must equal ceil(n_full/4)");
    if (!sidx && n_full != h->m.n) return fail(-2, "sample_ids length != expected sample count");
    if (rows == 0) return 0;
    Model& m = h->m;
    int rc = ensure_capacity(h, rows, bps, false);
    if (rc) return rc;
    const size_t n = m.n;
    JXB_CUDA_OK(cudaMemcpyAsync(m.packed, packed, rows * bps, cudaMemcpyHostToDevice, m.stream));
    const int64_t* sidx_dev = nullptr;
    if (sidx) {
        rc = ensure_sample_idx(m, sidx, n, false, n_full);
        if (rc) return rc;
        sidx_dev = m.sample_idx;
    }
    // the 4-float LUT rows travel in the out buffer's head (f64[cap][8] = room for 16 floats per row)
    float* lut_dev = reinterpret_cast<float*>(m.out);
    JXB_CUDA_OK(cudaMemcpyAsync(lut_dev, row_lut_host, rows * 4 * sizeof(float), cudaMemcpyHostToDevice, m.stream));
    rc = ensure_stage_f32(m, rows * n);
    if (rc) return rc;
    rc = launch_decode_center(m.packed, bps, nullptr, nullptr, rows, n_full, sidx_dev, n, m.af, m.counts, 0, nullptr, 0,
                              m.stage_f32, n, m.stream, lut_dev);
    note_launch(1);
    if (rc) return rc;
    JXB_CUDA_OK(cudaMemcpyAsync(g_host, m.stage_f32, rows * n * sizeof(float), cudaMemcpyDeviceToHost, m.stream));
    JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
    return 0;
}

int jxb_debug_fetch_rot(jxb_model* h, size_t row0, size_t rows, float* rot_host) {
    if (!h || !rot_host) return fail(-2, "null argument");
    Model& m = h->m;
    if (!m.rot || row0 + rows > m.cap_rows) return fail(-2, "rows exceed the workspace");
    JXB_CUDA_OK(cudaSetDevice(m.device));
    JXB_CUDA_OK(cudaMemcpy2DAsync(rot_host, m.ldc * sizeof(float), m.rot + row0 * m.ldc, m.ldc * sizeof(float),
                                  m.ldc * sizeof(float), rows, cudaMemcpyDeviceToHost, m.stream));
    JXB_CUDA_OK(cudaStreamSynchronize(m.stream));
    return 0;
}

int jxb_last_stage_ms(jxb_model* h, float ms6[6]) {
    if (!h) return fail(-2, "model is null");
    for (int i = 0; i < 6; ++i) ms6[i] = h->stage_ms[i];
    return 0;
}

int jxb_last_stage_ms8(jxb_model* h, float ms8[8]) {
    if (!h) return fail(-2, "model is null");
    for (int i = 0; i < 8; ++i) ms8[i] = h->stage_ms[i];
    return 0;
}

}  // extern "C"
