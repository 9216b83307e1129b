// eigh.cu -- SURVEY 8(f) N2 / row A17: spectral decomposition of the GRM on the device.
//
// Replaces rust_eigh_from_array_f64[_inplace] (src/math/eigh.rs:1621-1705, 1883-1990; LAPACK dsyevd/dsyevr drivers
// :1201-1420) with ONE cuSOLVER library call (cusolverDnXsyevd, 64-bit API), as SURVEY A17 prescribes ("cuSOLVER
// library call, not a hand kernel").  The library is opened with dlopen so libjxb200.so has no link-time dependency
// on it.  Output convention: eigenvalues ascending; the matrix is overwritten by U^T row-major (row k = k-th
// eigenvector) -- cuSOLVER's column-major eigenvector matrix read as row-major IS U^T, the `Dh = U.T` the scan takes.
//
// cusolverDnXsyevd rejects n*n >= 2^31 (n > 46,340).  Above that -- BASELINE configs[3], n = 50,000; the reference
// switches its LAPACK driver there too (dsyevr for n >= 32,768, assoc/workflow.py:5462-5486) -- and whenever the caller
// asks for several devices (jxb_set_eigh_devices), the decomposition runs through cusolverMgSyevd: the symmetric
// matrix is dealt in 1-D block-cyclic column panels over the GPUs of the box (peer copies over NVLink), decomposed
// there, and the eigenvector panels are collected back into the caller's buffer.
#include <cusolverDn.h>
#include <cusolverMg.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/jxb200.h"
#include "jxb_common.cuh"

namespace jxb {
namespace {

struct Solver {
    void* so = nullptr;
    bool tried = false, ok = false;
    decltype(&cusolverDnCreate) create = nullptr;
    decltype(&cusolverDnDestroy) destroy = nullptr;
    decltype(&cusolverDnSetStream) set_stream = nullptr;
    decltype(&cusolverDnCreateParams) create_params = nullptr;
    decltype(&cusolverDnDestroyParams) destroy_params = nullptr;
    decltype(&cusolverDnXsyevd_bufferSize) buffer_size = nullptr;
    decltype(&cusolverDnXsyevd) syevd = nullptr;
    decltype(&cusolverDnSetDeterministicMode) set_det = nullptr;   // optional (CUDA >= 12.2)
};

Solver& solver() {
    static Solver g;
    if (g.tried) return g;
    g.tried = true;
    for (const char* name : {"libcusolver.so.11", "libcusolver.so.12", "libcusolver.so"}) {
        g.so = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (g.so) break;
    }
    if (!g.so) return g;
#define JXB_SYM(field, sym) g.field = (decltype(g.field))dlsym(g.so, #sym); if (!g.field) return g;
    JXB_SYM(create, cusolverDnCreate)
    JXB_SYM(destroy, cusolverDnDestroy)
    JXB_SYM(set_stream, cusolverDnSetStream)
    JXB_SYM(create_params, cusolverDnCreateParams)
    JXB_SYM(destroy_params, cusolverDnDestroyParams)
    JXB_SYM(buffer_size, cusolverDnXsyevd_bufferSize)
    JXB_SYM(syevd, cusolverDnXsyevd)
#undef JXB_SYM
    g.set_det = (decltype(g.set_det))dlsym(g.so, "cusolverDnSetDeterministicMode");
    g.ok = true;
    return g;
}

__global__ void add_diag_kernel(double* a, size_t n, double shift) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i * n + i] += shift;
}

__global__ void narrow_kernel(const double* __restrict__ src, float* __restrict__ dst, size_t count) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) dst[i] = (float)src[i];
}


// ---- cusolverMg (multi-GPU / large n) -----------------------------------------------------------------------------
struct SolverMg {
    void* so = nullptr;
    std::string err;
    bool tried = false, ok = false;
    decltype(&cusolverMgCreate) create = nullptr;
    decltype(&cusolverMgDestroy) destroy = nullptr;
    decltype(&cusolverMgDeviceSelect) device_select = nullptr;
    decltype(&cusolverMgCreateDeviceGrid) create_grid = nullptr;
    decltype(&cusolverMgDestroyGrid) destroy_grid = nullptr;
    decltype(&cusolverMgCreateMatrixDesc) create_desc = nullptr;
    decltype(&cusolverMgDestroyMatrixDesc) destroy_desc = nullptr;
    decltype(&cusolverMgSyevd_bufferSize) buffer_size = nullptr;
    decltype(&cusolverMgSyevd) syevd = nullptr;
};

SolverMg& solver_mg() {
    static SolverMg g;
    if (g.tried) return g;
    g.tried = true;
    // first choice: the libcusolverMg that sits next to the libcusolver already in the process (a Python process that
    // imported torch holds torch's bundled copy; the toolkit's Mg library may not resolve against it)
    Solver& dn = solver();
    Dl_info info;
    if (dn.ok && dladdr((void*)dn.create, &info) && info.dli_fname) {
        std::string dir(info.dli_fname);
        const size_t slash = dir.find_last_of('/');
        if (slash != std::string::npos) {
            dir.resize(slash + 1);
            for (const char* name : {"libcusolverMg.so.11", "libcusolverMg.so.12", "libcusolverMg.so"}) {
                g.so = dlopen((dir + name).c_str(), RTLD_NOW | RTLD_GLOBAL);
                if (g.so) break;
            }
        }
    }
    for (const char* name : {"libcusolverMg.so.11", "libcusolverMg.so.12", "libcusolverMg.so"}) {
        if (g.so) break;
        g.so = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (!g.so) { const char* e = dlerror(); if (e) g.err = e; }
    }
    if (!g.so) return g;
#define JXB_SYM(field, sym) g.field = (decltype(g.field))dlsym(g.so, #sym); if (!g.field) return g;
    JXB_SYM(create, cusolverMgCreate)
    JXB_SYM(destroy, cusolverMgDestroy)
    JXB_SYM(device_select, cusolverMgDeviceSelect)
    JXB_SYM(create_grid, cusolverMgCreateDeviceGrid)
    JXB_SYM(destroy_grid, cusolverMgDestroyGrid)
    JXB_SYM(create_desc, cusolverMgCreateMatrixDesc)
    JXB_SYM(destroy_desc, cusolverMgDestroyMatrixDesc)
    JXB_SYM(buffer_size, cusolverMgSyevd_bufferSize)
    JXB_SYM(syevd, cusolverMgSyevd)
#undef JXB_SYM
    g.ok = true;
    return g;
}

int g_eigh_devices = 0;   // 0 = one device unless n forces the Mg path; k > 1 = use the first k visible devices

// a_dev: n x n symmetric f64 on `device` (overwritten by U^T row-major).  Devices used: `device` first, then the other
// visible devices in index order, `ndev` in total.
int eigh_device_mg(int device, int ndev, size_t n, double* a_dev, double diag_shift, double* evals_dev, float* ut_f32_dev,
                   cudaStream_t st) {
    SolverMg& s = solver_mg();
    if (!s.ok)
        return fail(-110, "libcusolverMg (cusolverMgSyevd) could not be loaded" + (s.err.empty() ? std::string() : " (" + s.err + ")") +
                              ": no eigensolver for n > 46340 / multi-GPU");
    int visible = 0;
    JXB_CUDA_OK(cudaGetDeviceCount(&visible));
    ndev = std::max(1, std::min(ndev, visible));
    std::vector<int> devs{device};
    for (int d = 0; d < visible && (int)devs.size() < ndev; ++d)
        if (d != device) devs.push_back(d);
    JXB_CUDA_OK(cudaSetDevice(device));
    if (diag_shift != 0.0) {
        add_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a_dev, n, diag_shift);
        note_launch(1);
        JXB_CUDA_OK(cudaGetLastError());
    }
    JXB_CUDA_OK(cudaStreamSynchronize(st));
    // peer access between every pair (the library moves panels device to device)
    for (int a : devs)
        for (int b : devs) {
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, a, b);
            if (!can) return fail(-116, "eigh: devices " + std::to_string(a) + " and " + std::to_string(b) + " have no peer access");
            cudaSetDevice(a);
            cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(-100, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            (void)cudaGetLastError();
        }
    const int nb = (int)devs.size();
    const size_t T = 256;                                   // column panel width of the block-cyclic layout
    const size_t npanels = (n + T - 1) / T;
    std::vector<size_t> local_cols(nb, 0);
    for (size_t b = 0; b < npanels; ++b) local_cols[b % nb] += std::min(T, n - b * T);
    std::vector<void*> d_a(nb, nullptr), d_work(nb, nullptr);
    cusolverMgHandle_t h = nullptr;
    cudaLibMgGrid_t grid = nullptr;
    cudaLibMgMatrixDesc_t desc = nullptr;
    std::vector<double> w_host(n);
    auto done = [&](int code, const std::string& msg) {
        for (int i = 0; i < nb; ++i) {
            cudaSetDevice(devs[i]);
            if (d_a[i]) cudaFree(d_a[i]);
            if (d_work[i]) cudaFree(d_work[i]);
        }
        if (desc) s.destroy_desc(desc);
        if (grid) s.destroy_grid(grid);
        if (h) s.destroy(h);
        cudaSetDevice(device);
        (void)cudaGetLastError();
        return code ? fail(code, msg) : 0;
    };
    if (s.create(&h) != CUSOLVER_STATUS_SUCCESS) return done(-111, "cusolverMgCreate failed");
    if (s.device_select(h, nb, devs.data()) != CUSOLVER_STATUS_SUCCESS) return done(-111, "cusolverMgDeviceSelect failed");
    if (s.create_grid(&grid, 1, nb, devs.data(), CUDALIBMG_GRID_MAPPING_COL_MAJOR) != CUSOLVER_STATUS_SUCCESS)
        return done(-111, "cusolverMgCreateDeviceGrid failed");
    if (s.create_desc(&desc, (int64_t)n, (int64_t)n, (int64_t)n, (int64_t)T, CUDA_R_64F, grid) != CUSOLVER_STATUS_SUCCESS)
        return done(-111, "cusolverMgCreateMatrixDesc failed");
    for (int i = 0; i < nb; ++i) {
        cudaSetDevice(devs[i]);
        if (cudaMalloc(&d_a[i], std::max<size_t>(local_cols[i], 1) * n * sizeof(double)) != cudaSuccess)
            return done(-100, "eigh: panel allocation failed on device " + std::to_string(devs[i]));
    }
    // deal the panels: the matrix is symmetric, so panel b (n x T, column-major, lda = n) is rows [bT, bT+T) of the
    // row-major buffer -- one contiguous block
    cudaSetDevice(device);
    for (size_t b = 0; b < npanels; ++b) {
        const size_t w = std::min(T, n - b * T);
        const int i = (int)(b % nb);
        double* dst = (double*)d_a[i] + (b / nb) * T * n;
        if (cudaMemcpyPeerAsync(dst, devs[i], a_dev + b * T * n, device, w * n * sizeof(double), st) != cudaSuccess)
            return done(-100, "eigh: panel scatter failed");
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) return done(-100, "eigh: panel scatter failed");
    int64_t lwork = 0;
    cusolverStatus_t sb = s.buffer_size(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, d_a.data(), 1, 1, desc,
                                        w_host.data(), CUDA_R_64F, CUDA_R_64F, &lwork);
    if (sb != CUSOLVER_STATUS_SUCCESS)
        return done(-112, "cusolverMgSyevd_bufferSize failed with status " + std::to_string((int)sb) + " (n=" + std::to_string(n) + ")");
    for (int i = 0; i < nb; ++i) {
        cudaSetDevice(devs[i]);
        if (cudaMalloc(&d_work[i], (size_t)std::max<int64_t>(lwork, 1) * sizeof(double)) != cudaSuccess)
            return done(-100, "eigh: workspace allocation of " + std::to_string(lwork * 8) + " bytes failed on device " +
                                  std::to_string(devs[i]));
    }
    cudaSetDevice(device);
    int info = 0;
    cusolverStatus_t stt = s.syevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, d_a.data(), 1, 1, desc,
                                   w_host.data(), CUDA_R_64F, CUDA_R_64F, d_work.data(), lwork, &info);
    for (int i = 0; i < nb; ++i) { cudaSetDevice(devs[i]); cudaDeviceSynchronize(); }
    cudaSetDevice(device);
    if (stt != CUSOLVER_STATUS_SUCCESS) return done(-113, "cusolverMgSyevd failed with status " + std::to_string((int)stt));
    if (info != 0) return done(-114, "eigendecomposition did not converge (cusolverMg info=" + std::to_string(info) + ")");
    // collect: eigenvector k is column k of the distributed result; columns are contiguous, so the assembled column-major
    // matrix read row-major is U^T
    for (size_t b = 0; b < npanels; ++b) {
        const size_t w = std::min(T, n - b * T);
        const int i = (int)(b % nb);
        const double* src = (const double*)d_a[i] + (b / nb) * T * n;
        if (cudaMemcpyPeerAsync(a_dev + b * T * n, device, src, devs[i], w * n * sizeof(double), st) != cudaSuccess)
            return done(-100, "eigh: panel gather failed");
    }
    if (cudaMemcpyAsync(evals_dev, w_host.data(), n * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        return done(-100, "eigh: result gather failed");
    if (ut_f32_dev) {
        narrow_kernel<<<148 * 8, 256, 0, st>>>(a_dev, ut_f32_dev, n * n);
        note_launch(1);
        if (cudaGetLastError() != cudaSuccess) return done(-100, "narrow kernel launch failed");
    }
    return done(0, "");
}

int eigh_device(int device, size_t n, double* a_dev, double diag_shift, double* evals_dev, float* ut_f32_dev,
                cudaStream_t st) {
    const char* force_mg = getenv("JXB_EIGH_FORCE_MG");     // tests: exercise the cusolverMg path at small n / on one device
    if (n > 46340 || g_eigh_devices > 1 || (force_mg && force_mg[0] == '1')) return eigh_device_mg(device, std::max(1, g_eigh_devices), n, a_dev, diag_shift, evals_dev, ut_f32_dev, st);
    Solver& s = solver();
    if (!s.ok) return fail(-110, "libcusolver (cusolverDnXsyevd) could not be loaded: the eigendecomposition has no CPU fallback");
    JXB_CUDA_OK(cudaSetDevice(device));
    if (diag_shift != 0.0) {
        add_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a_dev, n, diag_shift);
        note_launch(1);
        JXB_CUDA_OK(cudaGetLastError());
    }
    cusolverDnHandle_t h = nullptr;
    cusolverDnParams_t prm = nullptr;
    void* ws_dev = nullptr;
    void* ws_host = nullptr;
    int* info_dev = nullptr;
    int rc = 0;
    auto done = [&](int code, const std::string& msg) {
        if (ws_dev) cudaFree(ws_dev);
        if (info_dev) cudaFree(info_dev);
        free(ws_host);
        if (prm) s.destroy_params(prm);
        if (h) s.destroy(h);
        return code ? fail(code, msg) : 0;
    };
    if (s.create(&h) != CUSOLVER_STATUS_SUCCESS) return done(-111, "cusolverDnCreate failed");
    if (s.set_stream(h, st) != CUSOLVER_STATUS_SUCCESS) return done(-111, "cusolverDnSetStream failed");
    // bit-reproducible decomposition: the TSV of a job must not depend on the run (or on the number of GPUs)
    if (s.set_det) (void)s.set_det(h, CUSOLVER_DETERMINISTIC_RESULTS);
    if (s.create_params(&prm) != CUSOLVER_STATUS_SUCCESS) return done(-111, "cusolverDnCreateParams failed");
    size_t bytes_dev = 0, bytes_host = 0;
    // symmetric input: "lower, column-major" of the row-major buffer is its upper triangle -- either is the matrix
    const cusolverStatus_t sb = s.buffer_size(h, prm, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int64_t)n,
                                              CUDA_R_64F, a_dev, (int64_t)n, CUDA_R_64F, evals_dev, CUDA_R_64F,
                                              &bytes_dev, &bytes_host);
    if (sb != CUSOLVER_STATUS_SUCCESS)
        return done(-112, "cusolverDnXsyevd_bufferSize failed with status " + std::to_string((int)sb) + " (n=" +
                              std::to_string(n) + ")");
    if (cudaMalloc(&ws_dev, bytes_dev ? bytes_dev : 16) != cudaSuccess || cudaMalloc((void**)&info_dev, sizeof(int)) != cudaSuccess) {
        cudaGetLastError();
        return done(-100, "eigh workspace allocation of " + std::to_string(bytes_dev) + " bytes failed");
    }
    if (bytes_host) {
        ws_host = malloc(bytes_host);
        if (!ws_host) return done(-100, "eigh host workspace allocation failed");
    }
    const cusolverStatus_t stt = s.syevd(h, prm, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int64_t)n, CUDA_R_64F,
                                         a_dev, (int64_t)n, CUDA_R_64F, evals_dev, CUDA_R_64F, ws_dev, bytes_dev, ws_host,
                                         bytes_host, info_dev);
    if (stt != CUSOLVER_STATUS_SUCCESS) return done(-113, "cusolverDnXsyevd failed with status " + std::to_string((int)stt));
    int info = 0;
    if (cudaMemcpyAsync(&info, info_dev, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
        const std::string e = cudaGetErrorString(cudaGetLastError());
        return done(-100, "eigh: " + e);
    }
    if (info != 0) return done(-114, "eigendecomposition did not converge (cusolver info=" + std::to_string(info) + ")");
    if (ut_f32_dev) {
        narrow_kernel<<<148 * 8, 256, 0, st>>>(a_dev, ut_f32_dev, n * n);
        note_launch(1);
        if (cudaGetLastError() != cudaSuccess) return done(-100, "narrow kernel launch failed");
    }
    rc = done(0, "");
    return rc;
}

}  // namespace
}  // namespace jxb

using jxb::fail;

extern "C" void jxb_set_eigh_devices(int n_devices) { jxb::g_eigh_devices = n_devices < 0 ? 0 : n_devices; }

extern "C" int jxb_eigh_dev(int device, size_t n, double* a_dev, double diag_shift, double* evals_dev,
                            float* ut_f32_dev, void* stream) {
    if (!a_dev || !evals_dev || n == 0) return fail(-2, "null argument");
    return jxb::eigh_device(device, n, a_dev, diag_shift, evals_dev, ut_f32_dev, (cudaStream_t)stream);
}

extern "C" int jxb_eigh(int device, size_t n, const double* a_host, double diag_shift, double* evals_host,
                        double* ut_host, float* ut_f32_host) {
    if (!a_host || !evals_host || n == 0) return fail(-2, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(-1, "no CUDA device is visible: janusx_b200 has no CPU fallback");
    }
    JXB_CUDA_OK(cudaSetDevice(device));
    double* a = nullptr;
    double* w = nullptr;
    float* u32 = nullptr;
    auto cleanup = [&]() { if (a) cudaFree(a); if (w) cudaFree(w); if (u32) cudaFree(u32); };
    cudaError_t e = cudaMalloc((void**)&a, n * n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&w, n * sizeof(double));
    if (e == cudaSuccess && ut_f32_host) e = cudaMalloc((void**)&u32, n * n * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(a, a_host, n * n * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cleanup(); return fail(-100, std::string("eigh staging: ") + cudaGetErrorString(e)); }
    int rc = jxb::eigh_device(device, n, a, diag_shift, w, u32, nullptr);
    if (!rc) {
        e = cudaMemcpy(evals_host, w, n * sizeof(double), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && ut_host) e = cudaMemcpy(ut_host, a, n * n * sizeof(double), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && ut_f32_host) e = cudaMemcpy(ut_f32_host, u32, n * n * sizeof(float), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(-100, std::string("eigh readback: ") + cudaGetErrorString(e));
    }
    cleanup();
    return rc;
}
