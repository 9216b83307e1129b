// eigh.cu -- SURVEY 8(f) N2 / row A17: spectral decomposition of the GRM on the device.
//
// Replaces rust_eigh_from_array_f64[_inplace] (src/math/eigh.rs:1621-1705, 1883-1990; LAPACK dsyevd/dsyevr drivers
// :1201-1420) with ONE cuSOLVER library call (cusolverDnXsyevd, 64-bit API), as SURVEY A17 prescribes ("cuSOLVER
// library call, not a hand kernel").  The library is opened with dlopen so libjxb200.so has no link-time dependency
// on it.  Output convention: eigenvalues ascending; the matrix is overwritten by U^T row-major (row k = k-th
// eigenvector) -- cuSOLVER's column-major eigenvector matrix read as row-major IS U^T, the `Dh = U.T` the scan takes.
#include <cusolverDn.h>
#include <dlfcn.h>

#include <cstdlib>
#include <string>

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

int eigh_device(int device, size_t n, double* a_dev, double diag_shift, double* evals_dev, float* ut_f32_dev,
                cudaStream_t st) {
    if (n > 46340)
        return fail(-115, "eigendecomposition of n=" + std::to_string(n) + " is beyond cusolverDnXsyevd (it rejects n*n >= 2^31, "
                          "i.e. n > 46340); decompose on the host or per population block and pass (S, U^T) to jxb_model_create");
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
