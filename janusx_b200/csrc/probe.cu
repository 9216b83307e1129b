// probe.cu -- in-run measurement of the FP64 CUDA-core issue rate that bounds the per-SNP solve (bench.py's
// roofline denominator; MEASURED_PEAKS.json holds no FP64 figure).  Register-resident chains, no memory traffic.
#include "../../include/jxb200.h"
#include "jxb_common.cuh"

namespace jxb {
namespace {

template <int OP>
__global__ void __launch_bounds__(256) fp64_rate_kernel(double* out, int iters, double a, double b) {
    constexpr int ILP = 8;
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = a + k + threadIdx.x * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            if (OP == 0) x[k] = fma(x[k], b, a);
            else if (OP == 1) x[k] = __dadd_rn(x[k], b);
            else x[k] = __dmul_rn(x[k], b);
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    if (s == 12345.678) out[0] = s;
}

template <int OP>
int rate(int sms, double* d, double* tops) {
    const int blocks = sms * 8, threads = 256, iters = 4096;
    cudaEvent_t e0, e1;
    JXB_CUDA_OK(cudaEventCreate(&e0));
    JXB_CUDA_OK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        fp64_rate_kernel<OP><<<blocks, threads>>>(d, iters, 1.0000001, 0.9999999);
        cudaEventRecord(e1);
        JXB_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double t = (double)blocks * threads * iters * 8 / (ms * 1e-3) / 1e12;
        if (rep && t > best) best = t;       // first launch = warm-up
    }
    note_launch(4);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tops = best;
    return 0;
}

}  // namespace
}  // namespace jxb

extern "C" int jxb_fp64_probe(int device, double tops3[3]) {
    using namespace jxb;
    if (!tops3) return fail(-2, "null argument");
    JXB_CUDA_OK(cudaSetDevice(device));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    double* d = nullptr;
    JXB_CUDA_OK(cudaMalloc((void**)&d, 8));
    int rc = rate<0>(sms, d, &tops3[0]);
    if (!rc) rc = rate<1>(sms, d, &tops3[1]);
    if (!rc) rc = rate<2>(sms, d, &tops3[2]);
    cudaFree(d);
    return rc;
}
