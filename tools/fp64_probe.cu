// Microbenchmarks that fix the FP64 CUDA-core roofline for the per-SNP solve kernel on this B200:
// DFMA / DADD / DMUL issue throughput per SM, dependent-chain latency, f64 divide and log cost.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OP>
__global__ void thr_kernel(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = a + k + threadIdx.x * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            if (OP == 0) x[k] = fma(x[k], b, a);
            else if (OP == 1) x[k] = x[k] + b;
            else if (OP == 2) x[k] = x[k] * b;
            else if (OP == 3) x[k] = 1.0 / (x[k] + a);
            else if (OP == 4) x[k] = log(x[k] + a) + 2.0;
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    if (s == 12345.678) out[0] = s;
}

template <int ILP, int OP>
double run(const char* name, int blocks, int threads, int iters) {
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    thr_kernel<ILP, OP><<<blocks, threads>>>(d, iters, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    thr_kernel<ILP, OP><<<blocks, threads>>>(d, iters, 1.0000001, 0.9999999);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double ops = (double)blocks * threads * iters * ILP;
    double per_clk_sm = ops / (ms * 1e-3) / (clk_khz * 1e3) / sms;
    printf("%-28s blocks=%d thr=%d ilp=%d: %.3f ms  %.2f lane-ops/clk/SM (nominal clk %d MHz)  %.2f Tops/s\n", name, blocks,
           threads, ILP, ms, per_clk_sm, clk_khz / 1000, ops / (ms * 1e-3) / 1e12);
    cudaFree(d);
    return per_clk_sm;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    // throughput: many warps, ILP 8
    run<8, 0>("DFMA throughput", sms * 8, 256, 4096);
    run<8, 1>("DADD throughput", sms * 8, 256, 4096);
    run<8, 2>("DMUL throughput", sms * 8, 256, 4096);
    run<4, 3>("f64 divide throughput", sms * 8, 256, 512);
    run<4, 4>("f64 log throughput", sms * 8, 256, 512);
    // latency: one warp per SM, ILP 1 -> clk per dependent op = 32 / (lane-ops/clk/SM)
    double r = run<1, 0>("DFMA dependent (1 warp/SM)", sms, 32, 1 << 16);
    printf("  => DFMA dependent latency ~ %.1f clk\n", 32.0 / r);
    r = run<1, 1>("DADD dependent (1 warp/SM)", sms, 32, 1 << 16);
    printf("  => DADD dependent latency ~ %.1f clk\n", 32.0 / r);
    r = run<1, 3>("divide dependent (1 warp/SM)", sms, 32, 1 << 13);
    printf("  => divide dependent latency ~ %.1f clk\n", 32.0 / r);
    r = run<1, 4>("log dependent (1 warp/SM)", sms, 32, 1 << 13);
    printf("  => log dependent latency ~ %.1f clk\n", 32.0 / r);
    // one warp per SMSP with ILP 8: can a single warp saturate its sub-partition?
    run<8, 0>("DFMA 4 warps/SM ilp8", sms, 128, 1 << 14);
    run<8, 0>("DFMA 16 warps/SM ilp8", sms, 512, 1 << 13);
    return 0;
}
