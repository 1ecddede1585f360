// cb2_peak.cu — FP32 FFMA peak micro-benchmark (a roofline denominator for the narrow phase;
// MEASURED_PEAKS.json only carries HBM and bf16 tensor peaks).  Not part of the C ABI product
// library: built as libcb2_peak.so and used by bench.py only.
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void __launch_bounds__(256) k_ffma_peak(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
          x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

extern "C" double cb2_peak_fp32_tflops(int device) {
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, device);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
    float* out = nullptr;
    if (cudaMalloc(&out, sizeof(float) * blocks * threads) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        k_ffma_peak<<<blocks, threads>>>(out, iters, 1.0000001f, 1e-9f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return best;
}
