// fp64_peak_bench.cu -- measured FP64 peaks of the B200 the dense-batch roofline is quoted against:
//   (1) DFMA issue rate of the FP64 pipe (register-resident dependent chains, ILP independent chains per thread)
//   (2) DMMA.8x8x4 rate of the FP64 tensor pipe (mma.sync.aligned.m8n8k4.f64)
// Prints one JSON line: {"dfma_tflops": ..., "dmma_tflops": ..., "sm_count": ..., "clock_mhz": ...}
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_peak_bench.cu -o fp64_peak_bench
#include <cuda_runtime.h>
#include <cstdio>

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc[j] = threadIdx.x * 1e-3 + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) acc[j] = fma(acc[j], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += acc[j];
    if (s == 1.2345e300) out[0] = s;
}

__global__ void __launch_bounds__(256) dmma_kernel(double *out, int iters) {
    double c0[4] = {0.0, 0.0, 0.0, 0.0}, c1[4] = {0.0, 0.0, 0.0, 0.0};
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[j]), "+d"(c1[j]) : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
    for (int j = 0; j < 4; ++j) s += c0[j] + c1[j];
    if (s == 1.2345e300) out[0] = s;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    double *out;
    cudaMalloc(&out, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int sms = prop.multiProcessorCount;
    double best_dfma = 0.0, best_dmma = 0.0;
    for (int ctas = 2; ctas <= 8; ctas *= 2) {
        const int iters = 20000;
        const int grid = sms * ctas;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            dfma_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double tf = 2.0 * 8 * (double)iters * 256.0 * grid / (ms * 1e-3) / 1e12;
            if (rep > 0 && tf > best_dfma) best_dfma = tf;
        }
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            dmma_kernel<<<grid, 256>>>(out, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            // one m8n8k4 = 8*8*4 FMAs = 512 flops per warp instruction
            const double tf = 512.0 * 4 * (double)iters * 8.0 * grid / (ms * 1e-3) / 1e12;
            if (rep > 0 && tf > best_dmma) best_dmma = tf;
        }
    }
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"sm_count\": %d, \"clock_mhz\": %.0f, \"device\": \"%s\"}\n", best_dfma, best_dmma,
           sms, clk / 1e3, prop.name);
    return 0;
}
