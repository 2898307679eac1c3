// gather_bench.cu -- what is the B200's ceiling for uniformly random 8-byte gathers out of an L2-resident
// vector?  (The x-gathers of a CSR SpMV on a matrix with random columns, e.g. configs[4].)
// Every thread streams coalesced int32 indices and gathers x[idx] with ILP independent loads in flight.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a gather_bench.cu -o gather_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

template <int ILP, int MODE>
__global__ void gather_kernel(const int *__restrict__ idx, const double *x, double *out, size_t n) {
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t base = (size_t)blockIdx.x * blockDim.x + threadIdx.x; base + (ILP - 1) * stride < n; base += ILP * stride) {
        int c[ILP];
        double v[ILP];
#pragma unroll
        for (int j = 0; j < ILP; ++j) c[j] = __ldg(idx + base + j * stride);
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            if (MODE == 0) v[j] = x[c[j]];                     // ld.global (L1 allocate)
            else if (MODE == 1) v[j] = __ldcg(x + c[j]);       // ld.global.cg (L2 only)
            else if (MODE == 2) v[j] = __ldg(x + c[j]);        // ld.global.nc
            else asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v[j]) : "l"(x + c[j]));
        }
#pragma unroll
        for (int j = 0; j < ILP; ++j) acc += v[j];
    }
    if (acc == 1.2345e300) out[0] = acc;
}

template <int ILP, int MODE>
float run(const int *idx, const double *x, double *out, size_t n, int blocks, int threads, int reps, size_t smem = 0) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    // smem > 0: an (unused) dynamic shared-memory allocation per CTA -- it shrinks the L1 data cache
    cudaFuncSetAttribute(gather_kernel<ILP, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gather_kernel<ILP, MODE><<<blocks, threads, smem>>>(idx, x, out, n);
    cudaEventRecord(a);
    for (int r = 0; r < reps; ++r) gather_kernel<ILP, MODE><<<blocks, threads, smem>>>(idx, x, out, n);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char **argv) {
    const size_t nvec = argc > 1 ? (size_t)atoll(argv[1]) : 3000000;      // doubles in x (24 MB: L2 resident)
    const size_t n = argc > 2 ? (size_t)atoll(argv[2]) : 36000000;        // gathers per launch
    std::vector<int> h(n);
    unsigned long long s = 88172645463325252ULL;
    for (size_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % nvec); }
    int *idx; double *x, *out;
    cudaMalloc(&idx, n * sizeof(int)); cudaMalloc(&x, nvec * sizeof(double)); cudaMalloc(&out, 8);
    cudaMemcpy(idx, h.data(), n * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemset(x, 0, nvec * sizeof(double));
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("device %s, %d SMs; x = %.1f MB, %zu gathers per launch (index stream %.0f MB)\n", p.name, sms, nvec * 8 / 1e6, n, n * 4 / 1e6);
    printf("%-28s %8s %10s %12s %14s\n", "variant", "thr/SM", "ms", "Ggather/s", "gathers/clk/SM");
    const char *modes[] = {"ld.global", "ld.global.cg", "ld.global.nc", "L1::no_allocate"};
    for (int thr_per_sm : {512, 1024, 2048}) {
        const int threads = 512, blocks = sms * thr_per_sm / threads;
#define RUN(ILP, MODE)                                                                                      \
        {                                                                                                   \
            float ms = run<ILP, MODE>(idx, x, out, n, blocks, threads, 10);                                 \
            char name[64]; snprintf(name, sizeof(name), "%s ILP=%d", modes[MODE], ILP);                     \
            printf("%-28s %8d %10.4f %12.1f %14.3f\n", name, thr_per_sm, ms, n / ms / 1e6,                  \
                   n / (ms * 1e-3) / (sms * (double)p.clockRate * 1e3));                                   \
        }
        RUN(4, 0) RUN(8, 0) RUN(16, 0) RUN(8, 1) RUN(8, 2) RUN(8, 3) RUN(16, 1)
    }
    // L1 capacity vs gathers in flight: 2 CTAs x 512 threads per SM, ILP 4 / 8, with a shared-memory
    // allocation per CTA like the SpMV kernel's (every KB of shared memory is a KB less L1)
    printf("\n%-28s %8s %10s %12s %14s\n", "1024 thr/SM, smem/CTA", "KB", "ms", "Ggather/s", "gathers/clk/SM");
    for (int kb : {0, 16, 33, 50, 75, 100}) {
        const int threads = 512, blocks = sms * 2;
        float m4 = run<4, 0>(idx, x, out, n, blocks, threads, 10, (size_t)kb * 1024);
        float m8 = run<8, 0>(idx, x, out, n, blocks, threads, 10, (size_t)kb * 1024);
        printf("%-28s %8d %10.4f %12.1f %14.3f\n", "ld.global ILP=4", kb, m4, n / m4 / 1e6, n / (m4 * 1e-3) / (sms * (double)p.clockRate * 1e3));
        printf("%-28s %8d %10.4f %12.1f %14.3f\n", "ld.global ILP=8", kb, m8, n / m8 / 1e6, n / (m8 * 1e-3) / (sms * (double)p.clockRate * 1e3));
    }
    // the same sweep for loads that should not need an L1 line per miss in flight
    printf("\n%-28s %8s %10s %12s %14s\n", "1024 thr/SM, smem/CTA", "KB", "ms", "Ggather/s", "gathers/clk/SM");
    for (int kb : {0, 50, 100}) {
        const int threads = 512, blocks = sms * 2;
        float a = run<8, 1>(idx, x, out, n, blocks, threads, 10, (size_t)kb * 1024);
        float b = run<8, 3>(idx, x, out, n, blocks, threads, 10, (size_t)kb * 1024);
        float c = run<8, 2>(idx, x, out, n, blocks, threads, 10, (size_t)kb * 1024);
        printf("%-28s %8d %10.4f %12.1f %14.3f\n", "ld.global.cg ILP=8", kb, a, n / a / 1e6, n / (a * 1e-3) / (sms * (double)p.clockRate * 1e3));
        printf("%-28s %8d %10.4f %12.1f %14.3f\n", "L1::no_allocate ILP=8", kb, b, n / b / 1e6, n / (b * 1e-3) / (sms * (double)p.clockRate * 1e3));
        printf("%-28s %8d %10.4f %12.1f %14.3f\n", "ld.global.nc ILP=8", kb, c, n / c / 1e6, n / (c * 1e-3) / (sms * (double)p.clockRate * 1e3));
    }
    return 0;
}
