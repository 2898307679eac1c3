// texgather_bench.cu -- does the texture path give uniformly random 8-byte gathers a second data pipe?
// The SpMV tile engine is bound by the L1TEX LSU data pipe: one wavefront per scattered gather plus the shared-memory
// wavefronts of the tile (ncu: l1tex__data_pipe_lsu_wavefronts 85 % on a banded matrix, the cfg5 H pass time equals
// gathers + shared wavefronts at one per clock).  Here: the same gather loop with ld.global and with
// tex1Dfetch<int2> on a linear texture object, alone and next to a shared-memory load stream.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a texgather_bench.cu -o texgather_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

template <int ILP, int MODE, int SMEM_LOADS>
__global__ void __launch_bounds__(256, 4) gather_kernel(const int *__restrict__ idx, const double *x, cudaTextureObject_t tex,
                                                        double *out, size_t n) {
    __shared__ double sbuf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sbuf[i] = i;
    __syncthreads();
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t base = (size_t)blockIdx.x * blockDim.x + threadIdx.x; base + (ILP - 1) * stride < n; base += ILP * stride) {
        int c[ILP];
        double v[ILP];
#pragma unroll
        for (int j = 0; j < ILP; ++j) c[j] = __ldg(idx + base + j * stride);
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            if (MODE == 0) v[j] = x[c[j]];
            else {
                const int2 t = tex1Dfetch<int2>(tex, c[j]);
                v[j] = __hiloint2double(t.y, t.x);
            }
        }
#pragma unroll
        for (int j = 0; j < ILP; ++j) acc += v[j];
#pragma unroll
        for (int j = 0; j < SMEM_LOADS; ++j) acc += sbuf[(threadIdx.x + 256 * j + (c[0] & 1)) & 2047];   // conflict-free shared loads
    }
    if (acc == 1.2345e300) out[0] = acc;
}

template <int ILP, int MODE, int SMEM_LOADS>
float run(const int *idx, const double *x, cudaTextureObject_t tex, double *out, size_t n, int blocks, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    gather_kernel<ILP, MODE, SMEM_LOADS><<<blocks, 256>>>(idx, x, tex, out, n);
    cudaEventRecord(a);
    for (int r = 0; r < reps; ++r) gather_kernel<ILP, MODE, SMEM_LOADS><<<blocks, 256>>>(idx, x, tex, out, n);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char **argv) {
    const size_t nvec = argc > 1 ? (size_t)atoll(argv[1]) : 3000000;
    const size_t n = argc > 2 ? (size_t)atoll(argv[2]) : 36000000;
    std::vector<int> h(n);
    unsigned long long s = 88172645463325252ULL;
    for (size_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % nvec); }
    int *idx; double *x, *out;
    cudaMalloc(&idx, n * sizeof(int)); cudaMalloc(&x, nvec * sizeof(double)); cudaMalloc(&out, 8);
    cudaMemcpy(idx, h.data(), n * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemset(x, 0, nvec * sizeof(double));
    cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = x;
    rd.res.linear.desc = cudaCreateChannelDesc<int2>(); rd.res.linear.sizeInBytes = nvec * sizeof(double);
    cudaTextureDesc td{}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    cudaError_t e = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    if (e != cudaSuccess) { printf("texture object: %s\n", cudaGetErrorString(e)); return 1; }
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 4;
    printf("device %s; x = %.1f MB, %zu gathers per launch, 4 CTAs x 256 threads per SM\n", p.name, nvec * 8 / 1e6, n);
    printf("%-44s %10s %12s %14s\n", "variant", "ms", "Ggather/s", "gathers/clk/SM");
#define RUN(ILP, MODE, SL, NAME)                                                                            \
    {                                                                                                       \
        float ms = run<ILP, MODE, SL>(idx, x, tex, out, n, blocks, 10);                                     \
        printf("%-44s %10.4f %12.1f %14.3f\n", NAME, ms, n / ms / 1e6, n / (ms * 1e-3) / (sms * (double)p.clockRate * 1e3)); \
    }
    RUN(4, 0, 0, "ld.global ILP=4")
    RUN(8, 0, 0, "ld.global ILP=8")
    RUN(4, 1, 0, "tex1Dfetch ILP=4")
    RUN(8, 1, 0, "tex1Dfetch ILP=8")
    RUN(4, 0, 4, "ld.global ILP=4 + 4 shared loads")
    RUN(4, 1, 4, "tex1Dfetch ILP=4 + 4 shared loads")
    RUN(4, 0, 8, "ld.global ILP=4 + 8 shared loads")
    RUN(4, 1, 8, "tex1Dfetch ILP=4 + 8 shared loads")
    RUN(8, 0, 8, "ld.global ILP=8 + 8 shared loads")
    RUN(8, 1, 8, "tex1Dfetch ILP=8 + 8 shared loads")
    return 0;
}
