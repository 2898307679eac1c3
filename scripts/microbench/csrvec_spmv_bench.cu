// csrvec_spmv_bench.cu -- would a shared-memory-free "CSR-vector" SpMV beat the tiled CSR-stream kernel on a
// matrix with uniformly random columns (configs[4])?  Every KB of shared memory is a KB less L1, and on B200 the
// number of random gathers in flight is bounded by the L1 capacity (profiles/r1_gather_ceiling_smem_sweep.txt).
// Here LPR lanes own a row: they stream its (col, val) pairs with L1::no_allocate loads, keep UNROLL independent
// x-gathers in flight each, and combine with shuffles -- no shared memory at all.
// Matrices: H-like (rows x 36 nnz, columns uniform in [0, 3 rows)) and A-like (2 rows x 5 nnz, columns in [0, rows)).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a csrvec_spmv_bench.cu -o csrvec_spmv_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ int ld_stream_i32(const int *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream_f64(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

template <int LPR, int UNROLL, int MINB>
__global__ void __launch_bounds__(256, MINB) spmv_csrvec(int rows, const int *__restrict__ rowptr, const int *__restrict__ col,
                                                         const double *__restrict__ val, const double *x, double *y) {
    const int groups_per_cta = 256 / LPR;
    const int g = threadIdx.x / LPR, gl = threadIdx.x % LPR;
    // contiguous block of rows per CTA (static ownership, like the persistent kernel), groups interleaved inside it
    const int per_cta = (rows + gridDim.x - 1) / gridDim.x;
    const int r0 = blockIdx.x * per_cta, r1 = min(rows, r0 + per_cta);
    for (int row = r0 + g; row < r1; row += groups_per_cta) {
        const int a = __ldg(rowptr + row), b = __ldg(rowptr + row + 1);
        double acc = 0.0;
        for (int k = a + gl; k < b; k += LPR * UNROLL) {
            int c[UNROLL];
            double v[UNROLL], xv[UNROLL];
#pragma unroll
            for (int j = 0; j < UNROLL; ++j) {
                const int kk = k + j * LPR;
                c[j] = kk < b ? ld_stream_i32(col + kk) : 0;
                v[j] = kk < b ? ld_stream_f64(val + kk) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < UNROLL; ++j) xv[j] = (k + j * LPR < b) ? x[c[j]] : 0.0;
#pragma unroll
            for (int j = 0; j < UNROLL; ++j) acc += v[j] * xv[j];
        }
#pragma unroll
        for (int o = LPR >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, LPR);
        if (gl == 0) y[row] = acc;
    }
}

__global__ void flush_kernel(const double *buf, size_t n, double *sink) {
    double s = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += buf[i];
    if (s == 123.456) *sink = s;
}

struct Mat {
    int rows, cols, per_row;
    int *rowptr, *col;
    double *val, *x, *y;
    std::vector<int> h_rowptr, h_col;
    std::vector<double> h_val;
};

static Mat make(int rows, int cols, int per_row) {
    Mat M{rows, cols, per_row};
    const size_t nnz = (size_t)rows * per_row;
    M.h_rowptr.resize(rows + 1);
    M.h_col.resize(nnz);
    M.h_val.resize(nnz);
    unsigned long long s = 88172645463325252ULL;
    for (int r = 0; r <= rows; ++r) M.h_rowptr[r] = r * per_row;
    for (size_t i = 0; i < nnz; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        M.h_col[i] = (int)(s % (unsigned long long)cols);
        M.h_val[i] = 1.0 + (double)(i % 7);
    }
    cudaMalloc(&M.rowptr, (rows + 1) * sizeof(int)); cudaMalloc(&M.col, nnz * sizeof(int)); cudaMalloc(&M.val, nnz * sizeof(double));
    cudaMalloc(&M.x, (size_t)cols * sizeof(double)); cudaMalloc(&M.y, (size_t)rows * sizeof(double));
    cudaMemcpy(M.rowptr, M.h_rowptr.data(), (rows + 1) * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(M.col, M.h_col.data(), nnz * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(M.val, M.h_val.data(), nnz * sizeof(double), cudaMemcpyHostToDevice);
    std::vector<double> hx(cols);
    for (int i = 0; i < cols; ++i) hx[i] = 1.0 + 1e-3 * (i % 1000);
    cudaMemcpy(M.x, hx.data(), (size_t)cols * sizeof(double), cudaMemcpyHostToDevice);
    return M;
}

static double *g_flush; static size_t g_flush_n = (size_t)48 << 20;

template <int LPR, int UNROLL, int MINB>
static void run(const Mat &M, int sms, const char *name) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmv_csrvec<LPR, UNROLL, MINB>, 256, 0);
    const int grid = sms * per_sm;
    double total = 0.0; const int reps = 10;
    for (int r = -2; r < reps; ++r) {
        flush_kernel<<<2048, 256>>>(g_flush, g_flush_n, M.y);
        cudaEventRecord(a);
        spmv_csrvec<LPR, UNROLL, MINB><<<grid, 256>>>(M.rows, M.rowptr, M.col, M.val, M.x, M.y);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (r >= 0) total += ms;
    }
    // check one row
    std::vector<double> hy(M.rows);
    cudaMemcpy(hy.data(), M.y, (size_t)M.rows * sizeof(double), cudaMemcpyDeviceToHost);
    double maxerr = 0.0;
    for (int row = 0; row < M.rows; row += M.rows / 97 + 1) {
        double s = 0.0;
        for (int k = M.h_rowptr[row]; k < M.h_rowptr[row + 1]; ++k) s += M.h_val[k] * (1.0 + 1e-3 * (M.h_col[k] % 1000));
        maxerr = fmax(maxerr, fabs(s - hy[row]) / (1.0 + fabs(s)));
    }
    const double ms = total / reps;
    const double nnz = (double)M.rows * M.per_row;
    const double bytes = 12.0 * nnz + 4.0 * (M.rows + 1) + 8.0 * M.cols + 8.0 * M.rows;
    printf("{\"matrix\": \"%s\", \"lpr\": %d, \"unroll\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"GBs\": %.1f, \"Ggather_s\": %.1f, \"max_rel_err\": %.2e}\n",
           name, LPR, UNROLL, per_sm, ms, bytes / ms / 1e6, nnz / ms / 1e6, maxerr);
    fflush(stdout);
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 1000000;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    cudaMalloc(&g_flush, g_flush_n * sizeof(double)); cudaMemset(g_flush, 0, g_flush_n * sizeof(double));
    cudaFuncSetCacheConfig(flush_kernel, cudaFuncCachePreferL1);
    {
        Mat H = make(n, 3 * n, 36);
        run<4, 5, 4>(H, sms, "H");  run<4, 9, 4>(H, sms, "H");  run<8, 5, 4>(H, sms, "H");  run<8, 5, 6>(H, sms, "H");
        run<8, 5, 8>(H, sms, "H");  run<4, 5, 8>(H, sms, "H");  run<4, 9, 6>(H, sms, "H");  run<16, 3, 8>(H, sms, "H");
        run<2, 9, 4>(H, sms, "H");  run<2, 9, 8>(H, sms, "H");
        cudaFree(H.rowptr); cudaFree(H.col); cudaFree(H.val); cudaFree(H.x); cudaFree(H.y);
    }
    {
        Mat A = make(2 * n, n, 5);
        run<1, 5, 4>(A, sms, "A");  run<1, 5, 8>(A, sms, "A");  run<2, 3, 8>(A, sms, "A");  run<4, 2, 8>(A, sms, "A");
    }
    return 0;
}
