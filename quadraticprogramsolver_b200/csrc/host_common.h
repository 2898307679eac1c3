// host_common.h -- host-side helpers shared by the translation units of libqpb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>
#include <vector>

#include "qpb200.h"

namespace qpb {

// thread-local error text returned by qpb200_last_error()
std::string &last_error();
int fail(int code, const char *fmt, ...);

#define QPB_CUDA(call)                                                                               \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return ::qpb::fail(QPB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                               __FILE__, __LINE__);                                                  \
    } while (0)

// Owns device allocations of one handle.
struct DeviceArena {
    std::vector<void *> ptrs;
    size_t bytes = 0;
    template <class T>
    cudaError_t alloc(T **out, size_t count, bool zero = false) {
        void *p = nullptr;
        const size_t nb = (count > 0 ? count : 1) * sizeof(T);
        cudaError_t e = cudaMalloc(&p, nb);
        if (e != cudaSuccess) return e;
        ptrs.push_back(p);
        bytes += nb;
        if (zero) {
            e = cudaMemset(p, 0, nb);
            if (e != cudaSuccess) return e;
        }
        *out = static_cast<T *>(p);
        return cudaSuccess;
    }
    void release() {
        for (void *p : ptrs) cudaFree(p);
        ptrs.clear();
        bytes = 0;
    }
};

// Uninitialised POD buffer: std::vector::resize would zero-fill (and page-fault) hundreds of MB on one
// thread; here the first touch happens in the parallel fill loops.
template <class T>
struct PodBuf {
    T *p = nullptr;
    size_t n = 0;
    PodBuf() = default;
    PodBuf(const PodBuf &) = delete;
    PodBuf &operator=(const PodBuf &) = delete;
    ~PodBuf() { free(p); }
    void resize(size_t k) {
        free(p);
        p = static_cast<T *>(malloc((k > 0 ? k : 1) * sizeof(T)));
        n = k;
    }
    T *data() { return p; }
    const T *data() const { return p; }
    size_t size() const { return n; }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
};

// Host CSR with int32 indices (the device layout before tiling).
struct HostCsr {
    int rows = 0, cols = 0;
    std::vector<int> ptr;      // rows + 1
    PodBuf<int> idx;           // nnz
    PodBuf<double> val;        // nnz
    std::vector<int> mid;      // rows (optional)
    int64_t nnz() const { return ptr.empty() ? 0 : ptr.back(); }
};

struct HostTiles {
    std::vector<int4> tiles;
    std::vector<int> cta_begin;   // grid + 1
    int lpr = 1;
};

// Julia SparseMatrixCSC (cols = ncols) -> CSR of the same matrix.
void csc_to_csr(int64_t nrows, int64_t ncols, const int64_t *colptr, const int64_t *rowval, const double *nzval,
                int64_t base, HostCsr &out);
// Julia SparseMatrixCSC -> CSR of the transpose (the same arrays re-indexed).
void csc_as_csr_of_transpose(int64_t nrows, int64_t ncols, const int64_t *colptr, const int64_t *rowval,
                             const double *nzval, int64_t base, HostCsr &out);
// Validate a CSC triplet: monotone colptr, indices in range.  Returns 0 or an error code (message set).
int validate_csc(const char *name, int64_t nrows, int64_t ncols, const int64_t *colptr, const int64_t *rowval,
                 const double *nzval, int64_t base);
void build_tiles(const HostCsr &M, int tile_nnz, HostTiles &out);
void assign_tiles(HostTiles &t, int grid);
int choose_lpr(const HostCsr &M);

int host_threads();
void parallel_chunks(int64_t count, const std::function<void(int, int64_t, int64_t)> &fn, int64_t min_chunk);

// true iff every entry is finite (vectorisable: v * 0 is NaN exactly for NaN / +-Inf)
bool all_finite(const double *v, size_t count);

int check_device(int device);   // 0 or QPB200_ERR_DEVICE / QPB200_ERR_CUDA

}  // namespace qpb
