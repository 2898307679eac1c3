// host_common.h -- host-side helpers shared by the translation units of libqpb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>
#include <utility>
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
// Large blocks (device and host) are recycled through small process-wide caches: a caller that creates and
// destroys handles repeatedly (MPC-style re-solves, the end-to-end benchmark) otherwise pays cudaMalloc/cudaFree
// of ~1 GB (25-150 ms, measured) and first-touch page faults of ~0.5 GB of host staging per create.
// Bounded (QPB200_CACHE_MB, default 4096 MB per kind; 0 disables); blocks are reused only for requests of
// similar size.
cudaError_t cached_device_alloc(void **p, size_t bytes, size_t *capacity);
void cached_device_free(void *p, size_t capacity);
void *cached_host_alloc(size_t bytes, size_t *capacity);
void cached_host_free(void *p, size_t capacity);
// Page-locked flavour for the arrays that are uploaded (H2D from pinned memory runs at PCIe speed and can overlap the
// host conversion; pageable copies measured 11 GB/s).  *pinned tells the caller what it got: when cudaHostAlloc
// fails (or no device is present) the block is plain malloc memory.  At most QPB200_PINNED_CACHE_MB (default 1536)
// of page-locked blocks are kept between handles.
void *cached_pinned_alloc(size_t bytes, size_t *capacity, bool *pinned);
void cached_pinned_free(void *p, size_t capacity, bool pinned);

struct DeviceArena {
    std::vector<std::pair<void *, size_t>> ptrs;
    size_t bytes = 0;
    template <class T>
    cudaError_t alloc(T **out, size_t count, bool zero = false) {
        void *p = nullptr;
        const size_t nb = (count > 0 ? count : 1) * sizeof(T);
        size_t cap = 0;
        cudaError_t e = cached_device_alloc(&p, nb, &cap);
        if (e != cudaSuccess) return e;
        ptrs.push_back({p, cap});
        bytes += nb;
        if (zero) {
            e = cudaMemset(p, 0, nb);
            if (e != cudaSuccess) return e;
        }
        *out = static_cast<T *>(p);
        return cudaSuccess;
    }
    void release() {
        for (auto &pc : ptrs) cached_device_free(pc.first, pc.second);
        ptrs.clear();
        bytes = 0;
    }
};

// Uninitialised POD buffer: std::vector::resize would zero-fill (and page-fault) hundreds of MB on one
// thread; here the first touch happens in the parallel fill loops.
template <class T>
struct PodBuf {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    bool want_pinned = false;   // set before resize(): page-locked memory for buffers that get uploaded
    bool is_pinned = false;
    PodBuf() = default;
    PodBuf(const PodBuf &) = delete;
    PodBuf &operator=(const PodBuf &) = delete;
    ~PodBuf() { drop(); }
    void drop() {
        if (!p) return;
        if (is_pinned) cached_pinned_free(p, cap, true);
        else cached_host_free(p, cap);
        p = nullptr;
    }
    void resize(size_t k) {
        drop();
        const size_t bytes = (k > 0 ? k : 1) * sizeof(T);
        is_pinned = false;
        if (want_pinned) p = static_cast<T *>(cached_pinned_alloc(bytes, &cap, &is_pinned));
        else p = static_cast<T *>(cached_host_alloc(bytes, &cap));
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

// Julia SparseMatrixCSC (cols = ncols) -> CSR of the same matrix.  gap (nrows + 1 offsets, optional): leave
// gap[r+1] - gap[r] free slots behind row r and record their start in out.mid[r] (in-place assembly of [P A']).
void csc_to_csr(int64_t nrows, int64_t ncols, const int64_t *colptr, const int64_t *rowval, const double *nzval,
                int64_t base, HostCsr &out, const int64_t *gap = nullptr);
// H = [P A'] (n x (n+m), SURVEY 8(d) / DESIGN 3) straight from the two CSC inputs: row j = row j of P, then column j
// of A with its row indices shifted by n; H.mid[j] marks the split.  Also diag(P) and the column square sums of A.
void assemble_h_direct(int64_t n, int64_t m, const int64_t *Pp, const int64_t *Pi, const double *Pv, const int64_t *Ap,
                       const int64_t *Ai, const double *Av, int64_t base, HostCsr &H, std::vector<double> &dP,
                       std::vector<double> &dAA);
// Julia SparseMatrixCSC -> CSR of the transpose (the same arrays re-indexed).
void csc_as_csr_of_transpose(int64_t nrows, int64_t ncols, const int64_t *colptr, const int64_t *rowval,
                             const double *nzval, int64_t base, HostCsr &out);
// Validate a CSC triplet: monotone colptr, indices in range.  Returns 0 or an error code (message set).
int validate_csc(const char *name, int64_t nrows, int64_t ncols, const int64_t *colptr, const int64_t *rowval,
                 const double *nzval, int64_t base);
// tiles of <= tile_nnz non-zeros and <= max_rows rows (spmv_core.cuh: kTileFill, kTileRows)
void build_tiles(const HostCsr &M, int tile_nnz, int max_rows, HostTiles &out);
void assign_tiles(HostTiles &t, int grid);
int choose_lpr(const HostCsr &M);

int host_threads();
void parallel_chunks(int64_t count, const std::function<void(int, int64_t, int64_t)> &fn, int64_t min_chunk);

// Modified Ruiz equilibration of the KKT matrix [P A'; A 0] with cost scaling (Stellato et al., "OSQP", 2020,
// Algorithm 2): P <- c D P D, A <- E A D, q <- c D q.  Works on the row-major copies the solver already holds
// (P symmetric: CSR(P) rows = columns; At = CSR(A')), all three kept consistent.  Not in the reference
// (README.md:71-72 lists it as TODO): SURVEY 8(f) row 1.
struct RuizScaling {
    std::vector<double> D, E;   // n, m
    double c = 1.0;
};
void ruiz_equilibrate(HostCsr &P, HostCsr &A, HostCsr &At, std::vector<double> &q, int iters, RuizScaling &out);

// true iff every entry is finite (vectorisable: v * 0 is NaN exactly for NaN / +-Inf)
bool all_finite(const double *v, size_t count);

// Host -> device copy of a large PAGEABLE array at PCIe speed: the array is copied chunk by chunk into a small ring
// of page-locked buffers by all host threads (checking the values for NaN/Inf in the same pass when `finite_out` is
// given) and each chunk goes out with cudaMemcpyAsync while the next one is being staged.  A plain cudaMemcpy from
// pageable memory measured 11 GB/s on the B200 hosts, 5.5 GB of dense batch data need 0.5 s that way.
// Returns after the last chunk has been queued on `st` (the caller's array may be released: it has been staged).
cudaError_t staged_upload(void *dst_dev, const double *src_host, size_t count, cudaStream_t st, bool *finite_out);

int check_device(int device);   // 0 or QPB200_ERR_DEVICE / QPB200_ERR_CUDA

// ---- partition of one QP over R ranks (dist_partition.cpp)
struct DistSlice {
    int64_t i0 = 0, i1 = 0, j0 = 0, j1 = 0;   // rows of A / columns of P owned by the rank
    int64_t p_off = 0;                        // first non-zero of P the slice uses (offset into the caller's arrays)
    std::vector<int64_t> Pcolptr;             // n + 1, relative to p_off
    std::vector<int64_t> Acolptr, Arowval;    // row slice of A as CSC with local row indices
    std::vector<double> Anzval;
};
void balanced_blocks(const int64_t *counts, int64_t len, int parts, std::vector<int64_t> &bounds);
void dist_plan(int64_t n, int64_t m, const int64_t *Pp, const int64_t *Ap, const int64_t *Ai, int64_t base, int nranks,
               std::vector<int64_t> &rows, std::vector<int64_t> &cols);
void dist_slice(int64_t n, int64_t m, const int64_t *Pp, const int64_t *Ap, const int64_t *Ai, const double *Av, int64_t base,
                int rank, const std::vector<int64_t> &rows, const std::vector<int64_t> &cols, DistSlice &out);
// Host threads of this process are divided by `ranks_on_node` (one process per GPU shares the node's cores)
void set_host_thread_share(int ranks_on_node);

}  // namespace qpb
