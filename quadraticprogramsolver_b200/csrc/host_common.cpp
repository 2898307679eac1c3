// host_common.cpp -- CSC -> tiled-CSR conversion and misc host helpers (no device code).
#include "host_common.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace qpb {

std::string &last_error() {
    static thread_local std::string msg;
    return msg;
}

int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

int validate_csc(const char *name, int64_t nrows, int64_t ncols, const int64_t *colptr, const int64_t *rowval,
                 const double *nzval, int64_t base) {
    if (!colptr) return fail(QPB200_ERR_ARG, "%s: colptr is NULL", name);
    if (colptr[0] != base) return fail(QPB200_ERR_ARG, "%s: colptr[0] = %lld, expected index_base %lld", name,
                                       (long long)colptr[0], (long long)base);
    for (int64_t j = 0; j < ncols; ++j)
        if (colptr[j + 1] < colptr[j]) return fail(QPB200_ERR_ARG, "%s: colptr not monotone at column %lld", name, (long long)j);
    const int64_t nnz = colptr[ncols] - base;
    if (nnz >= (int64_t(1) << 31) - 64) return fail(QPB200_ERR_ARG, "%s: nnz = %lld exceeds the int32 index range", name, (long long)nnz);
    if (nnz > 0 && (!rowval || !nzval)) return fail(QPB200_ERR_ARG, "%s: rowval/nzval is NULL", name);
    for (int64_t k = 0; k < nnz; ++k) {
        const int64_t i = rowval[k] - base;
        if (i < 0 || i >= nrows) return fail(QPB200_ERR_ARG, "%s: row index %lld out of range at nnz %lld", name, (long long)rowval[k], (long long)k);
        if (!std::isfinite(nzval[k])) return fail(QPB200_ERR_NONFINITE, "%s: non-finite value at nnz %lld", name, (long long)k);
    }
    return QPB200_OK;
}

void csc_to_csr(int64_t nrows, int64_t ncols, const int64_t *colptr, const int64_t *rowval, const double *nzval,
                int64_t base, HostCsr &out) {
    const int64_t nnz = colptr[ncols] - base;
    out.rows = (int)nrows;
    out.cols = (int)ncols;
    out.ptr.assign((size_t)nrows + 1, 0);
    out.idx.resize((size_t)nnz);
    out.val.resize((size_t)nnz);
    for (int64_t k = 0; k < nnz; ++k) out.ptr[(size_t)(rowval[k] - base) + 1]++;
    for (int64_t i = 0; i < nrows; ++i) out.ptr[(size_t)i + 1] += out.ptr[(size_t)i];
    std::vector<int> cur(out.ptr.begin(), out.ptr.end() - 1);
    for (int64_t j = 0; j < ncols; ++j)
        for (int64_t k = colptr[j] - base; k < colptr[j + 1] - base; ++k) {
            const int p = cur[(size_t)(rowval[k] - base)]++;
            out.idx[(size_t)p] = (int)j;       // ascending j within each row: columns stay sorted
            out.val[(size_t)p] = nzval[k];
        }
}

void csc_as_csr_of_transpose(int64_t nrows, int64_t ncols, const int64_t *colptr, const int64_t *rowval,
                             const double *nzval, int64_t base, HostCsr &out) {
    const int64_t nnz = colptr[ncols] - base;
    out.rows = (int)ncols;
    out.cols = (int)nrows;
    out.ptr.resize((size_t)ncols + 1);
    out.idx.resize((size_t)nnz);
    out.val.resize((size_t)nnz);
    for (int64_t j = 0; j <= ncols; ++j) out.ptr[(size_t)j] = (int)(colptr[j] - base);
    for (int64_t k = 0; k < nnz; ++k) out.idx[(size_t)k] = (int)(rowval[k] - base);
    if (nnz) std::memcpy(out.val.data(), nzval, (size_t)nnz * sizeof(double));
}

int choose_lpr(const HostCsr &M) {
    const double avg = M.rows > 0 ? (double)M.nnz() / (double)M.rows : 0.0;
    int lpr = 1;
    while (lpr < 32 && avg > 6.0 * lpr) lpr *= 2;
    return lpr;
}

void build_tiles(const HostCsr &M, int tile_nnz, HostTiles &out) {
    const int max_rows = 4096;   // bound on rows per tile (runs of empty rows)
    out.tiles.clear();
    int r = 0;
    while (r < M.rows) {
        const int k0 = M.ptr[(size_t)r];
        const int len = M.ptr[(size_t)r + 1] - k0;
        if (len > tile_nnz) {   // long row: consecutive segments, flagged, kept on one CTA
            int off = 0;
            while (off < len) {
                const int nk = std::min(tile_nnz, len - off);
                int w = nk;
                if (off > 0) w |= (1 << 30);
                if (off + nk < len) w |= (1 << 29);
                out.tiles.push_back(make_int4(r, 1, k0 + off, w));
                off += nk;
            }
            ++r;
            continue;
        }
        int r1 = r, nk = 0;
        while (r1 < M.rows && (r1 - r) < max_rows) {
            const int l1 = M.ptr[(size_t)r1 + 1] - M.ptr[(size_t)r1];
            if (l1 > tile_nnz || nk + l1 > tile_nnz) break;
            nk += l1;
            ++r1;
        }
        out.tiles.push_back(make_int4(r, r1 - r, k0, nk));
        r = r1;
    }
    out.lpr = choose_lpr(M);
}

void assign_tiles(HostTiles &t, int grid) {
    const size_t nt = t.tiles.size();
    std::vector<double> cost(nt);
    double total = 0.0;
    for (size_t i = 0; i < nt; ++i) {
        const int nk = t.tiles[i].w & ((1 << 24) - 1);
        cost[i] = (double)nk + 2.0 * t.tiles[i].y + 64.0;
        total += cost[i];
    }
    t.cta_begin.assign((size_t)grid + 1, (int)nt);
    t.cta_begin[0] = 0;
    size_t i = 0;
    double acc = 0.0;
    for (int b = 0; b < grid; ++b) {
        t.cta_begin[(size_t)b] = (int)i;
        const double target = total * (double)(b + 1) / (double)grid;
        while (i < nt && acc + 0.5 * cost[i] <= target) acc += cost[i++];
        while (i < nt && (t.tiles[i].w & (1 << 30))) acc += cost[i++];   // never split a long row
    }
    t.cta_begin[(size_t)grid] = (int)nt;
    // anything left (rounding) goes to the last CTA
}

bool all_finite(const double *v, size_t count) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t i = 0;
    for (; i + 8 <= count; i += 8)
        for (int j = 0; j < 8; ++j) acc[j] += v[i + j] * 0.0;
    for (; i < count; ++i) acc[0] += v[i] * 0.0;
    double s = 0.0;
    for (int j = 0; j < 8; ++j) s += acc[j];
    return s == 0.0;
}

int check_device(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(QPB200_ERR_DEVICE, "no CUDA device available (%s); libqpb200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0) {
        e = cudaGetDevice(&device);
        if (e != cudaSuccess) return fail(QPB200_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    }
    if (device >= count) return fail(QPB200_ERR_ARG, "device %d out of range (%d visible)", device, count);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(QPB200_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(QPB200_ERR_DEVICE, "device %d (%s) is sm_%d%d; libqpb200 is built for sm_100a only", device, prop.name,
                    prop.major, prop.minor);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(QPB200_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return QPB200_OK;
}

}  // namespace qpb
