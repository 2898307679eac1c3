// host_common.cpp -- CSC -> tiled-CSR conversion and misc host helpers (no device code).
#include "host_common.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

namespace qpb {

static std::atomic<int> g_thread_share{1};
void set_host_thread_share(int ranks_on_node) { g_thread_share = std::max(1, ranks_on_node); }

// QPB200_HOST_THREADS, else OMP_NUM_THREADS when it asks for more than one thread (torch.distributed.run exports
// OMP_NUM_THREADS=1 for every rank: that is a default, not a request), else all hardware threads; capped at 32 and
// divided by the number of ranks that share the node (qpb200_dist_create*)
int host_threads() {
    static int nt = [] {
        const char *e = getenv("QPB200_HOST_THREADS");
        const char *o = getenv("OMP_NUM_THREADS");
        int v = e ? atoi(e) : ((o && atoi(o) > 1) ? atoi(o) : (int)std::thread::hardware_concurrency());
        return std::max(1, std::min(v, 32));
    }();
    if (getenv("QPB200_HOST_THREADS")) return nt;
    return std::max(1, std::min(nt, (int)std::thread::hardware_concurrency() / g_thread_share.load()));
}

// fn(t, begin, end) on contiguous chunks of [0, count) -- plain std::thread, no OpenMP dependency
void parallel_chunks(int64_t count, const std::function<void(int, int64_t, int64_t)> &fn, int64_t min_chunk) {
    int nt = host_threads();
    if (count < 2 * min_chunk) nt = 1;
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, count / std::max<int64_t>(1, min_chunk)));
    if (nt <= 1) {
        fn(0, 0, count);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) {
        const int64_t b = count * t / nt, e = count * (t + 1) / nt;
        th.emplace_back([&fn, t, b, e] { fn(t, b, e); });
    }
    for (auto &x : th) x.join();
}

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
    std::atomic<int64_t> bad_idx(-1), bad_val(-1);
    parallel_chunks(nnz, [&](int, int64_t b, int64_t e) {
        double acc = 0.0;
        for (int64_t k = b; k < e; ++k) {
            const int64_t i = rowval[k] - base;
            if (i < 0 || i >= nrows) { bad_idx = k; return; }
            acc += nzval[k] * 0.0;
        }
        if (acc != 0.0) bad_val = b;
    }, 1 << 16);
    if (bad_idx >= 0) return fail(QPB200_ERR_ARG, "%s: row index %lld out of range at nnz %lld", name, (long long)rowval[bad_idx], (long long)bad_idx.load());
    if (bad_val >= 0) {
        int64_t k = bad_val;
        while (k < nnz && std::isfinite(nzval[k])) ++k;
        return fail(QPB200_ERR_NONFINITE, "%s: non-finite value at nnz %lld", name, (long long)k);
    }
    return QPB200_OK;
}

namespace {
// Stable counting sort on one thread (small matrices): rows keep ascending column order.
void csc_to_csr_serial(int64_t nrows, int64_t ncols, const int64_t *colptr, const int64_t *rowval, const double *nzval,
                       int64_t base, const int64_t *gap, HostCsr &out) {
    std::vector<int> cur((size_t)nrows + 1, 0);
    const int64_t nnz = colptr[ncols] - base;
    for (int64_t k = 0; k < nnz; ++k) cur[(size_t)(rowval[k] - base) + 1]++;
    if (gap) out.mid.resize((size_t)nrows);
    int pos = 0;
    for (int64_t i = 0; i < nrows; ++i) {
        const int cnt = cur[(size_t)i + 1];
        out.ptr[(size_t)i] = pos;
        cur[(size_t)i] = pos;
        pos += cnt;
        if (gap) {
            out.mid[(size_t)i] = pos;
            pos += (int)(gap[i + 1] - gap[i]);
        }
    }
    out.ptr[(size_t)nrows] = pos;
    for (int64_t j = 0; j < ncols; ++j)
        for (int64_t k = colptr[j] - base; k < colptr[j + 1] - base; ++k) {
            const int p = cur[(size_t)(rowval[k] - base)]++;
            out.idx[(size_t)p] = (int)j;
            out.val[(size_t)p] = nzval[k];
        }
}

struct StagedEntry {   // one non-zero on its way from column-major to row-major order
    int row, col;
    double val;
};
}  // namespace

void csc_to_csr(int64_t nrows, int64_t ncols, const int64_t *colptr, const int64_t *rowval, const double *nzval,
                int64_t base, HostCsr &out, const int64_t *gap) {
    // Two-pass bucketed transpose.  A plain counting sort scatters every non-zero to a random cache line of the
    // output (cfg5: 26 M entries -> ~100 ms on 16 cores); here pass A partitions the entries by row block into a
    // staging array with <= 512 sequential write streams per thread, and pass B finishes each row block inside a
    // cache-sized window of the output.  Thread t owns a contiguous block of columns and the staging array is ordered
    // (row block, thread), so every row keeps ascending column order whatever the thread count (deterministic layout).
    // gap != nullptr (rows + 1 offsets): row r is followed by gap[r+1] - gap[r] free slots and out.mid[r] marks
    // where they start -- H = [P A'] is assembled in place this way.
    const int64_t nnz = colptr[ncols] - base;
    const int64_t gap_total = gap ? gap[nrows] - gap[0] : 0;
    out.rows = (int)nrows;
    out.cols = (int)ncols;
    out.ptr.assign((size_t)nrows + 1, 0);
    out.idx.resize((size_t)(nnz + gap_total));
    out.val.resize((size_t)(nnz + gap_total));
    int nt = host_threads();
    if (nnz < (1 << 18) || nt == 1) {
        csc_to_csr_serial(nrows, ncols, colptr, rowval, nzval, base, gap, out);
        return;
    }
    int shift = 0;
    while ((nrows >> shift) > 512) ++shift;
    const int64_t nb = ((nrows - 1) >> shift) + 1;
    // column block boundaries balanced by nnz
    std::vector<int64_t> cb((size_t)nt + 1, ncols);
    cb[0] = 0;
    for (int t = 1; t < nt; ++t) {
        const int64_t target = base + nnz * t / nt;
        cb[(size_t)t] = std::lower_bound(colptr, colptr + ncols + 1, target) - colptr;
        if (cb[(size_t)t] > ncols) cb[(size_t)t] = ncols;
        if (cb[(size_t)t] < cb[(size_t)t - 1]) cb[(size_t)t] = cb[(size_t)t - 1];
    }
    auto run = [&](const std::function<void(int)> &f) {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(f, t);
        for (auto &x : th) x.join();
    };
    // pass 0: entries per (thread, row block)
    std::vector<int64_t> off((size_t)nt * (size_t)nb, 0);
    run([&](int t) {
        int64_t *h = off.data() + (size_t)t * (size_t)nb;
        for (int64_t k = colptr[cb[(size_t)t]] - base; k < colptr[cb[(size_t)t + 1]] - base; ++k) h[(rowval[k] - base) >> shift]++;
    });
    std::vector<int64_t> boff((size_t)nb + 1, 0);
    {
        int64_t pos = 0;
        for (int64_t b = 0; b < nb; ++b) {
            boff[(size_t)b] = pos;
            for (int t = 0; t < nt; ++t) {
                const int64_t c = off[(size_t)t * (size_t)nb + (size_t)b];
                off[(size_t)t * (size_t)nb + (size_t)b] = pos;
                pos += c;
            }
        }
        boff[(size_t)nb] = pos;
    }
    // pass A: partition into the staging array
    PodBuf<StagedEntry> stage;
    stage.resize((size_t)nnz);
    run([&](int t) {
        int64_t *o = off.data() + (size_t)t * (size_t)nb;
        StagedEntry *st = stage.data();
        for (int64_t j = cb[(size_t)t]; j < cb[(size_t)t + 1]; ++j)
            for (int64_t k = colptr[j] - base; k < colptr[j + 1] - base; ++k) {
                const int64_t i = rowval[k] - base;
                st[o[i >> shift]++] = StagedEntry{(int)i, (int)j, nzval[k]};
            }
    });
    // pass B: one row block at a time (dynamic schedule: block sizes follow the row lengths)
    if (gap) out.mid.resize((size_t)nrows);
    std::atomic<int64_t> next(0);
    run([&](int) {
        std::vector<int> cur((size_t)1 << shift);
        for (;;) {
            const int64_t b = next.fetch_add(1);
            if (b >= nb) break;
            const int64_t r0 = b << shift, r1 = std::min<int64_t>(nrows, r0 + ((int64_t)1 << shift));
            const StagedEntry *st = stage.data();
            std::fill(cur.begin(), cur.begin() + (r1 - r0), 0);
            for (int64_t e = boff[(size_t)b]; e < boff[(size_t)b + 1]; ++e) cur[(size_t)(st[e].row - r0)]++;
            int64_t pos = boff[(size_t)b] + (gap ? gap[r0] - gap[0] : 0);
            for (int64_t r = r0; r < r1; ++r) {
                const int cnt = cur[(size_t)(r - r0)];
                out.ptr[(size_t)r] = (int)pos;
                cur[(size_t)(r - r0)] = (int)pos;
                pos += cnt;
                if (gap) {
                    out.mid[(size_t)r] = (int)pos;
                    pos += gap[r + 1] - gap[r];
                }
            }
            for (int64_t e = boff[(size_t)b]; e < boff[(size_t)b + 1]; ++e) {
                const int p = cur[(size_t)(st[e].row - r0)]++;
                out.idx[(size_t)p] = st[e].col;
                out.val[(size_t)p] = st[e].val;
            }
        }
    });
    out.ptr[(size_t)nrows] = (int)(nnz + gap_total);
}

void assemble_h_direct(int64_t n, int64_t m, const int64_t *Pp, const int64_t *Pi, const double *Pv, const int64_t *Ap,
                       const int64_t *Ai, const double *Av, int64_t base, HostCsr &H, std::vector<double> &dP,
                       std::vector<double> &dAA) {
    csc_to_csr(n, n, Pp, Pi, Pv, base, H, Ap);   // rows of P, each followed by room for the matching column of A
    H.cols = (int)(n + m);
    dP.assign((size_t)n, 0.0);
    dAA.assign((size_t)n, 0.0);
    parallel_chunks(n, [&](int, int64_t j0, int64_t j1) {
        for (int64_t j = j0; j < j1; ++j) {
            for (int k = H.ptr[(size_t)j]; k < H.mid[(size_t)j]; ++k)
                if (H.idx[(size_t)k] == j) dP[(size_t)j] += H.val[(size_t)k];
            int pos = H.mid[(size_t)j];
            for (int64_t k = Ap[j] - base; k < Ap[j + 1] - base; ++k) {
                H.idx[(size_t)pos] = (int)(Ai[k] - base + n);
                H.val[(size_t)pos] = Av[k];
                dAA[(size_t)j] += Av[k] * Av[k];
                ++pos;
            }
        }
    }, 4096);
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
    parallel_chunks(nnz, [&](int, int64_t b, int64_t e) {
        for (int64_t k = b; k < e; ++k) out.idx[(size_t)k] = (int)(rowval[k] - base);
        std::memcpy(out.val.data() + b, nzval + b, (size_t)(e - b) * sizeof(double));
    }, 1 << 16);
}

// Lanes per row of the row-sum step.  One lane per row costs the fewest instructions (no shuffles, every lane of a
// warp runs the phase epilogue) but chains `avg` dependent shared-memory adds; lpr ~ avg / 8 keeps the chain at 8-16
// adds (two accumulators) while a 1016-nnz tile still fills at most one round of the 256 threads.
int choose_lpr(const HostCsr &M) {
    static const int forced = [] {
        const char *e = getenv("QPB200_LPR");   // A/B experiments only
        return e ? atoi(e) : 0;
    }();
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8 || forced == 16 || forced == 32) return forced;
    const double avg = M.rows > 0 ? (double)M.nnz() / (double)M.rows : 0.0;
    int lpr = 1;
    while (lpr < 32 && avg >= 16.0 * lpr) lpr *= 2;
    return lpr;
}

void build_tiles(const HostCsr &M, int tile_nnz, int max_rows, HostTiles &out) {
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
    std::atomic<int> bad(0);
    parallel_chunks((int64_t)count, [&](int, int64_t b, int64_t e) {
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int64_t i = b;
        for (; i + 8 <= e; i += 8)
            for (int j = 0; j < 8; ++j) acc[j] += v[i + j] * 0.0;
        for (; i < e; ++i) acc[0] += v[i] * 0.0;
        double s = 0.0;
        for (int j = 0; j < 8; ++j) s += acc[j];
        if (s != 0.0) bad = 1;
    }, 1 << 18);
    return bad == 0;
}

// ---- staged uploads ---------------------------------------------------------------------------------
namespace {
struct StagingRing {
    static constexpr int kSlots = 4;
    static constexpr size_t kChunk = (size_t)64 << 20;   // bytes per slot
    std::mutex mu;                                         // one upload at a time per device
    double *slot[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t done[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    bool ok = false, tried = false;
    bool init() {
        if (tried) return ok;
        tried = true;
        for (int i = 0; i < kSlots; ++i) {
            if (cudaHostAlloc((void **)&slot[i], kChunk, cudaHostAllocPortable) != cudaSuccess ||
                cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                return ok = false;
            }
        }
        return ok = true;
    }
};
// events belong to a device: one ring per device ordinal (a process normally drives one GPU)
StagingRing &staging_ring(int dev) { static StagingRing r[16]; return r[dev & 15]; }
}  // namespace

cudaError_t staged_upload(void *dst_dev, const double *src_host, size_t count, cudaStream_t st, bool *finite_out) {
    if (finite_out) *finite_out = true;
    const size_t bytes = count * sizeof(double);
    int dev = 0;
    cudaGetDevice(&dev);
    StagingRing &ring = staging_ring(dev);
    std::unique_lock<std::mutex> lock(ring.mu);
    if (bytes < 2 * StagingRing::kChunk || !ring.init()) {          // small array or no page-locked memory: plain copy
        lock.unlock();
        if (finite_out) *finite_out = all_finite(src_host, count);
        return cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st);
    }
    const size_t per = StagingRing::kChunk / sizeof(double);
    std::atomic<int> bad(0);
    size_t off = 0;
    for (int i = 0; off < count; ++i, off += per) {
        const int sl = i % StagingRing::kSlots;
        const size_t cnt = std::min(per, count - off);
        if (i >= StagingRing::kSlots) {                             // the slot's previous chunk must have left
            cudaError_t e = cudaEventSynchronize(ring.done[sl]);
            if (e != cudaSuccess) return e;
        }
        double *dst = ring.slot[sl];
        const double *src = src_host + off;
        parallel_chunks((int64_t)cnt, [&](int, int64_t b, int64_t e) {
            std::memcpy(dst + b, src + b, (size_t)(e - b) * sizeof(double));
            if (finite_out) {
                double acc[4] = {0, 0, 0, 0};
                int64_t k = b;
                for (; k + 4 <= e; k += 4)
                    for (int j = 0; j < 4; ++j) acc[j] += dst[k + j] * 0.0;
                for (; k < e; ++k) acc[0] += dst[k] * 0.0;
                if ((acc[0] + acc[1]) + (acc[2] + acc[3]) != 0.0) bad = 1;
            }
        }, 1 << 16);
        cudaError_t e = cudaMemcpyAsync(static_cast<double *>(dst_dev) + off, dst, cnt * sizeof(double), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
        e = cudaEventRecord(ring.done[sl], st);
        if (e != cudaSuccess) return e;
    }
    // the ring is shared: leave it drained so that the next upload (possibly on another stream) starts clean
    for (int sl = 0; sl < StagingRing::kSlots; ++sl) {
        cudaError_t e = cudaEventSynchronize(ring.done[sl]);
        if (e != cudaSuccess) return e;
    }
    if (finite_out) *finite_out = bad == 0;
    return cudaSuccess;
}

// ---- block caches -----------------------------------------------------------------------------------
namespace {
struct BlockCache {
    std::mutex mu;
    std::vector<std::pair<void *, size_t>> blocks;   // (pointer, capacity)
    size_t total = 0;
    int device = -1;                                  // device cache: valid for one device at a time
};
size_t cache_limit() {
    static size_t lim = [] {
        const char *e = getenv("QPB200_CACHE_MB");
        return (size_t)(e ? atoll(e) : 4096) << 20;
    }();
    return lim;
}
constexpr size_t kCacheMinBlock = 1 << 20;   // only blocks >= 1 MB are worth keeping
BlockCache &dev_cache() { static BlockCache c; return c; }
BlockCache &host_cache() { static BlockCache c; return c; }
BlockCache &pinned_cache() { static BlockCache c; return c; }

void *take(BlockCache &c, size_t bytes, size_t *cap) {
    std::lock_guard<std::mutex> g(c.mu);
    size_t best = (size_t)-1, bi = 0;
    for (size_t i = 0; i < c.blocks.size(); ++i) {
        const size_t k = c.blocks[i].second;
        if (k >= bytes && k <= bytes + bytes / 4 + (1 << 20) && k < best) { best = k; bi = i; }
    }
    if (best == (size_t)-1) return nullptr;
    void *p = c.blocks[bi].first;
    *cap = best;
    c.total -= best;
    c.blocks.erase(c.blocks.begin() + (long)bi);
    return p;
}
size_t pinned_cache_limit() {          // page-locked memory is a scarcer resource: its own, smaller bound
    static size_t lim = [] {
        const char *e = getenv("QPB200_PINNED_CACHE_MB");
        return (size_t)(e ? atoll(e) : 1536) << 20;
    }();
    return std::min(lim, cache_limit());
}
bool give(BlockCache &c, void *p, size_t cap, size_t min_block, size_t limit = (size_t)-1) {
    if (cap < min_block) return false;
    std::lock_guard<std::mutex> g(c.mu);
    if (c.total + cap > std::min(limit, cache_limit())) return false;
    c.blocks.push_back({p, cap});
    c.total += cap;
    return true;
}
}  // namespace

cudaError_t cached_device_alloc(void **p, size_t bytes, size_t *capacity) {
    int dev = -1;
    cudaGetDevice(&dev);
    BlockCache &c = dev_cache();
    {
        std::lock_guard<std::mutex> g(c.mu);
        if (c.device != dev) {          // cached blocks belong to another device: drop them
            for (auto &b : c.blocks) cudaFree(b.first);
            c.blocks.clear();
            c.total = 0;
            c.device = dev;
        }
    }
    if (void *q = take(c, bytes, capacity)) {
        *p = q;
        return cudaSuccess;
    }
    *capacity = bytes;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {
        // out of memory with blocks parked in the cache: release them and retry once
        cudaGetLastError();
        {
            std::lock_guard<std::mutex> g(c.mu);
            for (auto &b : c.blocks) cudaFree(b.first);
            c.blocks.clear();
            c.total = 0;
        }
        e = cudaMalloc(p, bytes);
    }
    return e;
}

void cached_device_free(void *p, size_t capacity) {
    if (!p) return;
    int dev = -1;
    cudaGetDevice(&dev);
    BlockCache &c = dev_cache();
    if (dev == c.device && give(c, p, capacity, 0)) return;   // small blocks too: a real cudaFree can stall for tens of ms
    static const bool timing = getenv("QPB200_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    cudaFree(p);
    if (timing)
        fprintf(stderr, "[qpb200 cache] cudaFree of %.2f MB took %.2f ms (cache holds %.0f MB, device %d/%d)\n", capacity / 1048576.0,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), c.total / 1048576.0, dev,
                c.device);
}

void *cached_host_alloc(size_t bytes, size_t *capacity) {
    if (bytes >= kCacheMinBlock) {
        void *q = take(host_cache(), bytes, capacity);
        if (q) return q;
    }
    *capacity = bytes;
    return malloc(bytes);
}

void cached_host_free(void *p, size_t capacity) {
    if (!p) return;
    if (give(host_cache(), p, capacity, kCacheMinBlock)) return;
    free(p);
}

void *cached_pinned_alloc(size_t bytes, size_t *capacity, bool *pinned) {
    *pinned = false;
    if (bytes < kCacheMinBlock) {                    // small arrays: pinning costs more than it saves
        *capacity = bytes;
        return malloc(bytes);
    }
    if (void *q = take(pinned_cache(), bytes, capacity)) {
        *pinned = true;
        return q;
    }
    void *q = nullptr;
    if (cudaHostAlloc(&q, bytes, cudaHostAllocPortable) == cudaSuccess) {
        *capacity = bytes;
        *pinned = true;
        return q;
    }
    cudaGetLastError();                              // not fatal: fall back to pageable memory
    return cached_host_alloc(bytes, capacity);
}

void cached_pinned_free(void *p, size_t capacity, bool pinned) {
    if (!p) return;
    if (!pinned) {
        cached_host_free(p, capacity);
        return;
    }
    if (give(pinned_cache(), p, capacity, kCacheMinBlock, pinned_cache_limit())) return;
    cudaFreeHost(p);
}

namespace {
inline double limit_scaling(double v) { return v < 1e-4 ? 1.0 : (v > 1e4 ? 1e4 : v); }

void row_max_abs(const HostCsr &M, std::vector<double> &out, bool accumulate) {
    parallel_chunks(M.rows, [&](int, int64_t r0, int64_t r1) {
        for (int64_t r = r0; r < r1; ++r) {
            double v = accumulate ? out[(size_t)r] : 0.0;
            for (int k = M.ptr[(size_t)r]; k < M.ptr[(size_t)r + 1]; ++k) v = std::fmax(v, std::fabs(M.val.p[k]));
            out[(size_t)r] = v;
        }
    }, 4096);
}

// M[r, c] *= rs[r] * cs[c]
void scale_rows_cols(HostCsr &M, const std::vector<double> &rs, const std::vector<double> &cs, bool row_first) {
    parallel_chunks(M.rows, [&](int, int64_t r0, int64_t r1) {
        for (int64_t r = r0; r < r1; ++r)
            for (int k = M.ptr[(size_t)r]; k < M.ptr[(size_t)r + 1]; ++k) {
                const double a = rs[(size_t)r], b = cs[(size_t)M.idx.p[k]];
                // the same factor (variable scale * constraint scale) for A[i,j] and A'[j,i]: both copies stay bit-identical
                M.val.p[k] *= row_first ? (a * b) : (b * a);
            }
    }, 4096);
}
}  // namespace

void ruiz_equilibrate(HostCsr &P, HostCsr &A, HostCsr &At, std::vector<double> &q, int iters, RuizScaling &out) {
    const size_t n = (size_t)P.rows, m = (size_t)A.rows;
    out.D.assign(n, 1.0);
    out.E.assign(m, 1.0);
    out.c = 1.0;
    std::vector<double> dn(n), en(m);
    for (int it = 0; it < iters; ++it) {
        row_max_abs(P, dn, false);            // column norms of [P; A]
        if (m) row_max_abs(At, dn, true);
        if (m) row_max_abs(A, en, false);     // column norms of [A'; 0]
        for (size_t j = 0; j < n; ++j) dn[j] = 1.0 / std::sqrt(limit_scaling(dn[j]));
        for (size_t i = 0; i < m; ++i) en[i] = 1.0 / std::sqrt(limit_scaling(en[i]));
        scale_rows_cols(P, dn, dn, true);
        if (m) {
            scale_rows_cols(A, en, dn, false);    // factor d_j * e_i
            scale_rows_cols(At, dn, en, true);    // factor d_j * e_i
        }
        for (size_t j = 0; j < n; ++j) {
            q[j] *= dn[j];
            out.D[j] *= dn[j];
        }
        for (size_t i = 0; i < m; ++i) out.E[i] *= en[i];
        // cost scaling: mean column norm of P against |q|inf
        row_max_abs(P, dn, false);
        double mean = 0.0, qn = 0.0;
        for (size_t j = 0; j < n; ++j) {
            mean += dn[j];
            qn = std::fmax(qn, std::fabs(q[j]));
        }
        mean = limit_scaling(mean / (double)n);
        qn = limit_scaling(qn);
        const double gamma = 1.0 / std::fmax(mean, qn);
        parallel_chunks(P.rows, [&](int, int64_t r0, int64_t r1) {
            for (int64_t k = P.ptr[(size_t)r0]; k < P.ptr[(size_t)r1]; ++k) P.val.p[k] *= gamma;
        }, 4096);
        for (size_t j = 0; j < n; ++j) q[j] *= gamma;
        out.c *= gamma;
    }
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
    // attribute queries, not cudaGetDeviceProperties: the latter costs milliseconds (up to 100+ with other driver clients)
    int major = 0, minor = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
    if (e != cudaSuccess) return fail(QPB200_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    if (major != 10)
        return fail(QPB200_ERR_DEVICE, "device %d is sm_%d%d; libqpb200 is built for sm_100a only", device, major, minor);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(QPB200_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return QPB200_OK;
}

}  // namespace qpb
