// dist_solver.cu -- host side of the row-partitioned multi-GPU sparse path (one rank per GPU).
// NCCL is loaded at run time (dlopen "libnccl.so.2"), so libqpb200.so itself has no NCCL dependency and
// shares the NCCL a host application (e.g. torch) has already loaded.
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cmath>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "peer_kernels.cuh"
#include "host_common.h"
#include "sparse_solver.h"

namespace qpb {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) return nullptr;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
    api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
    if (!api.AllGather || !api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.GetErrorString) {
        dlclose(api.lib);
        api.lib = nullptr;
        return nullptr;
    }
    return &api;
}

#define QPB_NCCL(call)                                                                                          \
    do {                                                                                                        \
        ncclResult_t r_ = (call);                                                                               \
        if (r_ != ncclSuccess)                                                                                  \
            return ::qpb::fail(QPB200_ERR_NCCL, "%s failed: %s (%s:%d)", #call, api->GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

// Process-wide caches, one entry per device (one rank per GPU, whether the ranks are processes or threads of one
// process), guarded by a mutex and reference counted: an entry that a live handle uses is never destroyed or
// replaced -- a handle that cannot take the entry gets private resources it owns and frees itself.
constexpr int kMaxDevices = 64;
static std::mutex &cache_mutex() {
    static std::mutex m;
    return m;
}
struct CommCacheEntry {
    ncclComm_t comm = nullptr;
    int rank = -1, nranks = 0, users = 0;
};
static CommCacheEntry &comm_cache(int device) {
    static CommCacheEntry e[kMaxDevices];
    return e[device];
}

// Peer-visible region of this process, kept across handles like the communicator: cudaMalloc + IPC export +
// (R-1) cudaIpcOpenMemHandle at create and the matching close/free at destroy cost 100-500 ms per handle --
// as much as a whole solve on 8 GPUs.  One entry per device; a handle created while another one owns it allocates
// privately.
struct PeerRegionCache {
    bool in_use = false;
    int rank = -1, nranks = 0;
    double *region = nullptr;
    size_t doubles = 0;
    void *opened[kMaxPeers] = {nullptr};
};
static PeerRegionCache &region_cache(int device) {
    static PeerRegionCache c[kMaxDevices];
    return c[device];
}

struct DistContext {
    int rank = 0, nranks = 1, device = -1;
    ncclComm_t comm = nullptr;
    bool comm_cached = false;              // comm belongs to comm_cache(device) (users counted), else this handle owns it
    int64_t row_begin = 0, row_end = 0;    // rows of A this rank owns (qpb200_dist_create_full; -1 when pre-sliced)
    DistBuffers buf{};
    DistState *host_state = nullptr;   // pinned mirror (2 slots: the CG loop runs one step ahead of the host)
    cudaEvent_t ev[2] = {nullptr, nullptr};
    long long launches = 0, allreduces = 0;
    // in-kernel peer-memory path (peer_kernels.cuh)
    bool peer_ok = false;
    PeerDev peer{};
    double *region = nullptr;              // this rank's peer-visible allocation (not in the arena: IPC-exported)
    void *opened[kMaxPeers] = {nullptr};   // cudaIpcOpenMemHandle results
    size_t region_doubles = 0;
    bool region_cached = false;            // region / opened[] belong to region_cache(): handed back, not freed
    double *tiny = nullptr;                // 1 double for the host-level rendezvous all-reduce
};

void dist_destroy(DistContext *d) {
    if (!d) return;
    NcclApi *api = nccl_api();
    if (d->host_state) cudaFreeHost(d->host_state);
    for (auto &e : d->ev) if (e) cudaEventDestroy(e);
    std::lock_guard<std::mutex> g(cache_mutex());
    // the cached communicator outlives the handle (creating one costs ~1 s, re-solves / new handles reuse it);
    // a private one goes with its handle
    if (d->comm_cached) comm_cache(d->device).users -= 1;
    else if (d->comm && api) api->CommDestroy(d->comm);
    if (d->region_cached) {
        region_cache(d->device).in_use = false;     // mappings stay open for the next handle on this device
    } else {
        for (void *o : d->opened) if (o) cudaIpcCloseMemHandle(o);
        if (d->region) cudaFree(d->region);
    }
    delete d;
}

static int launch_seg(SparseSolver &s, DistContext &d, int seg, int do_check) {
    void *args[] = {(void *)&s.prob, (void *)&d.buf, (void *)&seg, (void *)&do_check};
    const void *fn = s.use_pre ? (const void *)admm_dist_kernel<1, true> : (const void *)admm_dist_kernel<1, false>;
    QPB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(s.grid), dim3(kThreads), args, sizeof(SpmvSmem), s.stream));
    ++d.launches;
    return QPB200_OK;
}

static int read_state(SparseSolver &s, DistContext &d) {
    QPB_CUDA(cudaMemcpyAsync(d.host_state, d.buf.state, sizeof(DistState), cudaMemcpyDeviceToHost, s.stream));
    QPB_CUDA(cudaStreamSynchronize(s.stream));
    return QPB200_OK;
}

// asynchronous variant: copy the control block into pinned slot `slot` and mark it with an event
static int post_read(SparseSolver &s, DistContext &d, int slot) {
    QPB_CUDA(cudaMemcpyAsync(d.host_state + slot, d.buf.state, sizeof(DistState), cudaMemcpyDeviceToHost, s.stream));
    QPB_CUDA(cudaEventRecord(d.ev[slot], s.stream));
    return QPB200_OK;
}

int dist_solve(SparseSolver &s, DistContext &d, double *x_inout, double *z_out, double *y_out, qpb200_info *info) {
    NcclApi *api = nccl_api();
    if (!api) return fail(QPB200_ERR_NCCL, "libnccl.so.2 could not be loaded");
    if (!x_inout) return fail(QPB200_ERR_ARG, "qpb200_dist_solve: x_inout is NULL");
    QPB_CUDA(cudaSetDevice(s.device));
    const int n = s.n, m = s.m;
    int rc = s.reset_state(x_inout);
    if (rc) return rc;
    DistState init;
    std::memset(&init, 0, sizeof(init));
    init.conv_flag = 1;
    init.rho = s.prob.s.rho;
    init.rho1 = 1.0 / init.rho;
    init.rhorho = init.rho;
    init.res_prim = NAN;
    init.res_dual = NAN;
    *d.host_state = init;
    QPB_CUDA(cudaMemcpyAsync(d.buf.state, d.host_state, sizeof(DistState), cudaMemcpyHostToDevice, s.stream));
    QPB_CUDA(cudaStreamSynchronize(s.stream));
    d.launches = 0;
    d.allreduces = 0;
    const long long max_iter = s.prob.s.max_iter, check_every = s.prob.s.check_every;
    QPB_CUDA(cudaEventRecord(s.ev0, s.stream));
    long long ii = 0;
    for (ii = 1; ii <= max_iter; ++ii) {
        if ((rc = launch_seg(s, d, kSegBegin, 0))) return rc;
        QPB_NCCL(api->AllReduce(d.buf.wbuf, d.buf.wbuf, (size_t)n, ncclDouble, ncclSum, d.comm, s.stream));
        ++d.allreduces;
        if ((rc = launch_seg(s, d, kSegPcgInit, 0))) return rc;
        // CG loop, run ONE STEP AHEAD of the host: step j+1 (all-reduce + segment + read-back) is enqueued
        // before the host has seen the `cont` flag of step j, so the host round trip never idles the GPU.
        // A step enqueued after the solve has converged finds cont == 0 on the device and returns at once
        // (its all-reduce ran on a dead buffer: one wasted 8 MB all-reduce per ADMM iteration).
        if ((rc = post_read(s, d, 0))) return rc;
        for (int j = 0;; ++j) {
            QPB_NCCL(api->AllReduce(d.buf.wbuf, d.buf.wbuf, (size_t)n, ncclDouble, ncclSum, d.comm, s.stream));
            ++d.allreduces;
            if ((rc = launch_seg(s, d, kSegPcgStep, 0))) return rc;
            if ((rc = post_read(s, d, (j + 1) & 1))) return rc;
            QPB_CUDA(cudaEventSynchronize(d.ev[j & 1]));
            if (!d.host_state[j & 1].cont) break;     // state after step j-1 (j = 0: after the init segment)
        }
        const int do_check = (ii % check_every) == 0;
        if ((rc = launch_seg(s, d, kSegUpdate, do_check))) return rc;
        if (do_check) {
            QPB_NCCL(api->AllReduce(d.buf.wbuf2, d.buf.wbuf2, (size_t)2 * n, ncclDouble, ncclSum, d.comm, s.stream));
            QPB_NCCL(api->AllReduce(d.buf.state->lmax, d.buf.state->lmax, 4, ncclDouble, ncclMax, d.comm, s.stream));
            d.allreduces += 2;
            if ((rc = launch_seg(s, d, kSegCheck, 1))) return rc;
            if ((rc = read_state(s, d))) return rc;
            if (d.host_state->conv_flag != 1) break;
        }
    }
    if (ii > max_iter) ii = max_iter;
    QPB_CUDA(cudaEventRecord(s.ev1, s.stream));
    if ((rc = read_state(s, d))) return rc;
    QPB_CUDA(cudaMemcpyAsync(x_inout, s.prob.XY, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    if (z_out && m) QPB_CUDA(cudaMemcpyAsync(z_out, s.prob.z, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    if (y_out && m) QPB_CUDA(cudaMemcpyAsync(y_out, s.prob.XY + n, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    QPB_CUDA(cudaStreamSynchronize(s.stream));
    float ms = 0.f;
    QPB_CUDA(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
    const DistState &S = *d.host_state;
    s.last_info.conv_flag = S.conv_flag;
    s.last_info.iterations = ii;
    s.last_info.pcg_iters_total = S.pcg_total;
    s.last_info.n_h_passes = S.n_h;
    s.last_info.n_a_passes = S.n_a;
    if (info) {
        std::memset(info, 0, sizeof(*info));
        info->conv_flag = S.conv_flag;
        info->iterations = ii;
        info->rho_final = S.rho;
        info->res_prim = S.res_prim;
        info->res_dual = S.res_dual;
        info->rho_updates = S.rho_updates;
        info->pcg_iters_total = S.pcg_total;
        info->pcg_maxed = S.pcg_maxed;
        info->solve_ms = ms;
        info->setup_ms = s.setup_ms;
        info->kernel_launches = d.launches;
    }
    return QPB200_OK;
}

static int peer_init(SparseSolver &s, DistContext &d);

int dist_init(SparseSolver &s, DistContext *&out, int rank, int nranks, const void *unique_id) {
    NcclApi *api = nccl_api();
    if (!api) return fail(QPB200_ERR_NCCL, "libnccl.so.2 could not be loaded (dlopen)");
    if (nranks < 1 || rank < 0 || rank >= nranks || !unique_id) return fail(QPB200_ERR_ARG, "qpb200_dist_create: bad rank / nranks / id");
    DistContext *d = new (std::nothrow) DistContext();
    if (!d) return fail(QPB200_ERR_ARG, "out of host memory");
    out = d;
    d->rank = rank;
    d->nranks = nranks;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    std::memcpy(&id, unique_id, sizeof(id));
    QPB_CUDA(cudaSetDevice(s.device));
    if (s.device < 0 || s.device >= kMaxDevices) return fail(QPB200_ERR_ARG, "qpb200_dist_create: device ordinal %d out of range", s.device);
    d->device = s.device;
    {
        // An all-zero id means "reuse this process's communicator for (device, rank, nranks)": ncclCommInitRank
        // is a ~1 s collective, far more than a solve, and a caller that re-solves related QPs should not pay it
        // per handle.  A fresh id replaces the cached communicator only when no live handle uses it; otherwise the
        // new communicator is private to this handle.  (Every rank takes the same branch as long as the ranks
        // create and destroy their handles in the same order, which the collective API requires anyway.)
        bool zero_id = true;
        for (size_t i = 0; i < sizeof(id); ++i) zero_id = zero_id && reinterpret_cast<const unsigned char *>(&id)[i] == 0;
        std::unique_lock<std::mutex> g(cache_mutex());
        CommCacheEntry &e = comm_cache(s.device);
        if (zero_id) {
            if (!e.comm || e.rank != rank || e.nranks != nranks)
                return fail(QPB200_ERR_NCCL, "qpb200_dist_create: zero NCCL id but no cached communicator for rank %d/%d on device %d", rank, nranks, s.device);
            e.users += 1;
            d->comm = e.comm;
            d->comm_cached = true;
        } else if (e.users == 0) {
            ncclComm_t old = e.comm;
            e.comm = nullptr;
            e.users = 1;                              // reserved while the (collective, blocking) init runs unlocked
            e.rank = rank; e.nranks = nranks;
            d->comm_cached = true;
            g.unlock();
            if (old) api->CommDestroy(old);
            ncclComm_t fresh = nullptr;
            const ncclResult_t r = api->CommInitRank(&fresh, nranks, id, rank);
            g.lock();
            if (r != ncclSuccess) {
                e.users = 0;
                e.rank = -1;
                d->comm_cached = false;
                return fail(QPB200_ERR_NCCL, "ncclCommInitRank failed: %s", api->GetErrorString(r));
            }
            e.comm = fresh;
            d->comm = fresh;
        } else {
            g.unlock();
            QPB_NCCL(api->CommInitRank(&d->comm, nranks, id, rank));   // owned by this handle
        }
    }
    QPB_CUDA(s.arena.alloc(&d->buf.state, 1, true));
    QPB_CUDA(s.arena.alloc(&d->buf.wbuf, (size_t)s.n + 8, true));
    QPB_CUDA(s.arena.alloc(&d->buf.wbuf2, (size_t)2 * s.n + 8, true));
    QPB_CUDA(cudaMallocHost(&d->host_state, 2 * sizeof(DistState)));
    QPB_CUDA(cudaEventCreateWithFlags(&d->ev[0], cudaEventDisableTiming));
    QPB_CUDA(cudaEventCreateWithFlags(&d->ev[1], cudaEventDisableTiming));
    for (const void *fn : {(const void *)admm_dist_kernel<1, false>, (const void *)admm_dist_kernel<1, true>}) {
        int per_sm = 0;
        if (int prc = prep_tile_kernel(fn, &per_sm)) return prc;
        if (per_sm * s.num_sms < s.grid)
            return fail(QPB200_ERR_CUDA, "segment kernel cannot be co-resident at grid %d (%d per SM)", s.grid, per_sm);
    }
    // Jacobi diagonal pieces are partial sums over the ranks' slices: combine once
    QPB_NCCL(api->AllReduce(s.prob.dP, const_cast<double *>(s.prob.dP), (size_t)s.n, ncclDouble, ncclSum, d->comm, s.stream));
    QPB_NCCL(api->AllReduce(s.prob.dAA, const_cast<double *>(s.prob.dAA), (size_t)s.n, ncclDouble, ncclSum, d->comm, s.stream));
    QPB_CUDA(cudaStreamSynchronize(s.stream));
    // collective choice (settings.reserved_i[1]): 0 = in-kernel peer-memory all-reduce when the GPUs can map each
    // other's memory, else NCCL; 1 = NCCL + host-driven segments; 2 = peer path required
    const int mode = s.settings.reserved_i[1];
    if (mode != 1) {
        const int prc = peer_init(s, *d);
        const std::string peer_msg = prc != QPB200_OK ? last_error() : std::string();
        // all ranks must agree: use the peer path only if every rank could map every peer (the agreement runs BEFORE
        // any rank returns an error, else the others would wait in the all-reduce forever)
        double ok = d->peer_ok ? 1.0 : 0.0, *dok = nullptr;
        QPB_CUDA(s.arena.alloc(&dok, 2, true));
        QPB_CUDA(cudaMemcpyAsync(dok, &ok, sizeof(double), cudaMemcpyHostToDevice, s.stream));
        QPB_NCCL(api->AllReduce(dok, dok, 1, ncclDouble, ncclMin, d->comm, s.stream));
        QPB_CUDA(cudaMemcpyAsync(&ok, dok, sizeof(double), cudaMemcpyDeviceToHost, s.stream));
        QPB_CUDA(cudaStreamSynchronize(s.stream));
        d->peer_ok = ok > 0.5;
        if (!d->peer_ok && mode == 2)
            return fail(QPB200_ERR_CUDA, "peer-memory path requested but not available on every rank%s%s", peer_msg.empty() ? "" : ": ",
                        peer_msg.c_str());
    }
    return QPB200_OK;
}


// ---- in-kernel peer-memory path --------------------------------------------------------------------
static int peer_init(SparseSolver &s, DistContext &d) {
    NcclApi *api = nccl_api();
    if (d.nranks > kMaxPeers) return QPB200_OK;      // stays on the NCCL path
    const size_t n = (size_t)s.n;
    auto up = [](size_t v) { return (v + 15) & ~size_t(15); };
    PeerDev &pd = d.peer;
    pd.rank = d.rank;
    pd.nranks = d.nranks;
    size_t off = 0;
    pd.off_flags = (long long)off; off += 16;
    pd.off_lmax = (long long)off; off += up(4 * (size_t)kMaxPeers);
    for (int q = 0; q <= d.nranks; ++q) pd.sb[q] = (int)((long long)s.n * q / d.nranks);
    for (int q = d.nranks + 1; q <= kMaxPeers; ++q) pd.sb[q] = s.n;
    pd.sstride = (long long)up(n / (size_t)d.nranks + 1);
    pd.off_wrecv = (long long)off; off += (size_t)pd.sstride * (size_t)d.nranks;
    pd.off_w2part = (long long)off; off += up(2 * n);
    pd.off_w2red = (long long)off; off += up(2 * n);
    // sliced variant: per-CTA dot partials of every rank, and the gathered vector pairs [u ; t], [x~ ; g]
    // (their n-parts are written by the peers) -- same offsets on every rank, hence sized with max_r m_r
    pd.grid_max = s.num_sms * kMinCtas;
    pd.off_cta = (long long)off; off += up((size_t)2 * ((size_t)pd.grid_max * 4 + kMaxPeers * 4));
    double mmax = (double)s.m, *dm = nullptr;
    QPB_CUDA(s.arena.alloc(&dm, 2, true));
    QPB_CUDA(cudaMemcpyAsync(dm, &mmax, sizeof(double), cudaMemcpyHostToDevice, s.stream));
    QPB_NCCL(api->AllReduce(dm, dm, 1, ncclDouble, ncclMax, d.comm, s.stream));
    QPB_CUDA(cudaMemcpyAsync(&mmax, dm, sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    QPB_CUDA(cudaStreamSynchronize(s.stream));
    const size_t pair = up(n + (size_t)mmax + 8);
    const size_t off_UT = off; off += pair;
    const size_t off_XG = off; off += pair;
    d.region_doubles = off;
    // ---- collective decision: reuse the process-cached region + mappings only if EVERY rank can
    // (the entry of this device is only touched by the one rank that drives the device; the lock covers the flags)
    PeerRegionCache &rc = region_cache(s.device);
    bool mine_ok;
    {
        std::lock_guard<std::mutex> g(cache_mutex());
        mine_ok = !rc.in_use && rc.region && rc.rank == d.rank && rc.nranks == d.nranks && rc.doubles >= off;
        if (mine_ok) rc.in_use = true;               // reserved; released below if the ranks do not all agree
    }
    double cannot = mine_ok ? 0.0 : 1.0;
    QPB_CUDA(cudaMemcpyAsync(dm, &cannot, sizeof(double), cudaMemcpyHostToDevice, s.stream));
    QPB_NCCL(api->AllReduce(dm, dm, 1, ncclDouble, ncclMax, d.comm, s.stream));
    QPB_CUDA(cudaMemcpyAsync(&cannot, dm, sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    QPB_CUDA(cudaStreamSynchronize(s.stream));
    if (cannot >= 0.5 && mine_ok) {
        std::lock_guard<std::mutex> g(cache_mutex());
        rc.in_use = false;
    }
    if (cannot < 0.5) {
        d.region_cached = true;
        d.region = rc.region;
        QPB_CUDA(cudaMemsetAsync(d.region, 0, off * sizeof(double), s.stream));
        for (int q = 0; q < d.nranks; ++q) pd.region[q] = q == d.rank ? d.region : static_cast<double *>(rc.opened[q]);
    } else {
        bool own_cache;                          // nobody holds the entry: replace it; else allocate privately
        {
            std::lock_guard<std::mutex> g(cache_mutex());
            own_cache = !rc.in_use;
            if (own_cache) rc.in_use = true;
        }
        if (own_cache) {
            for (void *&o : rc.opened) {
                if (o) cudaIpcCloseMemHandle(o);
                o = nullptr;
            }
        }
        // every rank has dropped its mappings of the old regions before anybody frees one
        QPB_NCCL(api->AllReduce(dm, dm, 1, ncclDouble, ncclMax, d.comm, s.stream));
        QPB_CUDA(cudaStreamSynchronize(s.stream));
        if (own_cache && rc.region) {
            cudaFree(rc.region);
            rc.region = nullptr;
            rc.doubles = 0;
        }
        QPB_CUDA(cudaMalloc(&d.region, off * sizeof(double)));
        QPB_CUDA(cudaMemset(d.region, 0, off * sizeof(double)));
        // exchange the IPC handles through the NCCL communicator we already have
        cudaIpcMemHandle_t mine;
        QPB_CUDA(cudaIpcGetMemHandle(&mine, d.region));
        unsigned char *hbuf = nullptr;
        const size_t hsz = sizeof(cudaIpcMemHandle_t);
        QPB_CUDA(s.arena.alloc(&hbuf, hsz * d.nranks, true));
        QPB_CUDA(cudaMemcpyAsync(hbuf + hsz * d.rank, &mine, hsz, cudaMemcpyHostToDevice, s.stream));
        QPB_NCCL(api->AllGather(hbuf + hsz * d.rank, hbuf, hsz, ncclChar, d.comm, s.stream));
        std::vector<cudaIpcMemHandle_t> all((size_t)d.nranks);
        QPB_CUDA(cudaMemcpyAsync(all.data(), hbuf, hsz * d.nranks, cudaMemcpyDeviceToHost, s.stream));
        QPB_CUDA(cudaStreamSynchronize(s.stream));
        for (int q = 0; q < d.nranks; ++q) {
            if (q == d.rank) {
                pd.region[q] = d.region;
                continue;
            }
            void *ptr = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&ptr, all[(size_t)q], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                if (own_cache) {                 // the entry was reserved above: release it empty; dist_destroy frees ours
                    std::lock_guard<std::mutex> g(cache_mutex());
                    rc.in_use = false;
                    rc.region = nullptr;
                    rc.doubles = 0;
                }
                return fail(QPB200_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s -- peer path unavailable", q, cudaGetErrorString(e));
            }
            d.opened[q] = ptr;
            pd.region[q] = static_cast<double *>(ptr);
        }
        if (own_cache) {                         // hand the new region to the cache; this handle borrows it
            rc.rank = d.rank; rc.nranks = d.nranks;
            rc.region = d.region; rc.doubles = off;
            for (int q = 0; q < kMaxPeers; ++q) rc.opened[q] = d.opened[q];
            d.region_cached = true;
        }
    }
    // the gathered pairs move into the peer-visible region (the arena copies stay unused)
    s.prob.UT = d.region + off_UT;
    s.prob.XG = d.region + off_XG;
    pd.info = s.prob.info;
    pd.dbg = nullptr;
    if (getenv("QPB200_TIMING")) QPB_CUDA(s.arena.alloc(&pd.dbg, 16, true));
    QPB_CUDA(s.arena.alloc(&d.tiny, 2, true));
    for (const void *fn : {(const void *)admm_peer_sliced_kernel<1, false>, (const void *)admm_peer_sliced_kernel<1, true>}) {
        int per_sm = 0;
        if (int prc = prep_tile_kernel(fn, &per_sm)) return prc;
        if (per_sm * s.num_sms < s.grid)
            return fail(QPB200_ERR_CUDA, "peer kernel cannot be co-resident at grid %d (%d per SM)", s.grid, per_sm);
    }
    d.peer_ok = true;
    return QPB200_OK;
}

int peer_solve(SparseSolver &s, DistContext &d, double *x_inout, double *z_out, double *y_out, qpb200_info *info) {
    NcclApi *api = nccl_api();
    if (!x_inout) return fail(QPB200_ERR_ARG, "qpb200_dist_solve: x_inout is NULL");
    QPB_CUDA(cudaSetDevice(s.device));
    const int n = s.n, m = s.m;
    int rc = s.reset_state(x_inout);
    if (rc) return rc;
    // epoch flags back to zero, then a host-level rendezvous: nobody launches before everybody has reset
    QPB_CUDA(cudaMemsetAsync(d.region, 0, (size_t)(d.peer.off_wrecv) * sizeof(double), s.stream));
    QPB_NCCL(api->AllReduce(d.tiny, d.tiny, 1, ncclDouble, ncclSum, d.comm, s.stream));
    QPB_CUDA(cudaStreamSynchronize(s.stream));
    QPB_CUDA(cudaEventRecord(s.ev0, s.stream));
    {
        void *args[] = {(void *)&s.prob, (void *)&d.peer};
        const void *fn = s.use_pre ? (const void *)admm_peer_sliced_kernel<1, true> : (const void *)admm_peer_sliced_kernel<1, false>;
        QPB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(s.grid), dim3(kThreads), args, sizeof(SpmvSmem), s.stream));
    }
    QPB_CUDA(cudaEventRecord(s.ev1, s.stream));
    QPB_CUDA(cudaMemcpyAsync(x_inout, s.prob.XY, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    if (z_out && m) QPB_CUDA(cudaMemcpyAsync(z_out, s.prob.z, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    if (y_out && m) QPB_CUDA(cudaMemcpyAsync(y_out, s.prob.XY + n, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    AdmmInfoDev hi;
    QPB_CUDA(cudaMemcpyAsync(&hi, s.prob.info, sizeof(hi), cudaMemcpyDeviceToHost, s.stream));
    QPB_CUDA(cudaStreamSynchronize(s.stream));
    // nobody tears its buffers down (or starts the next solve) while a peer may still be reading them
    QPB_NCCL(api->AllReduce(d.tiny, d.tiny, 1, ncclDouble, ncclSum, d.comm, s.stream));
    QPB_CUDA(cudaStreamSynchronize(s.stream));
    float ms = 0.f;
    QPB_CUDA(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
    s.last_info = hi;
    if (d.peer.dbg) {
        unsigned long long t[16];
        QPB_CUDA(cudaMemcpy(t, d.peer.dbg, sizeof(t), cudaMemcpyDeviceToHost));
        QPB_CUDA(cudaMemset(d.peer.dbg, 0, sizeof(t)));
        const double k = hi.pcg_iters_total > 0 ? 1e-3 / (double)hi.pcg_iters_total : 0.0;
        fprintf(stderr, "[qpb200 peer rank %d] us per CG iteration: A pass + grid barrier %.1f, H pass with push %.1f, system barrier + z.w sum %.1f, "
                        "slice update + z push %.1f, system barrier + 3 sums %.1f, loop head %.1f (solve %.1f ms, %lld CG its)\n",
                d.rank, t[0] * k, t[1] * k, t[2] * k, t[3] * k, t[4] * k, t[5] * k, ms, (long long)hi.pcg_iters_total);
    }
    if (info) {
        std::memset(info, 0, sizeof(*info));
        info->conv_flag = hi.conv_flag;
        info->iterations = hi.iterations;
        info->rho_final = hi.rho_final;
        info->res_prim = hi.res_prim;
        info->res_dual = hi.res_dual;
        info->rho_updates = hi.rho_updates;
        info->pcg_iters_total = hi.pcg_iters_total;
        info->pcg_maxed = hi.pcg_maxed;
        info->solve_ms = ms;
        info->setup_ms = s.setup_ms;
        info->kernel_launches = 1;
    }
    return QPB200_OK;
}

}  // namespace qpb

using namespace qpb;

extern "C" {

int qpb200_dist_unique_id(void *id128) {
    NcclApi *api = nccl_api();
    if (!api) return fail(QPB200_ERR_NCCL, "libnccl.so.2 could not be loaded (dlopen)");
    if (!id128) return fail(QPB200_ERR_ARG, "qpb200_dist_unique_id: NULL");
    ncclUniqueId id;
    QPB_NCCL(api->GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof(id));
    return QPB200_OK;
}

int qpb200_dist_create(qpb200_handle **out, int32_t rank, int32_t nranks, const void *nccl_unique_id, int64_t n,
                       int64_t m_local, const int64_t *P_colptr, const int64_t *P_rowval, const double *P_nzval,
                       const int64_t *A_colptr, const int64_t *A_rowval, const double *A_nzval, const double *q,
                       const double *l_local, const double *u_local, const qpb200_settings *settings, int32_t index_base) {
    if (!out) return fail(QPB200_ERR_ARG, "qpb200_dist_create: out is NULL");
    *out = nullptr;
    qpb200_settings s;
    if (settings) s = *settings;
    else qpb200_default_settings(&s);
    set_host_thread_share(nranks);                 // the ranks of one node share its cores
    if (s.lin_solver != QPB200_LINSOLVE_PCG || s.reserved_i[QPB200_RSV_POLISH] != 0)
        return fail(QPB200_ERR_ARG, "qpb200_dist_create: the row-partitioned path implements lin_solver = PCG without polish "
                                    "(the exact solve and the polish need the whole K / KKT operator on one device)");
    if (s.reserved_i[QPB200_RSV_SCALING_ITERS] != 0)
        return fail(QPB200_ERR_ARG, "qpb200_dist_create: equilibration needs column norms over all ranks' rows; not implemented for the row-partitioned path");
    qpb200_handle *h = new (std::nothrow) qpb200_handle();
    if (!h) return fail(QPB200_ERR_ARG, "out of host memory");
    int rc = h->solver.init(n, m_local, P_colptr, P_rowval, P_nzval, A_colptr, A_rowval, A_nzval, q, l_local, u_local, s, index_base);
    if (rc == QPB200_OK) rc = dist_init(h->solver, h->dist, rank, nranks, nccl_unique_id);
    if (rc == QPB200_OK) h->dist->row_begin = h->dist->row_end = -1;   // caller-made slice (qpb200_dist_create_full fills them in)
    if (rc != QPB200_OK) {
        qpb200_destroy(h);
        return rc;
    }
    *out = h;
    return QPB200_OK;
}

int qpb200_dist_create_full(qpb200_handle **out, int32_t rank, int32_t nranks, const void *nccl_unique_id, int64_t n,
                            int64_t m, const int64_t *P_colptr, const int64_t *P_rowval, const double *P_nzval,
                            const int64_t *A_colptr, const int64_t *A_rowval, const double *A_nzval, const double *q,
                            const double *l, const double *u, const qpb200_settings *settings, int32_t index_base) {
    if (!out) return fail(QPB200_ERR_ARG, "qpb200_dist_create_full: out is NULL");
    *out = nullptr;
    if (n <= 0 || m < 0 || nranks < 1 || rank < 0 || rank >= nranks || (index_base != 0 && index_base != 1))
        return fail(QPB200_ERR_ARG, "qpb200_dist_create_full: bad n / m / rank / nranks / index_base");
    if (!q || (m > 0 && (!l || !u))) return fail(QPB200_ERR_ARG, "q, l, u must not be NULL");
    set_host_thread_share(nranks);                 // the ranks of one node share its cores
    int rc = validate_csc("P", n, n, P_colptr, P_rowval, P_nzval, index_base);
    if (rc) return rc;
    if ((rc = validate_csc("A", m, n, A_colptr, A_rowval, A_nzval, index_base))) return rc;
    std::vector<int64_t> rows, cols;
    dist_plan(n, m, P_colptr, A_colptr, A_rowval, index_base, nranks, rows, cols);
    DistSlice sl;
    dist_slice(n, m, P_colptr, A_colptr, A_rowval, A_nzval, index_base, rank, rows, cols, sl);
    rc = qpb200_dist_create(out, rank, nranks, nccl_unique_id, n, sl.i1 - sl.i0, sl.Pcolptr.data(), P_rowval + sl.p_off,
                            P_nzval + sl.p_off, sl.Acolptr.data(), sl.Arowval.data(), sl.Anzval.data(), q,
                            m ? l + sl.i0 : l, m ? u + sl.i0 : u, settings, index_base);
    if (rc == QPB200_OK) {
        (*out)->dist->row_begin = sl.i0;
        (*out)->dist->row_end = sl.i1;
    }
    return rc;
}

int qpb200_dist_rows(qpb200_handle *h, int64_t *row_begin, int64_t *row_end) {
    if (!h || !h->dist) return fail(QPB200_ERR_ARG, "qpb200_dist_rows: not a distributed handle");
    if (h->dist->row_end < 0) return fail(QPB200_ERR_ARG, "qpb200_dist_rows: the handle was created from a caller-made slice");
    if (row_begin) *row_begin = h->dist->row_begin;
    if (row_end) *row_end = h->dist->row_end;
    return QPB200_OK;
}

int64_t qpb200_debug_partition(int64_t n, int64_t m, const int64_t *P_colptr, const int64_t *A_colptr, const int64_t *A_rowval,
                               int32_t index_base, int32_t nranks, int64_t *row_bounds_out, int64_t *col_bounds_out) {
    if (n <= 0 || m < 0 || nranks < 1 || !P_colptr || !A_colptr) return fail(QPB200_ERR_ARG, "qpb200_debug_partition: bad argument");
    std::vector<int64_t> rows, cols;
    dist_plan(n, m, P_colptr, A_colptr, A_rowval, index_base, nranks, rows, cols);
    if (row_bounds_out) std::copy(rows.begin(), rows.end(), row_bounds_out);
    if (col_bounds_out) std::copy(cols.begin(), cols.end(), col_bounds_out);
    return nranks;
}

int qpb200_debug_slice(int64_t n, int64_t m, const int64_t *P_colptr, const int64_t *A_colptr, const int64_t *A_rowval,
                       const double *A_nzval, int32_t index_base, int32_t rank, int32_t nranks, int64_t *bounds4_out,
                       int64_t *p_off_out, int64_t *Pcolptr_out, int64_t *Acolptr_out, int64_t *Arowval_out, double *Anzval_out,
                       int64_t a_cap) {
    if (n <= 0 || m < 0 || nranks < 1 || rank < 0 || rank >= nranks) return fail(QPB200_ERR_ARG, "qpb200_debug_slice: bad argument");
    std::vector<int64_t> rows, cols;
    dist_plan(n, m, P_colptr, A_colptr, A_rowval, index_base, nranks, rows, cols);
    DistSlice sl;
    dist_slice(n, m, P_colptr, A_colptr, A_rowval, A_nzval, index_base, rank, rows, cols, sl);
    if (bounds4_out) { bounds4_out[0] = sl.i0; bounds4_out[1] = sl.i1; bounds4_out[2] = sl.j0; bounds4_out[3] = sl.j1; }
    if (p_off_out) *p_off_out = sl.p_off;
    if (Pcolptr_out) std::copy(sl.Pcolptr.begin(), sl.Pcolptr.end(), Pcolptr_out);
    if (Acolptr_out) std::copy(sl.Acolptr.begin(), sl.Acolptr.end(), Acolptr_out);
    const int64_t nnz = sl.Acolptr[(size_t)n] - index_base;
    if (nnz <= a_cap) {
        if (Arowval_out) std::copy(sl.Arowval.begin(), sl.Arowval.begin() + nnz, Arowval_out);
        if (Anzval_out) std::copy(sl.Anzval.begin(), sl.Anzval.begin() + nnz, Anzval_out);
    }
    return QPB200_OK;
}

int qpb200_dist_solve(qpb200_handle *h, double *x_inout, double *z_out, double *y_out, qpb200_info *info) {
    if (!h || !h->dist) return fail(QPB200_ERR_ARG, "qpb200_dist_solve: not a distributed handle");
    // (a start point that differs between the ranks, or is non-finite on some only, would desynchronise the collective:
    //  x_inout is documented as replicated, so every rank takes the same branch here)
    if (x_inout && !all_finite(x_inout, (size_t)h->solver.n))
        return fail(QPB200_ERR_NONFINITE, "qpb200_dist_solve: the start point holds NaN or Inf");
    if (h->dist->peer_ok) return peer_solve(h->solver, *h->dist, x_inout, z_out, y_out, info);
    return dist_solve(h->solver, *h->dist, x_inout, z_out, y_out, info);
}

}  // extern "C"
