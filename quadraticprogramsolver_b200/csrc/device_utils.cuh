// device_utils.cuh -- PTX wrappers (mbarrier / 1-D TMA bulk copy), the grid-wide barrier with
// fused deterministic reductions, and block-level reductions.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qpb {

#ifndef QPB_THREADS
#define QPB_THREADS 256
#endif
#ifndef QPB_MIN_CTAS
#define QPB_MIN_CTAS 4
#endif
constexpr int kThreads = QPB_THREADS;   // threads per CTA of every sparse-path kernel (build-time tunable)
constexpr int kMinCtas = QPB_MIN_CTAS;  // __launch_bounds__ minimum resident CTAs per SM
constexpr int kWarps = kThreads / 32;
constexpr int kMaxRed = 8;      // widest fused reduction (values per barrier)

// ------------------------------------------------------------------------------------------
// mbarrier + cp.async.bulk (1-D TMA).  SASS: SYNCS.* / UBLKCP.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// global -> shared bulk copy; dst, src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// order this thread's generic-proxy shared-memory accesses before later async-proxy (TMA) ones
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// release/acquire fence at gpu scope (lighter than the sequentially-consistent __threadfence())
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ------------------------------------------------------------------------------------------
// Grid-wide barrier for the persistent (cooperatively launched) kernels.
//   count : monotonically increasing arrival counter (zeroed by the host before the launch)
//   flag  : epoch published by the last arriver; the others poll it (a different L2 line than
//           the one the atomics hit)
//   partials[2] : ping-pong [gridDim.x * kMaxRed] slots for the fused reductions
// ------------------------------------------------------------------------------------------
struct GridSync {
    unsigned long long *count;
    unsigned long long *flag;
    double *partials[2];
};

struct SyncState {          // per-thread (uniform) barrier bookkeeping
    unsigned long long epoch;
};

__device__ __forceinline__ void grid_barrier(const GridSync &gs, SyncState &st) {
    st.epoch += 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        // release: this CTA's writes (ordered before us by the bar.sync above) become visible before the arrival
        fence_acq_rel_gpu();
        const unsigned long long old = atomicAdd(gs.count, 1ULL);
        if (old == st.epoch * gridDim.x - 1ULL) {
            // last arriver: it has observed every other arrival through the RMW chain; the fence makes that
            // cumulative before the flag is published (only this one CTA pays for it)
            fence_acq_rel_gpu();
            st_release_gpu(gs.flag, st.epoch);
        } else {
            // a CTA that never arrives (a launch that is not co-resident, counters left over from another launch) must
            // fail loudly, not hang the GPU: trap after ~20 s of polling
            long long t0 = 0;
            unsigned spins = 0;
            while (ld_acquire_gpu(gs.flag) < st.epoch) {
                if ((++spins & 0xfffffu) == 0) {
                    if (t0 == 0) t0 = clock64();
                    else if (clock64() - t0 > 40000000000LL) __trap();
                }
            }
        }
        // acquire + L1 invalidate (SASS: MEMBAR ; CCTL.IVALL): the phase that follows reads vectors other CTAs
        // wrote with plain cached loads
        fence_acq_rel_gpu();
    }
    __syncthreads();
}

// NaN-propagating max of absolute values (norm(., Inf) semantics)
__device__ __forceinline__ double nanmax(double s, double v) { return (v > s || v != v) ? v : s; }

// Block-level reduction of NV per-thread values (sum or max); result in thread 0's v[].
// red: shared scratch of kWarps * NV doubles.
template <int NV, bool IS_MAX>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double y = __shfl_down_sync(0xffffffffu, x, o);
            x = IS_MAX ? nanmax(x, y) : x + y;
        }
        if (lane == 0) red[warp * NV + i] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double x = (lane < kWarps) ? red[lane * NV + i] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double y = __shfl_down_sync(0xffffffffu, x, o);
                x = IS_MAX ? nanmax(x, y) : x + y;
            }
            v[i] = x;
        }
    }
    __syncthreads();
}

// Barrier + all-reduce: every CTA ends up with bit-identical totals in v[] (all threads), because
// every CTA combines the per-CTA partials in the same fixed order.  Deterministic run to run.
template <int NV, bool IS_MAX>
__device__ __forceinline__ void grid_barrier_reduce(const GridSync &gs, SyncState &st, double (&v)[NV], double *red,
                                                    double *bcast) {
    static_assert(NV <= kMaxRed, "too many fused reduction values");
    block_reduce<NV, IS_MAX>(v, red);
    double *slots = gs.partials[(st.epoch + 1) & 1];
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) slots[blockIdx.x * kMaxRed + i] = v[i];
    }
    grid_barrier(gs, st);
    // Every CTA combines the gridDim.x partials in the same fixed order.  All threads of the CTA take part: thread t
    // loads partials t, t + kThreads, ... (independent L2 loads, all in flight together), then one block reduction.
    // (Round 1 had a single warp walk the partials 32 at a time: ~19 dependent L2 round trips per reduction at 592
    //  CTAs, 4-6 us on the critical path of every dot product.)
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0;
    for (unsigned j = threadIdx.x; j < gridDim.x; j += kThreads) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const double y = __ldcg(slots + j * kMaxRed + i);
            acc[i] = IS_MAX ? nanmax(acc[i], y) : acc[i] + y;
        }
    }
    block_reduce<NV, IS_MAX>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) bcast[i] = acc[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = bcast[i];
    __syncthreads();   // bcast may be rewritten by the next reduction
}

}  // namespace qpb
