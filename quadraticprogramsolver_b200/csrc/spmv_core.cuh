// spmv_core.cuh -- tiled CSR SpMV ("CSR-stream") device core, used by the stand-alone SpMV kernel
// and by every matrix phase of the persistent ADMM kernels.
//
// Replaces SparseArrays/MKLSparse `mul!` (reference call sites LinearSystemSolvers.jl:135,139,
// 153-155 and SolveQuadraticProgram.jl:85-89).
//
// Layout (HBM): CSR with int32 row pointers / column indices and FP64 values.  The row range is cut
// on the host into *tiles*: maximal runs of consecutive rows holding <= kTileFill non-zeros and
// <= kTileRows rows (a row longer than that is split into segments that stay on one CTA).  Each CTA
// owns a contiguous, nnz-balanced run of tiles (static => bitwise reproducible).  Per tile:
//   1. the tile's (col, val) stream AND its row pointers arrive in shared memory by 1-D TMA bulk
//      copies (cp.async.bulk, mbarrier complete-tx, kStages deep) -- no global load of matrix data
//      is ever issued by a thread, so nothing queues behind the gathers,
//   2. every thread takes pairs of slots laid out so that the warp's 64-bit column loads and 128-bit
//      value loads / product stores are contiguous (conflict free), 4 independent gathers of x through
//      L1/L2 in flight per thread,
//   3. `lpr` lanes per row sum the row's products out of shared memory and hand the row sum to the
//      phase's epilogue functor (which fuses the vector update / norm / dot product of that phase).
// Round 2: step 2 used to be 4 strided scalar slots per thread with per-slot bounds predicates and
// step 3 fetched row pointers from global memory and divided by the run-time `lpr`; ncu counted 2.4
// warp instructions per non-zero (61 % issue utilisation at 0.55 of the HBM peak on a banded
// matrix whose gathers hit L1: the engine was instruction bound).  See DESIGN.md 4.1.
// Algorithmic bytes per launch: 12*nnz + 4*(rows+1) + 8*cols + 8*rows.
#pragma once
#include "device_utils.cuh"

namespace qpb {

#ifndef QPB_TILE_NNZ
#define QPB_TILE_NNZ 1024
#endif
#ifndef QPB_STAGES
#define QPB_STAGES 3
#endif
constexpr int kTileNnz = QPB_TILE_NNZ;    // (col, val) slots per stage (build-time tunable for A/B runs)
constexpr int kTileFill = kTileNnz - 8;   // most non-zeros the host packs into a tile: the staged range starts at a
                                          // 16-byte boundary (<= 3 slots early) and is rounded up to 4 slots
constexpr int kTileRows = kThreads;       // most rows per tile (one row-sum round when lpr = 1)
constexpr int kRowSlots = kTileRows + 8;  // staged row pointers: rows + 1, same alignment slack
constexpr int kStages = QPB_STAGES;       // TMA pipeline depth
constexpr int kSlotGroups = kTileNnz / (4 * kThreads);   // groups of 4 slots per thread and tile
static_assert(kTileNnz == kSlotGroups * 4 * kThreads, "a tile must be a whole number of 4-slot groups per thread");

// tile descriptor: x = first row, y = #rows, z = first nnz (k0), w = #nnz | flags
constexpr int kTileContFromPrev = 1 << 30;   // this tile continues a long row started earlier
constexpr int kTileContToNext = 1 << 29;     // the long row continues into the next tile
constexpr int kTileNkMask = (1 << 24) - 1;

struct CsrTiled {
    int rows, cols;
    const int *rowptr;     // rows + 1 (+ 8 readable padding entries)
    const int *rowmid;     // rows, optional: first nnz of the second column block (split sums)
    const int *col;        // nnz (+ padding)
    const double *val;     // nnz (+ padding)
    const int4 *tiles;     // ntiles
    const int *cta_begin;  // grid + 1: tile range of every CTA
    int ntiles;
    int lpr;               // lanes per row in the row-sum step (power of two, <= 32)
};

// Shared memory of one CTA (40.6 KB with the defaults -> 4 CTAs per SM leave ~90 KB of L1 to the gathers).
struct __align__(128) SpmvSmem {
    double val[kStages][kTileNnz];   // staged values, overwritten in place by the products
    int col[kStages][kTileNnz];
    int rp[kStages][kRowSlots];      // staged row pointers of the tile
    int4 tdq[kStages];               // descriptor of the tile in each stage (written by the issuing thread)
    uint64_t full[kStages];          // mbarriers
    double red[kWarps * kMaxRed];
    double bcast[kMaxRed];
    double carry[2];                 // running sums of a long row
};

struct PipeState {     // uniform across the CTA
    uint32_t parity;   // bit s = parity to wait for on stage s
};

__device__ __forceinline__ void tma_issue_tile(const CsrTiled &M, const int4 td, SpmvSmem &sm, int stage) {
    const int k0 = td.z, nk = td.w & kTileNkMask;
    const int k0a = k0 & ~3;                              // 16-byte aligned start for all three arrays
    const int cnt = ((k0 + nk - k0a) + 3) & ~3;           // <= kTileFill + 6 <= kTileNnz
    const int r0a = td.x & ~3;
    const int rcnt = ((td.x + td.y + 1 - r0a) + 3) & ~3;  // <= kTileRows + 7 <= kRowSlots
    sm.tdq[stage] = td;                                   // released to the consumers by the mbarrier arrival below
    mbar_expect_tx(&sm.full[stage], static_cast<uint32_t>(cnt) * 12u + static_cast<uint32_t>(rcnt) * 4u);
    if (cnt) {
        tma_load_1d(sm.val[stage], M.val + k0a, static_cast<uint32_t>(cnt) * 8u, &sm.full[stage]);
        tma_load_1d(sm.col[stage], M.col + k0a, static_cast<uint32_t>(cnt) * 4u, &sm.full[stage]);
    }
    tma_load_1d(sm.rp[stage], M.rowptr + r0a, static_cast<uint32_t>(rcnt) * 4u, &sm.full[stage]);
}

// ---- step 2: products val[k] * x[col[k]] in place -------------------------------------------------------
// Every thread takes kSlotGroups x 2 PAIRS of slots; pair h of thread t sits at slot 2 (h kThreads + t), so a warp's
// 64-bit column loads and 128-bit value loads / product stores are contiguous: conflict free (a 128-bit access is
// served a quarter warp at a time; 4 consecutive slots per thread would put those 8 lanes 32 bytes apart = 2-way
// bank conflicts, measured as 2.3 of 5.1 M store wavefronts per H pass).  All gathers of x are issued before the
// first use.  x is mutable between phases: plain loads, never ld.global.nc.  Slots past the staged range keep stale
// data and are not touched: every slot that is multiplied holds a column index of this matrix.
constexpr int kPairs = 2 * kSlotGroups;
struct Gathered {          // the x values of one tile's slots owned by this thread, in flight or landed
    double xa[kPairs], xb[kPairs];
};

__device__ __forceinline__ void tile_gather(const int *col, int cnt, const double *x, Gathered &g) {
    int2 c[kPairs];
#pragma unroll
    for (int h = 0; h < kPairs; ++h) {
        const int s2 = (h * kThreads + threadIdx.x) * 2;
        c[h] = s2 < cnt ? *reinterpret_cast<const int2 *>(col + s2) : make_int2(0, 0);
    }
#pragma unroll
    for (int h = 0; h < kPairs; ++h) {
        const int s2 = (h * kThreads + threadIdx.x) * 2;
        g.xa[h] = g.xb[h] = 0.0;   // (defined on every path: otherwise the values are loop carried and get spilled)
        if (s2 < cnt) {
            g.xa[h] = x[static_cast<unsigned>(c[h].x)];
            g.xb[h] = x[static_cast<unsigned>(c[h].y)];
        }
    }
}

__device__ __forceinline__ void tile_multiply(double *val, int cnt, const Gathered &g) {
#pragma unroll
    for (int h = 0; h < kPairs; ++h) {
        const int s2 = (h * kThreads + threadIdx.x) * 2;
        if (s2 < cnt) {
            double2 v = *reinterpret_cast<const double2 *>(val + s2);
            v.x *= g.xa[h];
            v.y *= g.xb[h];
            *reinterpret_cast<double2 *>(val + s2) = v;
        }
    }
}

// sum of prod[a + gl], prod[a + gl + lpr], ... below b: two independent chains (the loop is latency bound)
__device__ __forceinline__ double strided_sum(const double *prod, int a, int b, int lpr) {
    double s0 = 0.0, s1 = 0.0;
    const double *p = prod + a, *e = prod + b;
    for (; p + lpr < e; p += 2 * lpr) {
        s0 += p[0];
        s1 += p[lpr];
    }
    if (p < e) s0 += p[0];
    return s0 + s1;
}

__device__ __forceinline__ double group_reduce(double s, int lpr) {   // lane 0 of every aligned group of lpr lanes
    if (lpr > 16) s += __shfl_down_sync(0xffffffffu, s, 16);
    if (lpr > 8) s += __shfl_down_sync(0xffffffffu, s, 8);
    if (lpr > 4) s += __shfl_down_sync(0xffffffffu, s, 4);
    if (lpr > 2) s += __shfl_down_sync(0xffffffffu, s, 2);
    if (lpr > 1) s += __shfl_down_sync(0xffffffffu, s, 1);
    return s;
}

// ---- step 3: row sums ------------------------------------------------------------------------------
// prod: products of this tile, non-zero k of the matrix at prod[k - k0a]; rp: staged row pointers, row r of the
// tile at rp[r], rp[r + 1].
template <bool SPLIT, class Epi>
__device__ __forceinline__ void tile_row_sums(const CsrTiled &M, const int4 td, const double *prod, const int *rp,
                                              int k0a, int lsh, SpmvSmem &sm, Epi &epi) {
    const int row0 = td.x, nrows = td.y;
    if (td.w & (kTileContFromPrev | kTileContToNext)) {
        // one segment of a long row: whole-CTA sum, carried across the row's tiles
        const bool from_prev = td.w & kTileContFromPrev, to_next = td.w & kTileContToNext;
        const int a = td.z - k0a, b = a + (td.w & kTileNkMask);
        double s[2] = {0.0, 0.0};
        const int mid = SPLIT ? (M.rowmid[row0] - k0a) : b;
        for (int k = a + threadIdx.x; k < b; k += kThreads) {
            if (!SPLIT || k < mid) s[0] += prod[k];
            else s[1] += prod[k];
        }
        block_reduce<2, false>(s, sm.red);
        if (threadIdx.x == 0) {
            const double c0 = (from_prev ? sm.carry[0] : 0.0) + s[0];
            const double c1 = (from_prev ? sm.carry[1] : 0.0) + s[1];
            if (to_next) { sm.carry[0] = c0; sm.carry[1] = c1; }
            else epi(row0, c0, c1);
        }
        return;
    }
    const int lpr = 1 << lsh;
    const int g = threadIdx.x >> lsh, gl = threadIdx.x & (lpr - 1);
    const int groups = kThreads >> lsh;
    for (int rb = 0; rb < nrows; rb += groups) {
        const int r = rb + g;
        const bool live = r < nrows;
        double s0 = 0.0, s1 = 0.0;
        if (live) {
            const int a = rp[r] - k0a, b = rp[r + 1] - k0a;
            if (!SPLIT) {
                s0 = strided_sum(prod, a + gl, b, lpr);
            } else {
                const int mid = __ldg(M.rowmid + row0 + r) - k0a;
                s0 = strided_sum(prod, a + gl, mid, lpr);
                s1 = strided_sum(prod, mid + gl, b, lpr);
            }
        }
        s0 = group_reduce(s0, lpr);
        if (SPLIT) s1 = group_reduce(s1, lpr);
        if (live && gl == 0) epi(row0 + r, s0, s1);
    }
}

// ---- the tile loop of one matrix phase ----------------------------------------------------------------
// LOADER is kept in the signature for the callers' template lists; there is one loader (TMA bulk copies).
template <int LOADER, bool SPLIT, class Epi>
__device__ __forceinline__ void spmv_tiles(const CsrTiled &M, const double *x, SpmvSmem &sm, PipeState &ps, Epi &epi) {
    const int tb = M.cta_begin[blockIdx.x], te = M.cta_begin[blockIdx.x + 1];
    const int nt = te - tb;
    const int lsh = 31 - __clz(M.lpr);
    // the stages were last touched through the generic proxy (previous phase's row sums)
    fence_proxy_async_smem();
    __syncthreads();
    constexpr int kAhead = kStages - 1;     // tiles in flight ahead of the one being processed
    if (threadIdx.x == 0) {
        const int pre = nt < kAhead ? nt : kAhead;
        for (int i = 0; i < pre; ++i) tma_issue_tile(M, __ldg(M.tiles + tb + i), sm, i % kStages);
    }
    // The issuing thread fetches tile descriptors ONE ITERATION AHEAD of the refill that needs them (a descriptor
    // load issued when it is needed queues behind the thousands of x-gathers the SM has in flight and serialises
    // the tile loop) and hands them to the other threads through shared memory next to the tile itself.
    int4 td_issue = (threadIdx.x == 0 && kAhead < nt) ? __ldg(M.tiles + tb + kAhead) : make_int4(0, 0, 0, 0);
    int s = 0;
    for (int i = 0; i < nt; ++i) {
        mbar_wait(&sm.full[s], (ps.parity >> s) & 1u);
        ps.parity ^= (1u << s);
        const int4 td = sm.tdq[s];
        const int k0a = td.z & ~3;
        const int cnt = ((td.z + (td.w & kTileNkMask) - k0a) + 3) & ~3;
        Gathered g;
        tile_gather(sm.col[s], cnt, x, g);
        tile_multiply(sm.val[s], cnt, g);
        fence_proxy_async_smem();   // our generic accesses to the stages (incl. tile i-1's reads) before the refill
        __syncthreads();
        if (threadIdx.x == 0 && i + kAhead < nt) {
            tma_issue_tile(M, td_issue, sm, s == 0 ? kStages - 1 : s - 1);
            if (i + 1 + kAhead < nt) td_issue = __ldg(M.tiles + tb + i + 1 + kAhead);
        }
        tile_row_sums<SPLIT>(M, td, sm.val[s], sm.rp[s] + (td.x & 3), k0a, lsh, sm, epi);
        s = (s + 1 == kStages) ? 0 : s + 1;
    }
    // (Issuing the gathers of tile i+1 before the row sums of tile i was measured twice, round 1 and round 2
    //  (profiles/r2e_spmv_variants.jsonl: cfg5 H pass 0.176 -> 0.201 ms): the L2 -> SM return path is already at
    //  ~90 % of the best rate measured on this GPU, extra loads in flight only block the warps' row sums.)
    __syncthreads();
}

__device__ __forceinline__ void spmv_smem_init(SpmvSmem &sm, PipeState &ps) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&sm.full[s], 1);
        mbar_init_fence();
    }
    ps.parity = 0;
    __syncthreads();
}

}  // namespace qpb
