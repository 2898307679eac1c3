// spmv_core.cuh -- tiled CSR SpMV ("CSR-stream") device core, used by the stand-alone SpMV kernel
// and by every matrix phase of the persistent ADMM kernel.
//
// Replaces SparseArrays/MKLSparse `mul!` (reference call sites LinearSystemSolvers.jl:135,139,
// 153-155 and SolveQuadraticProgram.jl:85-89).
//
// Layout (HBM): CSR with int32 row pointers / column indices and FP64 values.  The row range is cut
// on the host into *tiles*: maximal runs of consecutive rows holding <= kTileNnz non-zeros (a row
// longer than that is split into segments that stay on one CTA).  Each CTA owns a contiguous,
// nnz-balanced run of tiles (static => bitwise reproducible).  Per tile:
//   1. the tile's (col,val) stream is brought in fully coalesced -- either by a 1-D TMA bulk copy
//      (cp.async.bulk -> shared memory, mbarrier-tracked, multi-stage) or by coalesced LDGs,
//   2. every thread forms products val[k] * x[col[k]] (x gathered through L1/L2) into shared memory,
//   3. `lpr` lanes per row sum the row's products out of shared memory and hand the row sum to the
//      phase's epilogue functor (which fuses the vector update / norm / dot product of that phase).
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
constexpr int kTileNnz = QPB_TILE_NNZ;   // non-zeros per tile (build-time tunable for A/B runs)
constexpr int kTilePad = 8;             // alignment slack of a TMA-staged tile
constexpr int kStages = QPB_STAGES;      // TMA pipeline depth
constexpr int kTileCap = kTileNnz + kTilePad;
constexpr int kGatherBatch = kTileNnz / kThreads;   // independent x-gathers issued back to back per thread

// tile descriptor: x = first row, y = #rows, z = first nnz (k0), w = #nnz | flags
constexpr int kTileContFromPrev = 1 << 30;   // this tile continues a long row started earlier
constexpr int kTileContToNext = 1 << 29;     // the long row continues into the next tile
constexpr int kTileNkMask = (1 << 24) - 1;

struct CsrTiled {
    int rows, cols;
    const int *rowptr;     // rows + 1
    const int *rowmid;     // rows, optional: first nnz of the second column block (split sums)
    const int *col;        // nnz (+ padding)
    const double *val;     // nnz (+ padding)
    const int4 *tiles;     // ntiles
    const int *cta_begin;  // grid + 1: tile range of every CTA
    int ntiles;
    int lpr;               // lanes per row in the row-sum step (power of two, <= 32)
};

// Shared memory of one CTA.
struct __align__(128) SpmvSmem {
#ifdef QPB_SMEM_LEAN
    double val[2][kTileCap];         // lean variant: only the two product buffers live in shared memory
#else
    double val[kStages][kTileCap];   // TMA: staged values, overwritten in place by the products
#endif
                                     // LDG: buffers 0/1 hold the products (ping-pong)
#if !defined(QPB_COL_LDG) && !defined(QPB_SMEM_LEAN)
    int col[kStages][kTileCap];      // TMA only
#endif
    uint64_t full[kStages];          // mbarriers
    double red[kWarps * kMaxRed];
    double bcast[kMaxRed];
    double carry[2];                 // running sums of a long row
};

struct PipeState {     // uniform across the CTA
    uint32_t parity;   // bit s = parity to wait for on stage s
};

// Row pointers of the first row-sum round, fetched *before* the gather phase so that their L2 latency
// overlaps the gathers (ncu r1: 13 % of the stall samples sat on these loads).
struct RowPre {
    int a, b, mid;
};

template <bool SPLIT>
__device__ __forceinline__ RowPre prefetch_rowptr(const CsrTiled &M, const int4 td) {
    RowPre rp{0, 0, 0};
    const int r = threadIdx.x / M.lpr;
    if (r < td.y && !(td.w & (kTileContFromPrev | kTileContToNext))) {
        rp.a = __ldg(M.rowptr + td.x + r);
        rp.b = __ldg(M.rowptr + td.x + r + 1);
        if (SPLIT) rp.mid = __ldg(M.rowmid + td.x + r);
    }
    return rp;
}

// ---- row-sum step shared by both loaders ---------------------------------------------------
// prod: shared products of this tile, element k of the tile at prod[k] (k relative to k0).
template <bool SPLIT, class Epi>
__device__ __forceinline__ void tile_row_sums(const CsrTiled &M, const int4 td, const double *prod, SpmvSmem &sm,
                                              const RowPre pre, Epi &epi) {
    const int row0 = td.x, nrows = td.y, k0 = td.z;
    const int nk = td.w & kTileNkMask;
    const bool from_prev = td.w & kTileContFromPrev, to_next = td.w & kTileContToNext;
    if (from_prev || to_next) {
        // one segment of a long row: whole-CTA sum, carried across the row's tiles
        double s[2] = {0.0, 0.0};
        const int mid = SPLIT ? (M.rowmid[row0] - k0) : nk;
        for (int k = threadIdx.x; k < nk; k += kThreads) {
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
    const int lpr = M.lpr;
    const int groups = kThreads / lpr;
    const int g = threadIdx.x / lpr, gl = threadIdx.x % lpr;
    for (int rb = 0; rb < nrows; rb += groups) {
        const int r = rb + g;
        double s0 = 0.0, s1 = 0.0;
        if (r < nrows) {
            const int row = row0 + r;
            const int a = (rb == 0 ? pre.a : __ldg(M.rowptr + row)) - k0;
            const int b = (rb == 0 ? pre.b : __ldg(M.rowptr + row + 1)) - k0;
            if (!SPLIT) {
                for (int k = a + gl; k < b; k += lpr) s0 += prod[k];
            } else {
                const int mid = (rb == 0 ? pre.mid : __ldg(M.rowmid + row)) - k0;
                for (int k = a + gl; k < mid; k += lpr) s0 += prod[k];
                for (int k = mid + gl; k < b; k += lpr) s1 += prod[k];
            }
        }
        for (int o = lpr >> 1; o > 0; o >>= 1) {
            s0 += __shfl_down_sync(0xffffffffu, s0, o, lpr);
            if (SPLIT) s1 += __shfl_down_sync(0xffffffffu, s1, o, lpr);
        }
        if (r < nrows && gl == 0) epi(row0 + r, s0, s1);
    }
}

// ---- loader 1: coalesced LDG tiles -----------------------------------------------------------
// x: the gathered vector (mutable between phases: plain loads, never ld.global.nc).
template <bool SPLIT, class Epi>
__device__ __forceinline__ void spmv_tiles_ldg(const CsrTiled &M, const double *x, SpmvSmem &sm, Epi &epi) {
    const int tb = M.cta_begin[blockIdx.x], te = M.cta_begin[blockIdx.x + 1];
    for (int t = tb; t < te; ++t) {
        const int4 td = __ldg(M.tiles + t);
        const int k0 = td.z, nk = td.w & kTileNkMask;
        double *prod = sm.val[(t - tb) & 1];
        const RowPre pre = prefetch_rowptr<SPLIT>(M, td);
        const int *col = M.col + k0;
        const double *val = M.val + k0;
        for (int kb = threadIdx.x; kb < nk; kb += kThreads * kGatherBatch) {
            int c[kGatherBatch];
            double v[kGatherBatch], xv[kGatherBatch];
#pragma unroll
            for (int j = 0; j < kGatherBatch; ++j) {
                const int k = kb + j * kThreads;
                c[j] = (k < nk) ? __ldg(col + k) : 0;
                v[j] = (k < nk) ? __ldg(val + k) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < kGatherBatch; ++j) xv[j] = (kb + j * kThreads < nk) ? x[c[j]] : 0.0;   // independent gathers in flight
#pragma unroll
            for (int j = 0; j < kGatherBatch; ++j)
                if (kb + j * kThreads < nk) prod[kb + j * kThreads] = v[j] * xv[j];
        }
        __syncthreads();
        tile_row_sums<SPLIT>(M, td, prod, sm, pre, epi);
    }
    __syncthreads();
}

// ---- loader 1b (build variant QPB_SMEM_LEAN): register-prefetched stream, shared memory only for products ---
// Every KB of shared memory is a KB less L1, and on B200 the random x-gathers collapse once the sectors
// in flight exceed the L1 capacity (profiles/r1_gather_ceiling_smem_sweep.txt).  This loader keeps just two
// product buffers in shared memory; the (col, val) stream of tile i+1 is fetched into registers (no L1
// allocation) while the gathers of tile i are in flight.
__device__ __forceinline__ int ldg_stream_i32(const int *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldg_stream_f64(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

template <bool SPLIT, class Epi>
__device__ __forceinline__ void spmv_tiles_ldg_pf(const CsrTiled &M, const double *x, SpmvSmem &sm, Epi &epi) {
    static_assert(kTileNnz == kThreads * kGatherBatch, "one gather batch must cover a tile");
    const int tb = M.cta_begin[blockIdx.x], te = M.cta_begin[blockIdx.x + 1];
    const int nt = te - tb;
    int4 td_next = nt > 0 ? __ldg(M.tiles + tb) : make_int4(0, 0, 0, 0);
    int cn[kGatherBatch];
    double vn[kGatherBatch];
    auto fetch = [&](const int4 &tdi) {
        const int nk = tdi.w & kTileNkMask;
#pragma unroll
        for (int j = 0; j < kGatherBatch; ++j) {
            const int k = threadIdx.x + j * kThreads;
            cn[j] = (k < nk) ? ldg_stream_i32(M.col + tdi.z + k) : 0;
            vn[j] = (k < nk) ? ldg_stream_f64(M.val + tdi.z + k) : 0.0;
        }
    };
    if (nt > 0) fetch(td_next);
    for (int i = 0; i < nt; ++i) {
        const int4 td = td_next;
        if (i + 1 < nt) td_next = __ldg(M.tiles + tb + i + 1);
        const int nk = td.w & kTileNkMask;
        double *prod = sm.val[i & 1];
        const RowPre pre = prefetch_rowptr<SPLIT>(M, td);
        double v[kGatherBatch], xv[kGatherBatch];
#pragma unroll
        for (int j = 0; j < kGatherBatch; ++j) {
            v[j] = vn[j];
            xv[j] = (threadIdx.x + j * kThreads < nk) ? x[cn[j]] : 0.0;   // independent gathers in flight
        }
        if (i + 1 < nt) fetch(td_next);                                    // stream of the next tile, behind the gathers
#pragma unroll
        for (int j = 0; j < kGatherBatch; ++j) {
            const int k = threadIdx.x + j * kThreads;
            if (k < nk) prod[k] = v[j] * xv[j];
        }
        __syncthreads();
        tile_row_sums<SPLIT>(M, td, prod, sm, pre, epi);
    }
    __syncthreads();
}

// ---- loader 2: TMA bulk-copy staged tiles ------------------------------------------------------
__device__ __forceinline__ void tma_issue_tile(const CsrTiled &M, const int4 td, SpmvSmem &sm, int stage) {
    const int k0 = td.z, nk = td.w & kTileNkMask;
    const int k0a = k0 & ~3;                            // 16-byte aligned start for both arrays
    const int cnt = ((k0 + nk - k0a) + 3) & ~3;         // <= kTileNnz + 6
#if !defined(QPB_COL_LDG) && !defined(QPB_SMEM_LEAN)
    mbar_expect_tx(&sm.full[stage], static_cast<uint32_t>(cnt) * 12u);
    tma_load_1d(sm.val[stage], M.val + k0a, static_cast<uint32_t>(cnt) * 8u, &sm.full[stage]);
    tma_load_1d(sm.col[stage], M.col + k0a, static_cast<uint32_t>(cnt) * 4u, &sm.full[stage]);
#else
    mbar_expect_tx(&sm.full[stage], static_cast<uint32_t>(cnt) * 8u);
    tma_load_1d(sm.val[stage], M.val + k0a, static_cast<uint32_t>(cnt) * 8u, &sm.full[stage]);
#endif
}

template <bool SPLIT, class Epi>
__device__ __forceinline__ void spmv_tiles_tma(const CsrTiled &M, const double *x, SpmvSmem &sm, PipeState &ps,
                                               Epi &epi) {
    const int tb = M.cta_begin[blockIdx.x], te = M.cta_begin[blockIdx.x + 1];
    const int nt = te - tb;
    // the stages were last touched through the generic proxy (previous phase's row sums)
    fence_proxy_async_smem();
    __syncthreads();
#ifndef QPB_REFILL_AFTER_ROWSUMS
    constexpr int kAhead = kStages - 1;     // tiles in flight ahead of the one being processed
#else
    constexpr int kAhead = kStages;
#endif
    if (threadIdx.x == 0) {
        const int pre = nt < kAhead ? nt : kAhead;
        for (int i = 0; i < pre; ++i) tma_issue_tile(M, __ldg(M.tiles + tb + i), sm, i % kStages);
    }
    // Tile descriptors are fetched ONE ITERATION AHEAD: a descriptor load issued when it is needed queues
    // behind the thousands of x-gathers the SM has in flight (L1 is thrashed by them, so it is an L2 round
    // trip under load) and serialises the tile loop; the same holds for the descriptor thread 0 needs to
    // issue the next TMA refill.
    int4 td_next = nt > 0 ? __ldg(M.tiles + tb) : make_int4(0, 0, 0, 0);
#ifdef QPB_COL_LDG
    // column indices are NOT staged in shared memory (every KB of shared memory is a KB less L1, and L1
    // capacity bounds the number of gathers in flight): they are read with coalesced loads one tile ahead
    int cnext[kGatherBatch];
#pragma unroll
    for (int j = 0; j < kGatherBatch; ++j) {
        const int k = threadIdx.x + j * kThreads;
        cnext[j] = (nt > 0 && k < (td_next.w & kTileNkMask)) ? __ldg(M.col + td_next.z + k) : 0;
    }
#endif
    int4 td_issue = (threadIdx.x == 0 && kAhead < nt) ? __ldg(M.tiles + tb + kAhead) : make_int4(0, 0, 0, 0);
    for (int i = 0; i < nt; ++i) {
        const int s = i % kStages;
        const int4 td = td_next;
        const int4 td_refill = td_issue;
        if (i + 1 < nt) td_next = __ldg(M.tiles + tb + i + 1);
        if (threadIdx.x == 0 && i + 1 + kAhead < nt) td_issue = __ldg(M.tiles + tb + i + 1 + kAhead);
        const int k0 = td.z, nk = td.w & kTileNkMask;
        const int off = k0 & 3;
        const RowPre pre = prefetch_rowptr<SPLIT>(M, td);
        mbar_wait(&sm.full[s], (ps.parity >> s) & 1u);
        ps.parity ^= (1u << s);
        double *val = sm.val[s] + off;
#ifndef QPB_COL_LDG
        const int *col = sm.col[s] + off;
#endif
        for (int kb = threadIdx.x; kb < nk; kb += kThreads * kGatherBatch) {
            int c[kGatherBatch];
            double xv[kGatherBatch];
#pragma unroll
            for (int j = 0; j < kGatherBatch; ++j) {
#ifndef QPB_COL_LDG
                const int k = kb + j * kThreads;
                c[j] = (k < nk) ? col[k] : 0;
#else
                c[j] = cnext[j];
#endif
            }
#pragma unroll
            for (int j = 0; j < kGatherBatch; ++j) xv[j] = (kb + j * kThreads < nk) ? x[c[j]] : 0.0;   // independent gathers in flight
#ifdef QPB_COL_LDG
            if (i + 1 < nt) {
#pragma unroll
                for (int j = 0; j < kGatherBatch; ++j) {
                    const int k = threadIdx.x + j * kThreads;
                    cnext[j] = (k < (td_next.w & kTileNkMask)) ? __ldg(M.col + td_next.z + k) : 0;
                }
            }
#endif
#pragma unroll
            for (int j = 0; j < kGatherBatch; ++j)
                if (kb + j * kThreads < nk) val[kb + j * kThreads] *= xv[j];
        }
#ifndef QPB_REFILL_AFTER_ROWSUMS
        fence_proxy_async_smem();   // our generic accesses to the stages (incl. tile i-1's reads) before the refill
        __syncthreads();
        if (threadIdx.x == 0 && i + kStages - 1 < nt) tma_issue_tile(M, td_refill, sm, (i + kStages - 1) % kStages);
        tile_row_sums<SPLIT>(M, td, val, sm, pre, epi);
#else
        // variant for 2 big stages: the stage of tile i is handed back right after ITS row sums (one more
        // barrier per tile), so the refill (tile i + kStages) overlaps the whole next tile
        __syncthreads();
        tile_row_sums<SPLIT>(M, td, val, sm, pre, epi);
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0 && i + kStages < nt) tma_issue_tile(M, td_refill, sm, s);
#endif
    }
    __syncthreads();
}

// ---- loader 3: TMA staged + software pipelined ---------------------------------------------------
// Same staging as loader 2, but the gathers of tile i+1 are issued (into registers) BEFORE the row sums
// of tile i, so their L2 latency overlaps the row-sum step and the tile barrier of this CTA instead of
// relying on the other resident CTA (ncu r1: 16 % of the stall samples at the tile barrier, ~20 % on the
// gather scoreboard).  Safe because no phase writes the vector it gathers from.
template <bool SPLIT, class Epi>
__device__ __forceinline__ void spmv_tiles_tma_pipe(const CsrTiled &M, const double *x, SpmvSmem &sm, PipeState &ps,
                                                    Epi &epi) {
    static_assert(kTileNnz == kThreads * kGatherBatch, "one gather batch must cover a tile");
#ifdef QPB_COL_LDG
    spmv_tiles_tma<SPLIT>(M, x, sm, ps, epi);
#else
    if (kStages < 3) {   // the refill targets the stage of tile i-1 while tile i+1 is being read
        spmv_tiles_tma<SPLIT>(M, x, sm, ps, epi);
        return;
    }
    const int tb = M.cta_begin[blockIdx.x], te = M.cta_begin[blockIdx.x + 1];
    const int nt = te - tb;
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int pre = nt < (kStages - 1) ? nt : (kStages - 1);
        for (int i = 0; i < pre; ++i) tma_issue_tile(M, __ldg(M.tiles + tb + i), sm, i % kStages);
    }
    double v[kGatherBatch], xv[kGatherBatch];
    auto load_tile = [&](int i, const int4 &tdi) {     // wait for the stage, read (col, val), issue the gathers
        const int s = i % kStages;
        mbar_wait(&sm.full[s], (ps.parity >> s) & 1u);
        ps.parity ^= (1u << s);
        const int nk = tdi.w & kTileNkMask, off = tdi.z & 3;
        const double *val = sm.val[s] + off;
        const int *col = sm.col[s] + off;
        int c[kGatherBatch];
#pragma unroll
        for (int j = 0; j < kGatherBatch; ++j) {
            const int k = threadIdx.x + j * kThreads;
            c[j] = (k < nk) ? col[k] : 0;
            v[j] = (k < nk) ? val[k] : 0.0;
        }
        // volatile: the loads must be ISSUED here (before the row sums of the previous tile), not sunk to their use
#pragma unroll
        for (int j = 0; j < kGatherBatch; ++j) {
            xv[j] = 0.0;
            if (threadIdx.x + j * kThreads < nk) asm volatile("ld.global.f64 %0, [%1];" : "=d"(xv[j]) : "l"(x + c[j]) : "memory");
        }
    };
    int4 td = make_int4(0, 0, 0, 0);
    if (nt > 0) {
        td = __ldg(M.tiles + tb);
        load_tile(0, td);
    }
    for (int i = 0; i < nt; ++i) {
        const int s = i % kStages;
        const int nk = td.w & kTileNkMask, off = td.z & 3;
        const RowPre pre = prefetch_rowptr<SPLIT>(M, td);
        double *val = sm.val[s] + off;
#pragma unroll
        for (int j = 0; j < kGatherBatch; ++j) {
            const int k = threadIdx.x + j * kThreads;
            if (k < nk) val[k] = v[j] * xv[j];
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0 && i + kStages - 1 < nt)
            tma_issue_tile(M, __ldg(M.tiles + tb + i + kStages - 1), sm, (i + kStages - 1) % kStages);
        int4 tdn = td;
        if (i + 1 < nt) {
            tdn = __ldg(M.tiles + tb + i + 1);
            load_tile(i + 1, tdn);
        }
        tile_row_sums<SPLIT>(M, td, val, sm, pre, epi);
        td = tdn;
    }
    __syncthreads();
#endif
}

// LOADER: 0 = coalesced LDG, 1 = TMA staged, 2 = TMA staged + software pipelined gathers
template <int LOADER, bool SPLIT, class Epi>
__device__ __forceinline__ void spmv_tiles(const CsrTiled &M, const double *x, SpmvSmem &sm, PipeState &ps, Epi &epi) {
#ifdef QPB_SMEM_LEAN
    spmv_tiles_ldg_pf<SPLIT>(M, x, sm, epi);
#else
    if (LOADER == 2) spmv_tiles_tma_pipe<SPLIT>(M, x, sm, ps, epi);
    else if (LOADER == 1) spmv_tiles_tma<SPLIT>(M, x, sm, ps, epi);
    else spmv_tiles_ldg<SPLIT>(M, x, sm, epi);
#endif
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
