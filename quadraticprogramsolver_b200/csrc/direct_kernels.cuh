// direct_kernels.cuh -- the exact x~ step for ONE sparse QP of moderate size (settings.lin_solver =
// QPB200_LINSOLVE_CHOLESKY on qpb200_create): what the reference's direct plugins do,
//     LaLdlInit/LaLdl!, QDLdlInit/QDLdl!, FacLdlInit/FacLdl!     LinearSystemSolvers.jl:16-107
// i.e. the exact solution of [P + sigma I, A'; A, -rho^-1 I] [x~; nu] = [sigma x - q; z - rho^-1 y], in its reduced
// form (eliminate nu): K x~ = sigma x - q + A'(rho z - y), K = P + sigma I + A' diag(rho_i) A, z~ = A x~.
// FacLdl is the plugin the reference's own tests and benchmarks run (RunTests.jl:55-56, RunBenchmarks.jl:54-55).
//
// B200 design: the factorisation is needed once per rho (a handful of times per solve, LinearSystemSolvers.jl:30-32,
// 61-63, 93-95) while the solve runs thousands of times, and a sparse triangular solve is a latency chain of n
// dependent steps.  So K is formed DENSE in HBM (n <= 32768: 8.6 GB of 180) and inverted in place by a blocked
// symmetric sweep (block Gauss-Jordan without pivoting: every pivot block is a Schur complement of an SPD matrix,
// hence SPD), whose rank-32 trailing updates -- all of the n^3 flops -- run on the FP64 tensor pipe
// (mma.sync.m8n8k4.f64 -> SASS DMMA), lower tiles only, mirrored through shared memory.  Each ADMM iteration is then
// one streaming pass over -K^-1 (HBM bound for large n, one grid barrier for small n) inside the persistent kernel,
// applied as ONE STEP OF ITERATIVE REFINEMENT from the previous x~:  x~ += K^-1 (b - K x~), with b - K x~ coming out of
// the same H pass the CG path starts with -- the rounding error of the explicit inverse is multiplied by the
// (shrinking) distance between consecutive x~ instead of by |x~|.
//
// Sweep on pivot block k (D = K_kk^-1):  K_kk <- -D,  K_ik <- K_ik D,  K_ij <- K_ij - K_ik D K_kj  (i, j != k).
// Symmetry is kept at every step; after all blocks K holds -K^-1.
#pragma once
#include "spmv_core.cuh"

namespace qpb {

constexpr int kGjNb = 32;     // pivot block width
constexpr int kGjTile = 64;   // trailing-update tile
constexpr int kGjLds = 36;    // shared-memory leading dimension of a 32-wide panel (conflict-free DMMA fragments)

// x~ -= Kneg r  with Kneg = -K^-1 (dense, row-major, symmetric): one warp per row, fixed summation order.
// Called by every CTA of the persistent kernel; r was completed before the preceding grid barrier.
__device__ __forceinline__ void dense_symv_sub(const double *Kneg, int ld, int n, const double *r, double *xt) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * kWarps + (threadIdx.x >> 5), nw = gridDim.x * kWarps;
    for (int j = gw; j < n; j += nw) {
        const double *row = Kneg + (size_t)j * ld;
        double s0 = 0.0, s1 = 0.0;
        int c = 2 * lane;
        for (; c + 64 < n; c += 128) {
            const double2 k0 = __ldg(reinterpret_cast<const double2 *>(row + c));
            const double2 k1 = __ldg(reinterpret_cast<const double2 *>(row + c + 64));
            s0 += k0.x * r[c] + k0.y * r[c + 1];
            s1 += k1.x * r[c + 64] + k1.y * r[c + 65];
        }
        for (; c < n; c += 64) {                       // (r is padded to an even length; Kneg's padding columns are 0)
            const double2 k0 = __ldg(reinterpret_cast<const double2 *>(row + c));
            s0 += k0.x * r[c] + (c + 1 < n ? k0.y * r[c + 1] : 0.0);
        }
        double s = s0 + s1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (lane == 0) xt[j] -= s;
    }
}

}  // namespace qpb
