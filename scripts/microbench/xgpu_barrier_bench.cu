// xgpu_barrier_bench.cu -- latency of a cross-GPU flag handshake over NVLink peer memory (2 GPUs, one process):
// each GPU runs one kernel that K times {st.release.sys epoch -> peer flag ; spin ld.acquire.sys own flag}.
// Variants: flag polled in LOCAL memory (peer writes into it) vs polled in REMOTE memory (peer writes locally).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a xgpu_barrier_bench.cu -o xgpu_barrier_bench
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ unsigned long long ld_acq_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_rel_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// mode 0: push (write peer's flag, poll own local flag) with release/acquire
// mode 1: pull (write own local flag, poll the peer's flag remotely)
// mode 2: push with relaxed store/load + explicit fence.acq_rel.sys
__global__ void handshake(unsigned long long *mine, unsigned long long *theirs, int K, int mode, unsigned long long *out_ns) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int k = 1; k <= K; ++k) {
        if (mode == 0) {
            st_rel_sys(theirs, (unsigned long long)k);
            while (ld_acq_sys(mine) < (unsigned long long)k) {}
        } else if (mode == 1) {
            st_rel_sys(mine, (unsigned long long)k);
            while (ld_acq_sys(theirs) < (unsigned long long)k) {}
        } else {
            asm volatile("fence.acq_rel.sys;" ::: "memory");
            st_relaxed_sys(theirs, (unsigned long long)k);
            while (ld_relaxed_sys(mine) < (unsigned long long)k) {}
            asm volatile("fence.acq_rel.sys;" ::: "memory");
        }
    }
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    *out_ns = t1 - t0;
}

int main() {
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    if (ndev < 2) { printf("need 2 GPUs\n"); return 0; }
    unsigned long long *flag[2], *out[2];
    for (int d = 0; d < 2; ++d) {
        cudaSetDevice(d);
        cudaDeviceEnablePeerAccess(1 - d, 0);
        cudaMalloc(&flag[d], 256);
        cudaMalloc(&out[d], 8);
    }
    const int K = 20000;
    const char *names[] = {"push, release/acquire.sys", "pull (remote poll)", "push, relaxed + fence.acq_rel.sys"};
    for (int mode = 0; mode < 3; ++mode) {
        for (int d = 0; d < 2; ++d) { cudaSetDevice(d); cudaMemset(flag[d], 0, 256); cudaDeviceSynchronize(); }
        for (int d = 0; d < 2; ++d) { cudaSetDevice(d); handshake<<<1, 32>>>(flag[d], flag[1 - d], K, mode, out[d]); }
        unsigned long long ns[2];
        for (int d = 0; d < 2; ++d) { cudaSetDevice(d); cudaDeviceSynchronize(); cudaMemcpy(&ns[d], out[d], 8, cudaMemcpyDeviceToHost); }
        printf("%-36s %8.2f us per handshake (gpu0), %8.2f (gpu1)   err=%s\n", names[mode], ns[0] / 1e3 / K, ns[1] / 1e3 / K,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
