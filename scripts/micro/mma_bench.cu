// Microbenchmark: issue rate / duration of tcgen05.mma (kind::f16, bf16, cta_group::1) as a function of N,
// with and without per-group commits.  Development aid for the tower kernel design.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../kzero_b200/csrc/tc_common.cuh"
using namespace kzb::tc;

__global__ void __launch_bounds__(128, 1) bench(int m, int n, int iters, int commit_every, unsigned long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    if (warp == 1) {
        const uint32_t idesc = umma_idesc_bf16(m, n);
        const uint64_t hi = umma_desc_sw128_hi();
        const uint32_t a_lo = umma_desc_lo(smem_u32(smem)), b_lo = umma_desc_lo(smem_u32(smem + 16384));
        unsigned long long t0 = clock64();
        uint32_t phase = 0;
        int pending = 0;
        for (int i = 0; i < iters; i += 4) {
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 4; k++) umma_bf16(tmem + ((i >> 2) & 1) * 256, hi | uint64_t(a_lo + 2 * k), hi | uint64_t(b_lo + 2 * k), idesc, 1u);
                if (commit_every && (((i + 4) & (commit_every - 1)) == 0)) umma_commit(&bar[1]);
            }
            __syncwarp();
        }
        unsigned long long t1 = clock64();
        if (lane == 0) umma_commit(&bar[0]);
        __syncwarp();
        mbar_wait(&bar[0], phase);
        unsigned long long t2 = clock64();
        if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
        (void)pending;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory"); }
}

int main() {
    unsigned long long* d; cudaMalloc(&d, 148 * 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 4096;
    for (int grid : {148}) for (int m : {128}) for (int n : {64, 128, 192, 256}) for (int ce : {0, 4, 16}) {
        bench<<<grid, 128, 60 * 1024>>>(m, n, iters, ce, d);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("grid %3d M %3d N %3d commit_every %d : issue %.1f cyc/mma, complete %.1f cyc/mma (ideal %d)  %s\n", grid, m, n, ce,
               double(h[0]) / iters, double(h[1]) / iters, (m == 64 ? 128 : m) * n / 256, cudaGetErrorString(e));
    }
    return 0;
}
