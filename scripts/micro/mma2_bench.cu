// Microbenchmark + correctness probe for the CTA-pair MMA (tcgen05.mma.cta_group::2, M = 256): two CTAs of a cluster each
// hold 128 rows of A and HALF of B's rows (N/2 output columns) in their own shared memory, the leader CTA issues the MMA,
// each CTA's TMEM receives its 128 rows x N columns of D.  Questions for the next kernel generation:
// which half of B / which rows of D belong to which CTA, and how many cycles an M256 N{128,256} K16 MMA takes per SM
// compared with the 128 cycles of cta_group::1 M128 N256.  Every wait is bounded: a wrong guess ends in a message, not a hang.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/micro/mma2_bench scripts/micro/mma2_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../kzero_b200/csrc/tc_common.cuh"
using namespace kzb::tc;

constexpr int kAOff = 0;          // A: 128 rows x 64 k, SWIZZLE_128B, 16 KB
constexpr int kBOff = 20480;      // B half: up to 128 rows x 64 k, SWIZZLE_128B, 16 KB (A unswizzled: 8 x 152 x 16 = 19456 bytes)

__device__ __forceinline__ bool bounded_wait(uint64_t* bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 22); spin++) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return true;
    }
    return false;
}

// img: per CTA rank, 32 KB image of (A, B half).  mode 0: one K=64 pass, D -> out;  mode 1: timing
// a_nosw: A is laid out unswizzled, [k-chunk of 8][row][16 bytes] with kARows rows per k-chunk (the halo-tile operand of
// conv_tch.cu: LBO = kARows * 16, SBO = 128), and read from row `a_row0`
constexpr int kARows = 152;
__global__ void __launch_bounds__(128, 1) bench2(const uint8_t* img, int n, int iters, float* out, unsigned long long* cyc, int* status, int a_nosw,
                                                 int a_row0) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, peer_ready;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = cluster_ctarank();
    const uint8_t* mine = img + size_t(rank) * 40960;
    for (int i = threadIdx.x; i < 40960 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = reinterpret_cast<const uint32_t*>(mine)[i];
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_init(&peer_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // both CTAs' operands and barriers are in place
    tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    bool ok = true;
    if (warp == 1) {
        // the handshake a pipelined kernel needs per stage: the peer tells the leader "my half of the operands is in place"
        // with a remote arrive on the leader's barrier (here once; the cluster barrier above already made it true)
        if (rank == 1 && lane == 0) {
            uint32_t remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(smem_u32(&peer_ready)));
            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
        }
        if (rank == 0) {
            bool peer_ok = false;
            for (int spin = 0; spin < (1 << 22) && !peer_ok; spin++) {
                uint32_t done;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done)
                    : "r"(smem_u32(&peer_ready)), "r"(0u)
                    : "memory");
                peer_ok = done != 0;
            }
            if (!peer_ok && lane == 0) atomicExch(status, 2);
            tc_fence_after();
            const uint32_t idesc = umma_idesc_bf16(256, n);
            const uint64_t hi = umma_desc_sw128_hi();
            const uint32_t a_lo = umma_desc_lo(smem_u32(smem + kAOff)) + (a_nosw ? uint32_t(a_row0) : 0u), b_lo = umma_desc_lo(smem_u32(smem + kBOff));
            const uint64_t a_hi = a_nosw ? ((uint64_t(kARows * 16 >> 4) << 16) | (uint64_t(128 >> 4) << 32) | (uint64_t(1) << 46)) : hi;
            const uint32_t a_step = a_nosw ? uint32_t(2 * kARows * 16 >> 4) : 2u;  // 16 channels further along K
            const unsigned long long t0 = clock64();
            for (int it = 0; it < iters; it++) {
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint64_t da = a_hi | uint64_t(a_lo + a_step * j), db = hi | uint64_t(b_lo + 2 * j);
                        const uint32_t acc = (it | j) != 0;
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                            "l"(da), "l"(db), "r"(idesc), "r"(acc)
                            : "memory");
                    }
                }
                __syncwarp();
            }
            if (lane == 0)
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                             "h"(uint16_t(3))
                             : "memory");
            __syncwarp();
            ok = bounded_wait(&bar, 0);
            const unsigned long long t1 = clock64();
            if (lane == 0 && cyc) cyc[blockIdx.x / 2] = t1 - t0;
        } else {
            ok = bounded_wait(&bar, 0);
        }
        if (!ok && lane == 0) atomicExch(status, 1);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (out && blockIdx.x < 2) {
        const uint32_t taddr = tmem + (uint32_t(warp * 32) << 16);
        for (int c0 = 0; c0 < n; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(taddr + c0, r);
            tmem_ld_wait();
            for (int j = 0; j < 32; j++) out[(size_t(rank) * 128 + warp * 32 + lane) * 256 + c0 + j] = __uint_as_float(r[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
    }
}

static float bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

static void put_sw128(std::vector<uint8_t>& img, size_t base, int r, int k, float v) {
    // SWIZZLE_128B K-major [rows][64 k]: 16-byte chunk c of row r at r*128 + ((c ^ (r & 7)) * 16)
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    const size_t off = base + size_t(r) * 128 + ((size_t(k / 8) ^ size_t(r & 7)) * 16) + size_t(k % 8) * 2;
    memcpy(&img[off], &h, 2);
}

static cudaError_t launch(int grid, const uint8_t* d_img, int n, int iters, float* out, unsigned long long* cyc, int* status, int a_nosw, int a_row0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(grid));
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 48 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, bench2, d_img, n, iters, out, cyc, status, a_nosw, a_row0);
    if (e != cudaSuccess) return e;
    return cudaDeviceSynchronize();
}

int main() {
    uint8_t* d_img;
    float* d_out;
    unsigned long long* d_cyc;
    int* d_status;
    cudaMalloc(&d_img, 81920);
    cudaMalloc(&d_out, 256 * 256 * 4);
    cudaMalloc(&d_cyc, 148 * 8);
    cudaMalloc(&d_status, 4);
    cudaFuncSetAttribute(bench2, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(bench2, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cfg_i = 0; cfg_i < 4; cfg_i++) {
        const int n = cfg_i % 2 == 0 ? 256 : 128, a_nosw = cfg_i / 2, a_row0 = a_nosw ? 11 : 0;  // row 11: a tap-shifted start inside the halo tile
        printf("---- A operand %s\n", a_nosw ? "UNSWIZZLED halo tile (152 rows per k-chunk), read from row 11" : "SWIZZLE_128B");
        std::vector<float> A(256 * 64), B(size_t(n) * 64);
        srand(n);
        for (auto& v : A) v = bf((rand() % 200 - 100) / 64.0f);
        for (auto& v : B) v = bf((rand() % 200 - 100) / 64.0f);
        std::vector<uint8_t> img(81920, 0);
        for (int rank = 0; rank < 2; rank++) {
            for (int r = 0; r < 128; r++)
                for (int k = 0; k < 64; k++) {
                    const float v = A[size_t(rank * 128 + r) * 64 + k];
                    if (!a_nosw) {
                        put_sw128(img, size_t(rank) * 40960 + kAOff, r, k, v);
                    } else {
                        __nv_bfloat16 h = __float2bfloat16_rn(v);
                        const size_t off = size_t(rank) * 40960 + kAOff + size_t(k / 8) * kARows * 16 + size_t(a_row0 + r) * 16 + size_t(k % 8) * 2;
                        memcpy(&img[off], &h, 2);
                    }
                }
            for (int r = 0; r < n / 2; r++)
                for (int k = 0; k < 64; k++) put_sw128(img, size_t(rank) * 40960 + kBOff, r, k, B[size_t(rank * (n / 2) + r) * 64 + k]);
        }
        cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
        cudaMemset(d_out, 0, 256 * 256 * 4);
        cudaMemset(d_status, 0, 4);
        cudaError_t e = launch(2, d_img, n, 1, d_out, nullptr, d_status, a_nosw, a_row0);
        int status = 0;
        cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost);
        printf("N=%d correctness launch: %s%s\n", n, cudaGetErrorString(e),
               status == 2 ? "  (the peer's remote arrive never reached the leader)" : status ? "  (a barrier wait timed out)" : "  (remote arrive handshake ok)");
        if (e != cudaSuccess) return 1;
        std::vector<float> out(256 * 256);
        cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
        // hypotheses: D rows of CTA r = A rows of CTA (r ^ swap_m); D columns [0, n/2) = B rows of CTA (0 ^ swap_n)
        for (int swap_m = 0; swap_m < 2; swap_m++)
            for (int swap_n = 0; swap_n < 2; swap_n++) {
                double max_err = 0;
                for (int m = 0; m < 256; m++)
                    for (int c = 0; c < n; c++) {
                        const int am = ((m / 128) ^ swap_m) * 128 + m % 128;
                        const int bn = ((c / (n / 2)) ^ swap_n) * (n / 2) + c % (n / 2);
                        double ref = 0;
                        for (int k = 0; k < 64; k++) ref += double(A[size_t(am) * 64 + k]) * B[size_t(bn) * 64 + k];
                        max_err = std::fmax(max_err, std::fabs(ref - out[size_t(m) * 256 + c]));
                    }
                printf("  hypothesis rows %s, columns %s: max |err| %.3g%s\n", swap_m ? "swapped" : "own CTA", swap_n ? "swapped" : "CTA0 first", max_err,
                       max_err < 1e-2 ? "   <== matches" : "");
            }
        const int iters = 1024;
        cudaMemset(d_status, 0, 4);
        e = launch(148, d_img, n, iters, nullptr, d_cyc, d_status, a_nosw, a_row0);
        unsigned long long c[2];
        cudaMemcpy(c, d_cyc, 16, cudaMemcpyDeviceToHost);
        cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost);
        printf("N=%d timing, 74 CTA pairs: %.1f cycles per M256 N%d K16 MMA (each SM does 128 x %d x 16; cta_group::1 M128 N256 takes 128)  %s%s\n", n,
               double(c[0]) / (iters * 4), n, n, cudaGetErrorString(e), status ? "  (timed out)" : "");
    }
    return 0;
}
