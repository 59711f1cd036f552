// Microbenchmark + correctness probe: tcgen05.mma with a SWIZZLE_NONE K-major B operand whose 8-row groups are
// 144 bytes apart (8 positions + one zero pad row), so that a conv tap (dy, dx) is only a different descriptor start
// address: dx = +-16 bytes, dy = +-4 groups.  Question: is it correct, and does the misaligned (dx != 0) fetch run at
// full MMA rate?  Development aid for the tower kernel design.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../kzero_b200/csrc/tc_common.cuh"
using namespace kzb::tc;

constexpr int kGroup = 144;             // bytes per 8-position group (8 x 16 B + 16 B zero pad)
constexpr int kLbo = 36 * kGroup;       // k-chunk stride: 32 data groups + 4 shared halo groups
constexpr int kBBase = 16384 + 1024;    // B region starts after the A tile (+ slack so that base - 4 groups - 16 is valid)
constexpr int kBData = kBBase + 4 * kGroup;  // first data group of k-chunk 0

__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr) {
    uint64_t d = 0;
    d |= uint64_t((addr >> 4) & 0x3FFF);
    d |= uint64_t((kLbo >> 4) & 0x3FFF) << 16;
    d |= uint64_t((kGroup >> 4) & 0x3FFF) << 32;
    d |= uint64_t(1) << 46;
    return d;  // layout type 0 = no swizzle
}

// mode 0: correctness (one K=64 pass for tap (dy,dx), D -> out);  mode 1: timing (iters MMAs with the given tap)
__global__ void __launch_bounds__(128, 1) bench(const uint8_t* init, int init_bytes, int dy, int dx, int iters, int sw_b, float* out,
                                                unsigned long long* cyc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    for (int i = threadIdx.x; i < init_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = reinterpret_cast<const uint32_t*>(init)[i];
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    const uint32_t idesc = umma_idesc_bf16(128, 256);
    const uint64_t hi = umma_desc_sw128_hi();
    const uint32_t a_lo = umma_desc_lo(smem_u32(smem));
    const uint32_t b_addr = smem_u32(smem + kBData) + dy * 4 * kGroup + dx * 16;
    const uint32_t bsw_lo = umma_desc_lo(smem_u32(smem + 65536));  // a SWIZZLE_128B B tile for the timing comparison
    if (warp == 1) {
        uint64_t da[4], db[4];
        for (int j = 0; j < 4; j++) {
            da[j] = hi | uint64_t(a_lo + 2 * j);
            db[j] = sw_b ? (hi | uint64_t(bsw_lo + 2 * j)) : desc_nosw(b_addr + j * 2 * kLbo);
        }
        unsigned long long t0 = clock64();
        if (dy == 9) {  // cycle through all nine taps like the tower kernel does
            const uint32_t base = smem_u32(smem + kBData);
            for (int it = 0; it < iters; it++) {
                const int tap = it % 9, ty = tap / 3 - 1, tx = tap % 3 - 1;
                const uint64_t b0 = desc_nosw(base + ty * 4 * kGroup + tx * 16);
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < 4; j++) umma_bf16(tmem, da[j], b0 + uint64_t(j * (2 * kLbo / 16)), idesc, (it | j) != 0);
                }
                __syncwarp();
            }
        } else
        for (int it = 0; it < iters; it++) {
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < 4; j++) umma_bf16(tmem, da[j], db[j], idesc, (it | j) != 0);
            }
            __syncwarp();
        }
        if (lane == 0) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        unsigned long long t1 = clock64();
        if (lane == 0 && cyc) cyc[blockIdx.x] = t1 - t0;
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (out && blockIdx.x == 0) {
        const uint32_t taddr = tmem + (uint32_t(warp * 32) << 16);
        for (int c0 = 0; c0 < 256; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(taddr + c0, r);
            tmem_ld_wait();
            for (int j = 0; j < 32; j++) out[(warp * 32 + lane) * 256 + c0 + j] = __uint_as_float(r[j]);
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory"); }
}

static float bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

int main() {
    const int smem_bytes = 200 * 1024;
    std::vector<uint8_t> img(smem_bytes, 0);
    std::vector<float> W(128 * 64), X(256 * 64);
    srand(1);
    for (auto& v : W) v = bf((rand() % 200 - 100) / 64.0f);
    for (auto& v : X) v = bf((rand() % 200 - 100) / 64.0f);
    // A: SWIZZLE_128B K-major [128 rows][64 k]: 16-byte chunk c of row r at r*128 + ((c ^ (r & 7)) * 16)
    for (int r = 0; r < 128; r++)
        for (int k = 0; k < 64; k++) {
            __nv_bfloat16 h = __float2bfloat16_rn(W[r * 64 + k]);
            size_t off = size_t(r) * 128 + ((size_t(k / 8) ^ (r & 7)) * 16) + (k % 8) * 2;
            memcpy(&img[off], &h, 2);
        }
    // B: position n = y*32 + board*8 + x  ->  group g = n / 8 (= y*4 + board), row x = n % 8; k-chunk kc = k / 8
    for (int n = 0; n < 256; n++)
        for (int k = 0; k < 64; k++) {
            __nv_bfloat16 h = __float2bfloat16_rn(X[n * 64 + k]);
            size_t off = size_t(kBData) + size_t(k / 8) * kLbo + size_t(n / 8) * kGroup + (n % 8) * 16 + (k % 8) * 2;
            memcpy(&img[off], &h, 2);
        }
    uint8_t* d_img; float* d_out; unsigned long long* d_cyc;
    cudaMalloc(&d_img, smem_bytes); cudaMalloc(&d_out, 128 * 256 * 4); cudaMalloc(&d_cyc, 148 * 8);
    cudaMemcpy(d_img, img.data(), smem_bytes, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    std::vector<float> out(128 * 256);
    int bad_total = 0;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            bench<<<1, 128, 210 * 1024>>>(d_img, smem_bytes, dy, dx, 1, 0, d_out, nullptr);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
            double max_err = 0;
            int bad = 0;
            for (int m = 0; m < 128; m++)
                for (int n = 0; n < 256; n++) {
                    int y = n / 32, b = (n / 8) % 4, x = n % 8;
                    int yy = y + dy, xx = x + dx;
                    double ref = 0;
                    if (yy >= 0 && yy < 8 && xx >= 0 && xx < 8) {
                        int src = yy * 32 + b * 8 + xx;
                        for (int k = 0; k < 64; k++) ref += double(W[m * 64 + k]) * X[src * 64 + k];
                    }
                    double err = fabs(ref - out[m * 256 + n]);
                    if (err > max_err) max_err = err;
                    if (err > 1e-2) bad++;
                }
            printf("tap dy %+d dx %+d: max |err| %.3g, mismatches %d  (%s)\n", dy, dx, max_err, bad, cudaGetErrorString(e));
            bad_total += bad;
        }
    const int iters = 1024;
    for (int sw : {1, 0})
        for (int dx : {0, 1, -1}) {
            if (sw && dx) continue;
            bench<<<148, 128, 210 * 1024>>>(d_img, smem_bytes, 0, dx, iters, sw, nullptr, d_cyc);
            cudaError_t e = cudaDeviceSynchronize();
            unsigned long long c[2]; cudaMemcpy(c, d_cyc, 16, cudaMemcpyDeviceToHost);
            printf("%s B operand, dx %+d: %.1f cycles / MMA (M128 N256 K16; ideal 128)  %s\n", sw ? "SWIZZLE_128B" : "no-swizzle 144B-pitch", dx,
                   double(c[0]) / (iters * 4), cudaGetErrorString(e));
        }
    for (int dy : {-1, 1, 9}) {
        bench<<<148, 128, 210 * 1024>>>(d_img, smem_bytes, dy, 0, iters, 0, nullptr, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long c[2]; cudaMemcpy(c, d_cyc, 16, cudaMemcpyDeviceToHost);
        printf("no-swizzle, dy %+d%s: %.1f cycles / MMA  %s\n", dy, dy == 9 ? " (= all nine taps in turn)" : "", double(c[0]) / (iters * 4),
               cudaGetErrorString(e));
    }
    printf("%s\n", bad_total ? "FAILED" : "ALL TAPS CORRECT");
    return 0;
}
