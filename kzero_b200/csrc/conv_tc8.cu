// K1, 8x8-board specialisation ("mode 2"): conv3x3 as implicit GEMM with the activation tile loaded ONCE
// per horizontal tap and re-used for the three vertical taps, and the weight tile shared by two M tiles.
//
// Why: the generic kernel (conv_tc.cu) re-loads a shifted activation tile for each of the 9 taps and a weight
// tile per M tile -- 576 KB of L2->SM traffic per 128 output rows and layer.  ncu on B200 (profiles/) shows
// that kernel pinned by L2->SM bandwidth (9.7 TB/s, tensor pipe 38 % active), not by the tensor cores.
// This kernel moves 264 KB per 128 rows instead:
//
//   * one work unit = 4 boards = 256 output rows = two M=128 accumulators
//   * the TMA box for horizontal tap dx is (64 ch, 8 x, 4 boards, 10 y) at (c0, dx, b0, -1) through a tensor map
//     whose dimensions are ordered (c, x, board, y): in shared memory the 1 KiB swizzle atoms (8 rows = the 8
//     x-positions of one rank) end up ordered [y = -1..8][board 0..3].  x = -1 / 8 and y = -1 / 8 are outside
//     the tensor, so the TMA engine zero-fills them: that IS the conv's zero padding.
//   * a vertical tap dy is then nothing but a different start address for the UMMA descriptor:
//     M tile t (ranks 4t..4t+3 of the 4 boards) with tap dy reads the 16 consecutive atoms starting at
//     atom (4t + 1 + dy) * 4 -- always 1 KiB aligned, canonical K-major SWIZZLE_128B.
//   * per (k-block, dx) the three (dy) weight tiles are streamed through their own ring; each is used by
//     both M tiles before it is released.
//   * TMEM: 2 tiles x 2 buffers x N columns (N <= 128): the epilogue of unit i overlaps the MMAs of unit i+1.
//   * accumulator row m of tile t  <->  rank y = 4t + m/32, board = (m/8)%4, x = m%8.
#include "kernels.cuh"
#include "tc_common.cuh"

namespace kzb {
namespace {

using namespace tc;

constexpr int kThreads = 192;
constexpr int kBoards = 4;                         // boards per work unit
constexpr int kASlotBytes = 10 * kBoards * 1024;   // (8 + 2 halo ranks) x 4 boards x 1 KiB atom
constexpr int kASlots = 2;

struct Smem8 {
    uint8_t* a;  // kASlots x kASlotBytes
    uint8_t* b;  // b_slots x (n * 128)
    uint64_t *a_full, *a_empty, *b_full, *b_empty, *tmem_full, *tmem_empty;
    uint32_t* tmem_ptr;
    float* bias;
};

__device__ __forceinline__ Smem8 carve8(uint8_t* base, int n, int b_slots) {
    Smem8 s;
    s.a = base;
    s.b = base + kASlots * kASlotBytes;
    uint8_t* p = s.b + size_t(b_slots) * n * 128;
    s.a_full = reinterpret_cast<uint64_t*>(p);
    s.a_empty = s.a_full + kASlots;
    s.b_full = s.a_empty + kASlots;
    s.b_empty = s.b_full + b_slots;
    s.tmem_full = s.b_empty + b_slots;
    s.tmem_empty = s.tmem_full + 2;
    s.tmem_ptr = reinterpret_cast<uint32_t*>(s.tmem_empty + 2);
    s.bias = reinterpret_cast<float*>(s.tmem_ptr + 4);
    return s;
}

__global__ void __launch_bounds__(kThreads, 1)
    conv_tc8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const ConvTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_slots = p.stages;
    const Smem8 sm = carve8(smem, p.n, b_slots);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int b_bytes = p.n * 128;
    const int num_units = p.num_tiles;  // here: 4-board units
    unsigned long long* tl = p.timeline ? p.timeline + size_t(blockIdx.x) * 1024 : nullptr;
    if (tl && threadIdx.x == 0) tl[0] = clock64();

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
        for (int i = 0; i < kASlots; i++) {
            mbar_init(&sm.a_full[i], 1);
            mbar_init(&sm.a_empty[i], 1);
        }
        for (int i = 0; i < b_slots; i++) {
            mbar_init(&sm.b_full[i], 1);
            mbar_init(&sm.b_empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&sm.tmem_full[i], 1);
            mbar_init(&sm.tmem_empty[i], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.tmem_ptr)),
                     "r"(uint32_t(p.tmem_cols))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.n; i += kThreads) sm.bias[i] = p.bias[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_ptr;
    const uint32_t acc_cols = uint32_t(p.tmem_cols / 4);  // columns per (buffer, tile) accumulator
    if (tl && threadIdx.x == 0) tl[1] = clock64();

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int a_slot = 0, b_slot = 0;
            uint32_t a_phase = 0, b_phase = 0;
            for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
                for (int kb = 0; kb < p.kblocks; kb++) {
                    for (int dx = -1; dx <= 1; dx++) {
                        mbar_wait(&sm.a_empty[a_slot], a_phase ^ 1);
                        mbar_expect_tx(&sm.a_full[a_slot], kASlotBytes);
                        tma_load_4d(&tmap_a, &sm.a_full[a_slot], sm.a + size_t(a_slot) * kASlotBytes, kb * 64, dx,
                                    unit * kBoards, -1);
                        if (++a_slot == kASlots) {
                            a_slot = 0;
                            a_phase ^= 1;
                        }
                        for (int dy = -1; dy <= 1; dy++) {
                            const int tap = (dy + 1) * 3 + (dx + 1);
                            mbar_wait(&sm.b_empty[b_slot], b_phase ^ 1);
                            mbar_expect_tx(&sm.b_full[b_slot], uint32_t(b_bytes));
                            tma_load_2d(&tmap_b, &sm.b_full[b_slot], sm.b + size_t(b_slot) * b_bytes,
                                        tap * p.cin_pad + kb * 64, 0);
                            if (++b_slot == b_slots) {
                                b_slot = 0;
                                b_phase ^= 1;
                            }
                        }
                    }
                }
                if (tl && unit == blockIdx.x) tl[2] = clock64();  // all TMA loads of the first unit issued
            }
            if (tl) tl[3] = clock64();  // all TMA loads issued
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // whole warp walks the loops (uniform datapath for the address arithmetic); lane 0 issues MMAs + commits
        const uint32_t idesc = umma_idesc_bf16(128, p.n);
        const uint64_t desc_hi = umma_desc_sw128_hi();
        int a_slot = 0, b_slot = 0;
        uint32_t a_phase = 0, b_phase = 0;
        int local = 0;
        for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, local++) {
            const int buf = local & 1;
            mbar_wait(&sm.tmem_empty[buf], ((local >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + uint32_t(buf * 2) * acc_cols;
            bool first = true;
            for (int kb = 0; kb < p.kblocks; kb++) {
                for (int dx = -1; dx <= 1; dx++) {
                    mbar_wait(&sm.a_full[a_slot], a_phase);
                    if (tl && first && local < 2 && lane == 0) tl[4 + local] = clock64();  // first operands of the unit landed
                    const uint32_t a_lo = umma_desc_lo(smem_u32(sm.a + size_t(a_slot) * kASlotBytes));
                    for (int dy = -1; dy <= 1; dy++) {
                        mbar_wait(&sm.b_full[b_slot], b_phase);
                        tc_fence_after();
                        const uint32_t b_lo = umma_desc_lo(smem_u32(sm.b + size_t(b_slot) * b_bytes));
                        if (lane == 0) {
#pragma unroll
                            for (int t = 0; t < 2; t++) {
                                const uint32_t a_t = a_lo + uint32_t((4 * t + 1 + dy) * kBoards * (1024 >> 4));
#pragma unroll
                                for (int k = 0; k < 4; k++) {
                                    umma_bf16(tmem_d + uint32_t(t) * acc_cols, desc_hi | uint64_t(a_t + 2 * k),
                                              desc_hi | uint64_t(b_lo + 2 * k), idesc, (!first || k != 0) ? 1u : 0u);
                                }
                            }
                            umma_commit(&sm.b_empty[b_slot]);
                            if (dy == 1) umma_commit(&sm.a_empty[a_slot]);
                        }
                        __syncwarp();
                        first = false;
                        if (++b_slot == b_slots) {
                            b_slot = 0;
                            b_phase ^= 1;
                        }
                    }
                    if (++a_slot == kASlots) {
                        a_slot = 0;
                        a_phase ^= 1;
                    }
                }
            }
            if (lane == 0) {
                umma_commit(&sm.tmem_full[buf]);
                if (tl && local < 2) tl[6 + local] = clock64();  // all MMAs of the unit issued
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int quarter = warp % 4;
        int local = 0;
        for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, local++) {
            const int buf = local & 1;
            const int board = unit * kBoards + lane / 8;
            // tile t, TMEM lane quarter*32+lane  <->  rank 4t+quarter, board lane/8, file lane%8
            const int row0 = board * 64 + quarter * 8 + (lane % 8);
            const int row1 = row0 + 32;
            const bool store0 = row0 < p.valid_rows, store1 = row1 < p.valid_rows;
            const bool has_res = p.res != nullptr;

            constexpr int kResVec = 16;  // 16 x 8 channels = 128 = max N of this kernel
            uint4 res0[kResVec], res1[kResVec];
            if (has_res && store0) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.res + size_t(row0) * p.res_stride);
#pragma unroll
                for (int j = 0; j < kResVec; j++)
                    if (j * 8 < p.n_store) res0[j] = rp[j];
            }
            mbar_wait(&sm.tmem_full[buf], (local >> 1) & 1);
            tc_fence_after();
            if (tl && warp == 2 && lane == 0 && local < 2) tl[8 + local] = clock64();  // accumulators of the unit complete
            if (has_res && store1) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.res + size_t(row1) * p.res_stride);
#pragma unroll
                for (int j = 0; j < kResVec; j++)
                    if (j * 8 < p.n_store) res1[j] = rp[j];
            }

#pragma unroll
            for (int t = 0; t < 2; t++) {
                const int row = t == 0 ? row0 : row1;
                const bool store = t == 0 ? store0 : store1;
                const uint32_t taddr = tmem_base + uint32_t(buf * 2 + t) * acc_cols + (uint32_t(quarter * 32) << 16);
#pragma unroll
                for (int cc = 0; cc < 4; cc++) {  // 32 columns per iteration, n <= 128
                    const int c0 = cc * 32;
                    if (c0 >= p.n_store) break;
                    const bool second = c0 + 16 < p.n_store;
                    uint32_t r[32];
                    tmem_ld16(taddr + c0, r);
                    if (second) tmem_ld16(taddr + c0 + 16, r + 16);
                    tmem_ld_wait();
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        if (h == 1 && !second) break;
                        const int ch = c0 + h * 16;
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            float f = __uint_as_float(r[h * 16 + j]) + sm.bias[ch + j];
                            if (ch + j < p.relu_n) f = f < 0.0f ? 0.0f : f;
                            v[j] = f;
                        }
                        if (has_res && store) {
                            const uint4 q0 = t == 0 ? res0[cc * 4 + h * 2] : res1[cc * 4 + h * 2];
                            const uint4 q1 = t == 0 ? res0[cc * 4 + h * 2 + 1] : res1[cc * 4 + h * 2 + 1];
                            const __nv_bfloat16* h0 = reinterpret_cast<const __nv_bfloat16*>(&q0);
                            const __nv_bfloat16* h1 = reinterpret_cast<const __nv_bfloat16*>(&q1);
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                v[j] += __bfloat162float(h0[j]);
                                v[8 + j] += __bfloat162float(h1[j]);
                            }
                        }
                        if (store) {
                            uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + size_t(row) * p.out_stride + ch);
                            op[0] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                            op[1] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.tmem_empty[buf]);
            if (tl && warp == 2 && lane == 0 && local < 2) tl[10 + local] = clock64();  // epilogue of the unit done
        }
    }

    tc_fence_before();
    __syncthreads();
    if (tl && threadIdx.x == 0) tl[12] = clock64();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(p.tmem_cols))
                     : "memory");
    }
}

}  // namespace

size_t conv_tc8_smem_bytes(int n, int b_slots) {
    return 1024 + size_t(kASlots) * kASlotBytes + size_t(b_slots) * n * 128 + (2 * kASlots + 2 * b_slots + 4) * 8 + 16 + size_t(n) * 4;
}

int conv_tc8_pick_b_slots(int n) {
    int slots = 12;
    while (slots > 3 && conv_tc8_smem_bytes(n, slots) > 227 * 1024) slots--;
    return slots;
}

void conv_tc8_prepare() { cudaFuncSetAttribute(conv_tc8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }

void launch_conv_tc8(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const ConvTcParams& p, int grid, cudaStream_t s) {
    if (p.num_tiles <= 0) return;
    conv_tc8_kernel<<<std::min(grid, p.num_tiles), kThreads, conv_tc8_smem_bytes(p.n, p.stages), s>>>(tmap_a, tmap_b, p);
}

}  // namespace kzb
