// K1, whole-tower persistent kernel for 8x8 boards: every conv3x3 of the ResNet tower in ONE launch.
//
// A 3x3 conv never looks outside its own board, so a CTA that owns a set of boards can run them through all
// 2*D+1 layers without ever synchronising with another CTA.  Each CTA therefore keeps its 4-board work units
// for the whole tower and walks layer by layer; activations round-trip through global memory (they stay in
// the 126 MB L2: a chess batch of 1024 is 16.8 MB per tensor), weights are streamed from L2 by TMA.
// What this removes compared with one launch per layer (conv_tc8.cu, measured with clock64 stamps on B200):
// the ~1.6k-cycle prologue (TMEM alloc, barrier init), the ~2.4k-cycle first-operand latency and the fully
// exposed ~10k-cycle last epilogue of EVERY layer, plus the launch gaps.  Here the epilogue of (layer L, unit u)
// overlaps the MMAs of the next work item and the tensor pipe only drains once, at the end of the tower.
//
// Operand staging is the scheme of conv_tc8.cu: per (k-block, dx) ONE TMA box (64 ch, 8 x, 4 boards, 8 ranks)
// through the (c, x, board, y)-ordered tensor map; x = -1 / 8 zero-filled by the TMA engine; the y = -1 / 8 halo
// ranks are permanent zero atoms sitting between the A slots in shared memory, so a vertical tap dy is just a
// UMMA descriptor start address (always 1 KiB aligned).  Three dy weight tiles per A tile, shared by both
// M=128 accumulators of the unit.
//
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner), warps 2..5 epilogue
// (tcgen05.ld -> +bias -> relu -> +residual -> bf16 -> 256-bit global stores).
// Cross-layer dependency: the epilogue warps arrive on ready[unit] after their stores (threadfence + async-proxy
// fence); the producer waits on it before it lets the TMA engine read that unit's rows for the next layer.
#include "kernels.cuh"
#include "tc_common.cuh"

namespace kzb {
namespace {

using namespace tc;

constexpr int kThreads = 192;
constexpr int kBoards = 4;
constexpr int kAtom = 1024;
constexpr int kRankBytes = kBoards * kAtom;   // one rank of 4 boards
constexpr int kABox = 8 * kRankBytes;         // 32 KiB per TMA box
constexpr int kASlots = 3;
constexpr int kAStride = kABox + kRankBytes;  // slot + the zero rank that follows it
constexpr int kARegion = kRankBytes + kASlots * kAStride;
constexpr int kMaxLocalUnits = 16;

struct SmemT {
    uint8_t* a;  // [Z][A0][Z][A1][Z][A2][Z]
    uint8_t* b;
    uint64_t *a_full, *a_empty, *b_full, *b_empty, *tmem_full, *tmem_empty, *ready;
    uint32_t* tmem_ptr;
    float* bias;  // [2][128]
};

__device__ __forceinline__ SmemT carve_t(uint8_t* base, int n, int b_slots) {
    SmemT s;
    s.a = base;
    s.b = base + kARegion;
    uint8_t* p = s.b + size_t(b_slots) * n * 128;
    s.a_full = reinterpret_cast<uint64_t*>(p);
    s.a_empty = s.a_full + kASlots;
    s.b_full = s.a_empty + kASlots;
    s.b_empty = s.b_full + b_slots;
    s.tmem_full = s.b_empty + b_slots;
    s.tmem_empty = s.tmem_full + 2;
    s.ready = s.tmem_empty + 2;
    s.tmem_ptr = reinterpret_cast<uint32_t*>(s.ready + kMaxLocalUnits);
    s.bias = reinterpret_cast<float*>(s.tmem_ptr + 4);
    return s;
}

__device__ __forceinline__ void ldg256(const void* ptr, uint32_t* r) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(ptr));
}
__device__ __forceinline__ void stg256(void* ptr, const uint32_t* r) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

__global__ void __launch_bounds__(kThreads, 1)
    tower8_kernel(const __grid_constant__ Tower8Maps maps, const Tower8Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const SmemT sm = carve_t(smem, p.n, p.b_slots);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int b_bytes = p.n * 128;
    // development aid (KZB_TIMELINE=tower8): 1024 clock64() stamps per CTA, 8 per work item:
    //   [0] tmem_empty acquired  [1] first operands landed  [2] MMAs issued  [3] accumulators complete
    //   [4] epilogue done        [5] producer: ready acquired  [6] producer: unit's loads issued
    unsigned long long* tl = p.timeline ? p.timeline + size_t(blockIdx.x) * 1024 : nullptr;
#define KZB_STAMP(item, k) do { if (tl && (item) < 127) tl[8 + (item) * 8 + (k)] = clock64(); } while (0)
    if (tl && threadIdx.x == 0) tl[0] = clock64();

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 3; i++) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.a[i])) : "memory");
        for (int i = 0; i < 2; i++) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w[i])) : "memory");
        for (int i = 0; i < kASlots; i++) {
            mbar_init(&sm.a_full[i], 1);
            mbar_init(&sm.a_empty[i], 1);
        }
        for (int i = 0; i < p.b_slots; i++) {
            mbar_init(&sm.b_full[i], 1);
            mbar_init(&sm.b_empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&sm.tmem_full[i], 1);
            mbar_init(&sm.tmem_empty[i], 4);
        }
        for (int i = 0; i < kMaxLocalUnits; i++) mbar_init(&sm.ready[i], 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.tmem_ptr)),
                     "r"(uint32_t(p.tmem_cols))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // permanent zero ranks (the conv's vertical zero padding) around the A slots
    for (int z = 0; z <= kASlots; z++) {
        uint4* zp = reinterpret_cast<uint4*>(sm.a + size_t(z) * kAStride);
        for (int i = threadIdx.x; i < kRankBytes / 16; i += kThreads) zp[i] = make_uint4(0, 0, 0, 0);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zeros -> visible to UMMA reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_ptr;
    const uint32_t acc_cols = uint32_t(p.tmem_cols / 4);
    if (tl && threadIdx.x == 0) tl[1] = clock64();

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int a_slot = 0, b_slot = 0;
            uint32_t a_phase = 0, b_phase = 0;
            int pitem = 0;
            for (int L = 0; L < p.num_layers; L++) {
                const TowerLayerDev ld = p.layers[L];
                const CUtensorMap* amap = &maps.a[ld.a_map];
                const CUtensorMap* wmap = &maps.w[ld.w_map];
                int ul = 0;
                for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ul++, pitem++) {
                    // rows of this unit written by layer L-1's epilogue must be complete and visible
                    if (L > 0) mbar_wait(&sm.ready[ul], uint32_t(L - 1) & 1);
                    KZB_STAMP(pitem, 5);
                    for (int kb = 0; kb < ld.kblocks; kb++) {
                        for (int dx = -1; dx <= 1; dx++) {
                            mbar_wait(&sm.a_empty[a_slot], a_phase ^ 1);
                            if (p.debug & 1) {
                                mbar_arrive(&sm.a_full[a_slot]);
                            } else {
                                mbar_expect_tx(&sm.a_full[a_slot], kABox);
                                tma_load_4d(amap, &sm.a_full[a_slot], sm.a + kRankBytes + size_t(a_slot) * kAStride, kb * 64, dx,
                                            unit * kBoards, 0);
                            }
                            if (++a_slot == kASlots) {
                                a_slot = 0;
                                a_phase ^= 1;
                            }
                            for (int dy = -1; dy <= 1; dy++) {
                                const int tap = (dy + 1) * 3 + (dx + 1);
                                mbar_wait(&sm.b_empty[b_slot], b_phase ^ 1);
                                if (p.debug & 2) {
                                    mbar_arrive(&sm.b_full[b_slot]);
                                } else {
                                    mbar_expect_tx(&sm.b_full[b_slot], uint32_t(b_bytes));
                                    tma_load_2d(wmap, &sm.b_full[b_slot], sm.b + size_t(b_slot) * b_bytes,
                                                tap * ld.cin_pad + kb * 64, ld.w_row0);
                                }
                                if (++b_slot == p.b_slots) {
                                    b_slot = 0;
                                    b_phase ^= 1;
                                }
                            }
                        }
                    }
                    KZB_STAMP(pitem, 6);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // The whole warp walks the loops (warp-uniform control flow keeps the address arithmetic on the uniform
        // datapath); lane 0 alone issues tcgen05.mma and the commits that track them.
        const uint32_t idesc = umma_idesc_bf16(128, p.n);
        const uint64_t desc_hi = umma_desc_sw128_hi();
        int a_slot = 0, b_slot = 0;
        uint32_t a_phase = 0, b_phase = 0;
        int item = 0;
        for (int L = 0; L < p.num_layers; L++) {
            const int kblocks = p.layers[L].kblocks;
            for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, item++) {
                const int buf = item & 1;
                mbar_wait(&sm.tmem_empty[buf], ((item >> 1) & 1) ^ 1);
                tc_fence_after();
                if (lane == 0) KZB_STAMP(item, 0);
                const uint32_t tmem_d = tmem_base + uint32_t(buf * 2) * acc_cols;
                bool first = true;
                for (int kb = 0; kb < kblocks; kb++) {
                    for (int dx = -1; dx <= 1; dx++) {
                        mbar_wait(&sm.a_full[a_slot], a_phase);
                        if (first && lane == 0) KZB_STAMP(item, 1);
                        const uint32_t a_lo = umma_desc_lo(smem_u32(sm.a + kRankBytes + size_t(a_slot) * kAStride));
                        for (int dy = -1; dy <= 1; dy++) {
                            mbar_wait(&sm.b_full[b_slot], b_phase);
                            tc_fence_after();
                            const uint32_t b_lo = umma_desc_lo(smem_u32(sm.b + size_t(b_slot) * b_bytes));
                            if (lane == 0) {
#pragma unroll
                                for (int t = 0; t < 2; t++) {
                                    // tile t = ranks 4t..4t+3; tap dy starts dy ranks away (rank -1 / 8 = zero atoms)
                                    const uint32_t a_t = a_lo + uint32_t((4 * t + dy) * (kRankBytes >> 4));
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
                            if (++b_slot == p.b_slots) {
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
                    KZB_STAMP(item, 2);
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int quarter = warp % 4;
        const int et = threadIdx.x - 64;  // 0..127
        int item = 0;
        for (int L = 0; L < p.num_layers; L++) {
            const TowerLayerDev ld = p.layers[L];
            float* bias = sm.bias + (L & 1) * 128;
            if (et < p.n) bias[et] = ld.bias[et];
            asm volatile("bar.sync 1, 128;" ::: "memory");
            __nv_bfloat16* out = ld.out_buf == 1 ? p.x : p.t;
            int ul = 0;
            for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, item++, ul++) {
                const int buf = item & 1;
                const int board = unit * kBoards + lane / 8;
                // tile t, TMEM lane quarter*32+lane  <->  rank 4t+quarter, board lane/8, file lane%8
                const int row0 = board * 64 + quarter * 8 + (lane % 8);
                const int row1 = row0 + 32;
                const bool live = !(p.debug & 4);
                const bool store0 = live && row0 < p.valid_rows, store1 = live && row1 < p.valid_rows;

                // residual rows (bf16) fetched before the accumulator is waited for: latency hides behind the MMAs
                uint32_t res[2][64];
                if (ld.has_res) {
#pragma unroll
                    for (int t = 0; t < 2; t++) {
                        const bool st = t == 0 ? store0 : store1;
                        const __nv_bfloat16* rp = p.x + size_t(t == 0 ? row0 : row1) * p.stride;
                        if (st) {
#pragma unroll
                            for (int j = 0; j < 8; j++)
                                if (j * 16 < p.n_store) ldg256(rp + j * 16, &res[t][j * 8]);
                        }
                    }
                }
                mbar_wait(&sm.tmem_full[buf], (item >> 1) & 1);
                tc_fence_after();
                if (warp == 2 && lane == 0) KZB_STAMP(item, 3);

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
                                float f = __uint_as_float(r[h * 16 + j]) + bias[ch + j];
                                if (ch + j < ld.relu_n) f = f < 0.0f ? 0.0f : f;
                                v[j] = f;
                            }
                            if (ld.has_res && store) {
#pragma unroll
                                for (int j = 0; j < 8; j++) {
                                    const uint32_t q = res[t][(cc * 2 + h) * 8 + j];
                                    v[2 * j] += bf16_lo(q);
                                    v[2 * j + 1] += bf16_hi(q);
                                }
                            }
                            if (store) {
                                uint32_t o[8];
#pragma unroll
                                for (int j = 0; j < 8; j++) o[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
                                stg256(out + size_t(row) * p.stride + ch, o);
                            }
                        }
                    }
                }
                tc_fence_before();
                // make this unit's rows visible to the TMA engine (async proxy) before the next layer may load them
                __threadfence();
                asm volatile("fence.proxy.async;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&sm.tmem_empty[buf]);
                    mbar_arrive(&sm.ready[ul]);
                    if (warp == 2) KZB_STAMP(item, 4);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (tl && threadIdx.x == 0) tl[2] = clock64();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(p.tmem_cols))
                     : "memory");
    }
}

}  // namespace

size_t tower8_smem_bytes(int n, int b_slots) {
    return 1024 + size_t(kARegion) + size_t(b_slots) * n * 128 + (2 * kASlots + 2 * b_slots + 4 + kMaxLocalUnits) * 8 + 16 + 2 * 128 * 4;
}

int tower8_pick_b_slots(int n) {
    int slots = 12;
    while (slots > 3 && tower8_smem_bytes(n, slots) > 227 * 1024) slots--;
    return slots;
}

int tower8_max_local_units() { return kMaxLocalUnits; }

void tower8_prepare() { cudaFuncSetAttribute(tower8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }

void launch_tower8(const Tower8Maps& maps, const Tower8Params& p, int grid, cudaStream_t s) {
    if (p.num_units <= 0 || p.num_layers <= 0) return;
    tower8_kernel<<<std::min(grid, p.num_units), kThreads, tower8_smem_bytes(p.n, p.b_slots), s>>>(maps, p);
}

}  // namespace kzb
