// K1, whole-tower persistent kernel for 8x8 boards: every conv3x3 of the ResNet tower in ONE launch.
//
// A 3x3 conv never looks outside its own board, so a CTA that owns a set of boards can run them through all
// 2*D+1 layers without ever synchronising with another CTA.  Each CTA keeps its 4-board work units for the
// whole tower and walks layer by layer; activations round-trip through global memory (they stay in the
// 126 MB L2: a chess batch of 1024 is 16.8 MB per tensor), weights are streamed from L2 by TMA.
//
// GEMM orientation (per layer, per unit):  D^T[M = 128 out-channels, N = 256 positions] += W_tap[M, K] * X_tap[N, K]
//   * the WEIGHT tile is the UMMA "A" operand (M = 128 rows, K-major), the ACTIVATIONS are the "B" operand with
//     N = 256 = the unit's 4 boards x 64 squares.  One tcgen05.mma (M128 N256 K16) is 128 tensor cycles; measured on
//     B200 (scripts/micro/mma_bench.cu) a single issuing thread sustains one MMA per ~94 cycles, so N = 128 MMAs
//     (64 cycles) are issue-bound at ~66 % while N = 256 MMAs run the tensor pipe back to back.
//   * activations are staged by ONE TMA box per (k-block, dx): (64 ch, 8 x, 4 boards, 8 ranks) through a tensor map
//     whose dims are ordered (c, x, board, y); x = -1 / 8 is zero-filled by the TMA engine; the y = -1 / 8 halo
//     ranks are permanent zero atoms between the activation slots in shared memory.  In smem the 1 KiB swizzle
//     atoms (8 x-positions x 64 ch) are ordered [rank][board], so the vertical tap dy is only a different UMMA
//     descriptor start address (slot + dy ranks, always 1 KiB aligned) -- no data movement for 6 of the 9 taps.
//   * accumulator: TMEM lane = output channel, column n = rank*32 + board*8 + file.  2 buffers x 256 columns.
//
// Epilogue (warps 2..5, thread = one output channel): tcgen05.ld 32 columns (= one rank of the 4 boards) ->
//   +bias -> relu -> +residual -> bf16 -> transposed through a per-warp shared-memory staging tile -> TMA store
//   (NHWC rows, 32 channels x one rank of the 4 boards per store).
//   The reference block is x + relu(bn(conv(...))): relu BEFORE the add (python/lib/model/post_act.py:218-228).
//   The residual stream is additionally kept channel-major (XT[unit][rank][half][channel][16 positions], bf16,
//   bit-identical to X) so that the thread owning a channel reads / writes its residual with 32-byte accesses that
//   are 1 KiB-contiguous across the warp.
// Cross-layer dependency: after the unit's TMA stores have completed the epilogue arrives on ready[unit]; the
//   producer waits on it before it lets the TMA engine read those rows for the next layer.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner), warps 2..9 epilogue.
#include "kernels.cuh"
#include "tc_common.cuh"

namespace kzb {
namespace {

using namespace tc;

constexpr int kEpiWarps = 8;                   // 2 per TMEM lane quarter (16 measured no faster: the epilogue is bound by L2 traffic, not by latency)
constexpr int kChunks = 32 / kEpiWarps;        // ranks (32-position chunks) per epilogue warp and unit
constexpr int kStageBufs = kEpiWarps == 8 ? 2 : 1;
constexpr int kThreads = 64 + 32 * kEpiWarps;  // warp 0 TMA, warp 1 MMA, the rest epilogue
constexpr int kBoards = 4;
constexpr int kAtom = 1024;
constexpr int kRankBytes = kBoards * kAtom;   // one rank of 4 boards, 64 channels
constexpr int kXBox = 8 * kRankBytes;         // 32 KiB per activation TMA box
constexpr int kXSlots = 3;
constexpr int kXStride = kXBox + kRankBytes;  // slot + the zero rank that follows it
constexpr int kXRegion = kRankBytes + kXSlots * kXStride;
constexpr int kWBytes = 128 * 128;            // weight tile: 128 out-channels x 64 k, 16 KiB
constexpr int kStageBytes = kEpiWarps * kStageBufs * 32 * 64;  // output staging: per warp [32 positions][32 ch] bf16 tiles
constexpr int kMaxLocalUnits = 16;

struct SmemT {
    uint8_t* x;      // [Z][X0][Z][X1][Z][X2][Z]
    uint8_t* w;      // w_slots x 16 KiB
    uint8_t* stage;  // 2 x 8 KiB
    uint64_t *x_full, *x_empty, *w_full, *w_empty, *tmem_full, *tmem_empty, *ready;
    uint32_t* tmem_ptr;
};

__device__ __forceinline__ SmemT carve_t(uint8_t* base, int w_slots) {
    SmemT s;
    s.x = base;
    s.w = base + kXRegion;
    s.stage = s.w + size_t(w_slots) * kWBytes;
    uint8_t* p = s.stage + kStageBytes;
    s.x_full = reinterpret_cast<uint64_t*>(p);
    s.x_empty = s.x_full + kXSlots;
    s.w_full = s.x_empty + kXSlots;
    s.w_empty = s.w_full + w_slots;
    s.tmem_full = s.w_empty + w_slots;
    s.tmem_empty = s.tmem_full + 2;
    s.ready = s.tmem_empty + 2;
    s.tmem_ptr = reinterpret_cast<uint32_t*>(s.ready + kMaxLocalUnits);
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

// CL = CTAs per cluster.  CL = 2: the two CTAs of a pair walk the same (layer, item) sequence on different boards and
// share every weight tile: each loads half of it (n/2 output channels) and TMA-multicasts it into both CTAs' rings, which
// halves the weight traffic out of L2 (the kernel is otherwise pinned by L2->SM bandwidth, profiles/).
template <int CL>
__global__ void __launch_bounds__(kThreads, 1)
    tower8_kernel(const __grid_constant__ Tower8Maps maps, const Tower8Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const SmemT sm = carve_t(smem, p.b_slots);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    // development aid (KZB_TIMELINE=tower8): 1024 clock64() stamps per CTA, 8 per work item:
    //   [0] tmem_empty acquired  [1] first operands landed  [2] MMAs issued  [3] accumulators complete
    //   [4] epilogue done        [5] producer: ready acquired  [6] producer: unit's loads issued
    unsigned long long* tl = p.timeline ? p.timeline + size_t(blockIdx.x) * 1024 : nullptr;
#define KZB_STAMP(item, k) do { if (tl && (item) < 127) tl[8 + (item) * 8 + (k)] = clock64(); } while (0)
    if (tl && threadIdx.x == 0) tl[0] = clock64();

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 3; i++) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.a[i])) : "memory");
        for (int i = 0; i < 2; i++) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w[i])) : "memory");
        for (int i = 0; i < 2; i++) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.out[i])) : "memory");
        for (int i = 0; i < kXSlots; i++) {
            mbar_init(&sm.x_full[i], 1);
            mbar_init(&sm.x_empty[i], 1);
        }
        for (int i = 0; i < p.b_slots; i++) {
            mbar_init(&sm.w_full[i], 1);
            mbar_init(&sm.w_empty[i], CL);  // released by the MMA warps of every CTA the tile was multicast to
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&sm.tmem_full[i], 1);
            mbar_init(&sm.tmem_empty[i], kEpiWarps);
        }
        for (int i = 0; i < kMaxLocalUnits; i++) mbar_init(&sm.ready[i], kEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.tmem_ptr)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // permanent zero ranks (the conv's vertical zero padding) around the activation slots; the weight ring is
    // cleared too so that rows >= n of a narrow net (n < 128) never feed NaN bit patterns to the tensor core
    for (int z = 0; z <= kXSlots; z++) {
        uint4* zp = reinterpret_cast<uint4*>(sm.x + size_t(z) * kXStride);
        for (int i = threadIdx.x; i < kRankBytes / 16; i += kThreads) zp[i] = make_uint4(0, 0, 0, 0);
    }
    for (int i = threadIdx.x; i < p.b_slots * kWBytes / 16; i += kThreads) reinterpret_cast<uint4*>(sm.w)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zeros -> visible to UMMA reads
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast / committed to them
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_ptr;
    const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0;
    constexpr uint16_t kMask = uint16_t((1u << CL) - 1);
    if (tl && threadIdx.x == 0) tl[1] = clock64();

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            const uint32_t w_bytes = uint32_t(p.n) * 128u;
            int x_slot = 0, w_slot = 0;
            uint32_t x_phase = 0, w_phase = 0;
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
                            mbar_wait(&sm.x_empty[x_slot], x_phase ^ 1);
                            if (p.debug & 1) {
                                mbar_arrive(&sm.x_full[x_slot]);
                            } else {
                                mbar_expect_tx(&sm.x_full[x_slot], kXBox);
                                tma_load_4d(amap, &sm.x_full[x_slot], sm.x + kRankBytes + size_t(x_slot) * kXStride, kb * 64, dx,
                                            unit * kBoards, 0);
                            }
                            if (++x_slot == kXSlots) {
                                x_slot = 0;
                                x_phase ^= 1;
                            }
                            for (int dy = -1; dy <= 1; dy++) {
                                const int tap = (dy + 1) * 3 + (dx + 1);
                                mbar_wait(&sm.w_empty[w_slot], w_phase ^ 1);
                                if (p.debug & 2) {
                                    mbar_arrive(&sm.w_full[w_slot]);
                                } else {
                                    mbar_expect_tx(&sm.w_full[w_slot], w_bytes);
                                    if (CL == 1) {
                                        tma_load_2d(wmap, &sm.w_full[w_slot], sm.w + size_t(w_slot) * kWBytes,
                                                    tap * ld.cin_pad + kb * 64, ld.w_row0);
                                    } else {  // my n/CL rows of the tile, into every CTA of the cluster
                                        const int rows = p.n / CL;
                                        tma_load_2d_multicast(wmap, &sm.w_full[w_slot],
                                                              sm.w + size_t(w_slot) * kWBytes + size_t(cta_rank) * rows * 128,
                                                              tap * ld.cin_pad + kb * 64, ld.w_row0 + int(cta_rank) * rows, kMask);
                                    }
                                }
                                if (++w_slot == p.b_slots) {
                                    w_slot = 0;
                                    w_phase ^= 1;
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
        const uint32_t idesc = umma_idesc_bf16(128, 256);
        const uint64_t desc_hi = umma_desc_sw128_hi();
        int x_slot = 0, w_slot = 0;
        uint32_t x_phase = 0, w_phase = 0;
        int item = 0;
        for (int L = 0; L < p.num_layers; L++) {
            const int kblocks = p.layers[L].kblocks;
            for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, item++) {
                const int buf = item & 1;
                mbar_wait(&sm.tmem_empty[buf], ((item >> 1) & 1) ^ 1);
                tc_fence_after();
                if (lane == 0) KZB_STAMP(item, 0);
                const uint32_t tmem_d = tmem_base + uint32_t(buf) * 256u;
                bool first = true;
                for (int kb = 0; kb < kblocks; kb++) {
                    for (int dx = -1; dx <= 1; dx++) {
                        mbar_wait(&sm.x_full[x_slot], x_phase);
                        if (first && lane == 0) KZB_STAMP(item, 1);
                        const uint32_t x_lo = umma_desc_lo(smem_u32(sm.x + kRankBytes + size_t(x_slot) * kXStride));
                        for (int dy = -1; dy <= 1; dy++) {
                            mbar_wait(&sm.w_full[w_slot], w_phase);
                            tc_fence_after();
                            const uint32_t w_lo = umma_desc_lo(smem_u32(sm.w + size_t(w_slot) * kWBytes));
                            // N rows = ranks dy .. dy+7 of the slot (rank -1 / 8 = the zero atoms around it)
                            const uint32_t x_t = x_lo + uint32_t(dy * (kRankBytes >> 4));
                            if (lane == 0) {
#pragma unroll
                                for (int k = 0; k < 4; k++) {
                                    umma_bf16(tmem_d, desc_hi | uint64_t(w_lo + 2 * k), desc_hi | uint64_t(x_t + 2 * k), idesc,
                                              (!first || k != 0) ? 1u : 0u);
                                }
                                if (CL == 1) umma_commit(&sm.w_empty[w_slot]);
                                else umma_commit_multicast(&sm.w_empty[w_slot], kMask);
                                if (dy == 1) umma_commit(&sm.x_empty[x_slot]);
                            }
                            __syncwarp();
                            first = false;
                            if (++w_slot == p.b_slots) {
                                w_slot = 0;
                                w_phase ^= 1;
                            }
                        }
                        if (++x_slot == kXSlots) {
                            x_slot = 0;
                            x_phase ^= 1;
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
        // ------------------------------------------------------------------ epilogue (warps 2..9)
        // Eight warps, two per TMEM lane quarter (a warp may only read lanes 32*(warp%4)..+31): warp (q, half)
        // owns output channels 32q..32q+31 and ranks 4*half..4*half+3 of the unit.  Warps are independent of each
        // other: own staging buffers, own TMA stores (box = 32 channels x one rank of the 4 boards), own arrivals.
        // Two epilogue warps per scheduler hide each other's ALU / TMEM / shared-memory latencies.
        const int quarter = warp % 4;
        const int part = (warp - 2) / 4;     // which kChunks ranks of the unit
        const int c = quarter * 32 + lane;   // output channel = TMEM lane of this thread
        const bool warp_ok = quarter * 32 < p.n_store;  // narrow nets: upper warps have no channels
        const bool live = !(p.debug & 4) && warp_ok;
        uint8_t* const wstage = sm.stage + (warp - 2) * (kStageBufs * 32 * 64);
        int item = 0;
        for (int L = 0; L < p.num_layers; L++) {
            const TowerLayerDev ld = p.layers[L];
            const float bias = warp_ok ? ld.bias[c] : 0.0f;
            const bool relu = c < ld.relu_n;
            const bool to_x = ld.out_buf == 1;
            const bool has_res = ld.has_res != 0;
            const CUtensorMap* omap = &maps.out[to_x ? 0 : 1];
            int ul = 0;
            for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, item++, ul++) {
                const int buf = item & 1;
                // channel-major residual copy, laid out so that one warp-wide 32-byte access is 1 KiB contiguous:
                // XT[unit][rank][16-position half][channel][16 positions]
                __nv_bfloat16* xt = p.xt + size_t(unit) * (128 * 256) + size_t(c) * 16;
                uint32_t res[16];
                if (has_res && live) {
                    ldg256(xt + ((kChunks * part) * 2 + 0) * 2048, res);
                    ldg256(xt + ((kChunks * part) * 2 + 1) * 2048, res + 8);
                }
                mbar_wait(&sm.tmem_full[buf], (item >> 1) & 1);
                tc_fence_after();
                if (warp == 2 && lane == 0) KZB_STAMP(item, 3);
                const uint32_t taddr = tmem_base + uint32_t(buf) * 256u + (uint32_t(quarter * 32) << 16);

#pragma unroll 1
                for (int yy = 0; yy < kChunks; yy++) {
                    const int y = kChunks * part + yy;
                    uint32_t r[32];
                    tmem_ld32(taddr + y * 32, r);
                    tmem_ld_wait();
                    uint32_t packed[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        float f0 = __uint_as_float(r[2 * j]) + bias;
                        float f1 = __uint_as_float(r[2 * j + 1]) + bias;
                        if (relu) {
                            f0 = f0 < 0.0f ? 0.0f : f0;  // NaN stays NaN, like torch/ONNX Relu
                            f1 = f1 < 0.0f ? 0.0f : f1;
                        }
                        if (has_res) {
                            f0 += bf16_lo(res[j]);
                            f1 += bf16_hi(res[j]);
                        }
                        packed[j] = pack_bf16(f0, f1);
                    }
                    if (has_res && live && yy < kChunks - 1) {  // next rank's residual
                        ldg256(xt + ((y + 1) * 2 + 0) * 2048, res);
                        ldg256(xt + ((y + 1) * 2 + 1) * 2048, res + 8);
                    }
                    // staging buffer (yy & 1) was last read by the store of chunk yy-2: allow only the store of
                    // chunk yy-1 to be still reading
                    if (lane == 0) {
                        if (kStageBufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    __syncwarp();
                    uint8_t* stage = wstage + (yy & (kStageBufs - 1)) * (32 * 64);
                    if (live) {
                        // transposed tile [position j][32 channels]: the 32 lanes write 64 contiguous bytes per j
                        uint8_t* sp = stage + lane * 2;
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            *reinterpret_cast<uint16_t*>(sp + (2 * j) * 64) = uint16_t(packed[j] & 0xffffu);
                            *reinterpret_cast<uint16_t*>(sp + (2 * j + 1) * 64) = uint16_t(packed[j] >> 16);
                        }
                        if (to_x) {  // channel-major copy of the residual stream
                            stg256(xt + (y * 2 + 0) * 2048, packed);
                            stg256(xt + (y * 2 + 1) * 2048, packed + 8);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0 && live) {
                        tma_store_4d(omap, stage, quarter * 32, 0, unit * kBoards, y);
                        tma_store_commit();
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&sm.tmem_empty[buf]);
                    tma_store_wait_all();  // this warp's rows of the unit are in global memory (async proxy, like the loads)
                    mbar_arrive(&sm.ready[ul]);
                    if (warp == 2) KZB_STAMP(item, 4);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it or signal its barriers
    if (tl && threadIdx.x == 0) tl[2] = clock64();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace

size_t tower8_smem_bytes(int w_slots) {
    return 1024 + size_t(kXRegion) + size_t(w_slots) * kWBytes + kStageBytes + (2 * kXSlots + 2 * w_slots + 4 + kMaxLocalUnits) * 8 + 16;
}

int tower8_pick_b_slots(int /*n*/) {
    int slots = 6;
    while (slots > 3 && tower8_smem_bytes(slots) > 227 * 1024) slots--;
    return slots;
}

int tower8_max_local_units() { return kMaxLocalUnits; }

void tower8_prepare() {
    cudaFuncSetAttribute(tower8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(tower8_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

// cluster == 2 requires an even p.num_units (the executor pads the batch with one all-zero unit): then the two CTAs of a
// pair (blockIdx 2c, 2c+1; units blockIdx + k * gridDim) always own the same number of units and stay in lockstep on the
// shared weight ring.
void launch_tower8(const Tower8Maps& maps, const Tower8Params& p, int grid, cudaStream_t s) {
    if (p.num_units <= 0 || p.num_layers <= 0) return;
    if (p.cluster == 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(unsigned(std::min(grid & ~1, p.num_units)));
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = tower8_smem_bytes(p.b_slots);
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, tower8_kernel<2>, maps, p);
    } else {
        tower8_kernel<1><<<std::min(grid, p.num_units), kThreads, tower8_smem_bytes(p.b_slots), s>>>(maps, p);
    }
}

}  // namespace kzb
