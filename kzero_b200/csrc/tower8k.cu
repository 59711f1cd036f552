// K1, whole-tower persistent kernel for 8x8 boards, second generation: activations are staged ONCE per 64-channel
// k-block and every one of the nine 3x3 taps is only a different UMMA descriptor start address.
//
// Why: the first generation (tower8.cu) stages three x-shifted copies of every activation tile because a one-row
// shift is not expressible under the 128-byte swizzle.  Ablations on B200 show that kernel pinned
// by the bytes each SM ingests from L2 (480 KB per 4-board unit and layer, ~46 B/cycle/SM), not by the tensor pipe.
// Here the activation operand uses the SWIZZLE_NONE K-major canonical layout instead:
//
//   * "core matrix" = 8 positions x 8 channels (16 bytes per position, 128 contiguous bytes); a k-chunk (8 channels)
//     of a 4-board unit is 32 groups of 8 positions = the 8 files of one (rank, board), ordered [rank][board].
//   * group pitch (descriptor SBO) = 144 bytes: 8 positions + ONE ZERO PAD ROW.  A horizontal tap dx is start address
//     +-16 bytes: file -1 reads the previous group's pad row, file 8 reads the own pad row.  A vertical tap dy is
//     +-4 groups (576 bytes): rank -1 / 8 fall into the all-zero gap between consecutive k-chunks.
//   * k-chunk pitch (descriptor LBO) = 5248 bytes = 4608 data + 640 zero gap.  One K=16 MMA reads chunks 2j, 2j+1.
//   Measured (scripts/micro/nosw_bench.cu): all 9 taps exact, 128.0 cycles per M128 N256 K16 MMA = full rate, also
//   for the 16-byte-misaligned dx = +-1 start addresses.
//   * the TMA engine produces this layout directly: activations live in global memory k-chunk-major,
//     A[kc][board][y][x][8 channels] (128 contiguous bytes per (kc, board, rank)); the box is (72 elements, 4 boards,
//     8 ranks) of one k-chunk -- 72 > 64 = the tensor's inner extent, so the engine zero-fills the pad row itself.
//   SM ingress per unit and layer: 64 KB activations + 288 KB weights = 352 KB (was 480 KB).
//
// Everything else follows tower8.cu: GEMM orientation D^T[128 out-channels, 256 positions] (weights = UMMA A operand,
// SWIZZLE_128B tiles streamed through a ring, multicast between the two CTAs of a cluster), two TMEM accumulators,
// warp 0 TMA producer / warp 1 MMA issuer / warps 2..9 epilogue (thread = output channel), channel-major residual copy
// XT, ready[unit] barrier between a unit's layers.  The epilogue stores k-chunk-major through a bank-conflict-free
// staging tile [board][kc][80 elements] and one TMA store per rank; the LAST layer stores row-major [position][C]
// (what the head kernels read).  Reference semantics: block = x + relu(bn(conv(relu(bn(conv(x)))))), relu BEFORE the
// add (python/lib/model/post_act.py:218-228).
#include "kernels.cuh"
#include "tc_common.cuh"

namespace kzb {
namespace {

using namespace tc;

constexpr int kEpiWarps = 8;
constexpr int kChunks = 32 / kEpiWarps;        // ranks per epilogue warp and unit
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kGroup = 144;                    // 8 positions x 16 B + one zero pad row
constexpr int kChunkData = 32 * kGroup;        // 4608 B: the 8 ranks of one k-chunk of a 4-board unit
constexpr int kLbo = 5248;                     // k-chunk pitch (41 x 128: keeps every TMA destination 128-byte aligned)
constexpr int kLead = kLbo - kChunkData;       // 640 B zero gap in front of every k-chunk
constexpr int kHalfBytes = kLead + 8 * kLbo;   // one 64-channel k-block of a unit incl. its gaps: 42,624 B
constexpr int kXSlots = 2;
constexpr int kWBytes = 128 * 128;             // weight tile: 128 out-channels x 64 k
constexpr int kStageWarp = 4 * 4 * 160;        // per-warp output staging: [4 boards][4 kc][80 elements]
constexpr int kStageBytes = kEpiWarps * kStageWarp;
constexpr int kMaxLocalUnits = 16;

struct SmemK {
    uint8_t* x;      // kXSlots x kHalfBytes
    uint8_t* w;      // w_slots x 16 KiB
    uint8_t* stage;  // kEpiWarps x kStageWarp
    uint64_t *x_full, *x_empty, *w_full, *w_empty, *tmem_full, *tmem_empty, *ready;
    uint32_t* tmem_ptr;
};

__device__ __forceinline__ SmemK carve_k(uint8_t* base, int w_slots) {
    SmemK s;
    s.w = base;  // 1 KiB aligned (SWIZZLE_128B atoms)
    s.x = s.w + size_t(w_slots) * kWBytes;
    s.stage = s.x + size_t(kXSlots) * kHalfBytes;
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

// Work units.  Legacy scheme: 4-board units, unit u = blockIdx + ul * gridDim.  Balanced scheme (p.balanced): CTA c owns
// the contiguous boards [c*base + min(c, rem), ...) -- base or base+1 of them, 6..8 -- as two units of 4 or 3 boards, so
// that every SM carries the same work within one board (1024 boards on 148 SMs: 136 x 7 + 12 x 6 instead of 108 x 8 + 40 x 4).
// A 3-board unit is an N = 192 MMA (96 cycles, still above the ~94-cycle floor of an M128 MMA).
struct Unit {
    int board0, nb;
};
__device__ __forceinline__ int local_units(const Tower8Params& p) {
    if (p.balanced) return 2;
    return (p.num_units - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
}
__device__ __forceinline__ Unit unit_of(const Tower8Params& p, int ul) {
    if (!p.balanced) return {(int(blockIdx.x) + ul * int(gridDim.x)) * 4, 4};
    const int c = int(blockIdx.x);
    const int n = p.bal_base + (c < p.bal_rem ? 1 : 0);
    const int start = c * p.bal_base + min(c, p.bal_rem);
    const int u0 = (n + 1) / 2;
    return ul == 0 ? Unit{start, u0} : Unit{start + u0, n - u0};
}

// SWIZZLE_NONE K-major descriptor, address-independent part: LBO = k-chunk pitch, SBO = group pitch, version 1
__device__ __forceinline__ uint64_t umma_desc_nosw_hi() {
    return (uint64_t(kLbo >> 4) << 16) | (uint64_t(kGroup >> 4) << 32) | (uint64_t(1) << 46);
}

template <int CL>
__global__ void __launch_bounds__(kThreads, 1)
    tower8k_kernel(const __grid_constant__ Tower8kMaps maps, const Tower8Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const SmemK sm = carve_k(smem, p.b_slots);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    // development aid (KZB_TIMELINE=tower8): stamps as in tower8.cu
    unsigned long long* tl = p.timeline ? p.timeline + size_t(blockIdx.x) * 1024 : nullptr;
#define KZB_STAMP(item, k) do { if (tl && (item) < 127) tl[8 + (item) * 8 + (k)] = clock64(); } while (0)
    if (tl && threadIdx.x == 0) tl[0] = clock64();

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 6; i++) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.a[i / 2][i % 2])) : "memory");
        for (int i = 0; i < 2; i++) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w[i])) : "memory");
        for (int i = 0; i < 6; i++) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.out[i / 2][i % 2])) : "memory");
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
    // the activation slots start all-zero: the gaps between k-chunks (the conv's vertical padding) are never written
    // again; the weight ring is cleared so that rows >= n of a narrow net never feed NaN bit patterns to the tensor core
    for (int i = threadIdx.x; i < (p.b_slots * kWBytes + kXSlots * kHalfBytes) / 16; i += kThreads)
        reinterpret_cast<uint4*>(sm.w)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zeros -> visible to UMMA reads
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast / committed to them
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_ptr;
    const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0;
    constexpr uint16_t kMask = uint16_t((1u << CL) - 1);
    if (tl && threadIdx.x == 0) tl[1] = clock64();
    if (p.pdl) grid_dep_launch_dependents();  // the head kernel may set itself up on every SM this kernel leaves

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            if (p.pdl) grid_dep_wait();  // the planes of layer 0 are the encode kernel's output (everything above overlapped it)
            const uint32_t w_bytes = uint32_t(p.n) * 128u;
            const int n_local = local_units(p);
            int x_slot = 0, w_slot = 0;
            uint32_t x_phase = 0, w_phase = 0;
            int pitem = 0;
            for (int L = 0; L < p.num_layers; L++) {
                const TowerLayerDev ld = p.layers[L];
                const CUtensorMap* wmap = &maps.w[ld.w_map];
                for (int ul = 0; ul < n_local; ul++, pitem++) {
                    const Unit un = unit_of(p, ul);
                    const CUtensorMap* amap = &maps.a[ld.a_map][un.nb - 3];
                    // rows of this unit written by layer L-1's epilogue must be complete and visible
                    if (L > 0) mbar_wait(&sm.ready[ul], uint32_t(L - 1) & 1);
                    KZB_STAMP(pitem, 5);
                    for (int kb = 0; kb < ld.kblocks; kb++) {
                        mbar_wait(&sm.x_empty[x_slot], x_phase ^ 1);
                        if (p.debug & 1) {
                            mbar_arrive(&sm.x_full[x_slot]);
                        } else {
                            // box = (72 elements, nb boards, 9 ranks): rank 8 does not exist, so the engine also writes the zero
                            // rank that the dy = +1 taps of this chunk and the dy = -1 taps of the next one read
                            mbar_expect_tx(&sm.x_full[x_slot], uint32_t(ld.kchunks) * uint32_t(9 * un.nb * kGroup));
                            uint8_t* half = sm.x + size_t(x_slot) * kHalfBytes + kLead;
                            for (int c = 0; c < ld.kchunks; c++)
                                tma_load_4d(amap, &sm.x_full[x_slot], half + size_t(c) * kLbo, 0, un.board0, 0, kb * 8 + c);
                        }
                        if (++x_slot == kXSlots) {
                            x_slot = 0;
                            x_phase ^= 1;
                        }
                        for (int tap = 0; tap < 9; tap++) {
                            mbar_wait(&sm.w_empty[w_slot], w_phase ^ 1);
                            if (p.debug & 2) {
                                mbar_arrive(&sm.w_full[w_slot]);
                            } else if (CL == 1) {
                                mbar_expect_tx(&sm.w_full[w_slot], w_bytes);
                                tma_load_2d(wmap, &sm.w_full[w_slot], sm.w + size_t(w_slot) * kWBytes, tap * ld.cin_pad + kb * 64,
                                            ld.w_row0);
                            } else {  // my n/CL rows of the tile, into every CTA of the cluster
                                mbar_expect_tx(&sm.w_full[w_slot], w_bytes);
                                const int rows = p.n / CL;
                                tma_load_2d_multicast(wmap, &sm.w_full[w_slot],
                                                      sm.w + size_t(w_slot) * kWBytes + size_t(cta_rank) * rows * 128,
                                                      tap * ld.cin_pad + kb * 64, ld.w_row0 + int(cta_rank) * rows, kMask);
                            }
                            if (++w_slot == p.b_slots) {
                                w_slot = 0;
                                w_phase ^= 1;
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
        const int n_local = local_units(p);
        const uint64_t w_hi = umma_desc_sw128_hi();
        const uint64_t x_hi = umma_desc_nosw_hi();
        int x_slot = 0, w_slot = 0;
        uint32_t x_phase = 0, w_phase = 0;
        int item = 0;
        for (int L = 0; L < p.num_layers; L++) {
            const int kblocks = p.layers[L].kblocks, ksteps = p.layers[L].ksteps;
            for (int ul = 0; ul < n_local; ul++, item++) {
                const int nb = unit_of(p, ul).nb;
                const uint32_t idesc = umma_idesc_bf16(128, nb * 64);
                const int buf = item & 1;
                mbar_wait(&sm.tmem_empty[buf], ((item >> 1) & 1) ^ 1);
                tc_fence_after();
                if (lane == 0) KZB_STAMP(item, 0);
                const uint32_t tmem_d = tmem_base + uint32_t(buf) * 256u;
                bool first = true;
                for (int kb = 0; kb < kblocks; kb++) {
                    mbar_wait(&sm.x_full[x_slot], x_phase);
                    if (first && lane == 0) KZB_STAMP(item, 1);
                    const uint32_t x_lo = umma_desc_lo(smem_u32(sm.x + size_t(x_slot) * kHalfBytes + kLead));
                    for (int dy = -1; dy <= 1; dy++) {
                        for (int dx = -1; dx <= 1; dx++) {
                            mbar_wait(&sm.w_full[w_slot], w_phase);
                            tc_fence_after();
                            const uint32_t w_lo = umma_desc_lo(smem_u32(sm.w + size_t(w_slot) * kWBytes));
                            // tap (dy, dx): nb groups per rank, 16 bytes per file -- in 16-byte units
                            const uint32_t x_t = x_lo + uint32_t(dy * (nb * kGroup / 16) + dx);
                            if (lane == 0) {
#pragma unroll
                                for (int k = 0; k < 4; k++) {
                                    if (k < ksteps)
                                        umma_bf16(tmem_d, w_hi | uint64_t(w_lo + 2 * k), x_hi | uint64_t(x_t + k * (2 * kLbo / 16)), idesc,
                                                  (!first || k != 0) ? 1u : 0u);
                                }
                                if (CL == 1) umma_commit(&sm.w_empty[w_slot]);
                                else umma_commit_multicast(&sm.w_empty[w_slot], kMask);
                                if (dy == 1 && dx == 1) umma_commit(&sm.x_empty[x_slot]);
                            }
                            __syncwarp();
                            first = false;
                            if (++w_slot == p.b_slots) {
                                w_slot = 0;
                                w_phase ^= 1;
                            }
                        }
                    }
                    if (++x_slot == kXSlots) {
                        x_slot = 0;
                        x_phase ^= 1;
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
        // Two warps per TMEM lane quarter (a warp may only read lanes 32*(warp%4)..+31): warp (q, part) owns output
        // channels 32q..32q+31 and ranks kChunks*part.. of the unit; own staging tile, own TMA stores, own arrivals.
        const int quarter = warp % 4;
        const int part = (warp - 2) / 4;
        const int c = quarter * 32 + lane;   // output channel = TMEM lane of this thread
        const bool warp_ok = quarter * 32 < p.n_store;  // narrow nets: upper warps have no channels
        const bool live = !(p.debug & 4) && warp_ok;
        uint8_t* const stage = sm.stage + (warp - 2) * kStageWarp;
        const int n_local = local_units(p);
        int item = 0;
        for (int L = 0; L < p.num_layers; L++) {
            const TowerLayerDev ld = p.layers[L];
            const float bias = warp_ok ? ld.bias[c] : 0.0f;
            const bool relu = c < ld.relu_n;
            const bool to_x = ld.out_buf == 1;
            const bool has_res = ld.has_res != 0;
            const bool rowmajor = ld.out_rowmajor != 0;
            for (int ul = 0; ul < n_local; ul++, item++) {
                const Unit un = unit_of(p, ul);
                const int nb = un.nb;
                const CUtensorMap* omap = &maps.out[rowmajor ? 2 : (to_x ? 0 : 1)][nb - 3];
                const int buf = item & 1;
                // channel-major residual copy, laid out so that one warp-wide 32-byte access is 1 KiB contiguous:
                // XT[slot][rank][16-position half][channel][16 positions], one slot per (CTA, local unit)
                __nv_bfloat16* xt = p.xt + size_t(ul * int(gridDim.x) + int(blockIdx.x)) * (128 * 256) + size_t(c) * 16;
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
                    tmem_ld32(taddr + y * (nb * 8), r);  // columns of rank y: nb boards x 8 files (a 3-board unit ignores the last 8)
                    tmem_ld_wait();
                    uint32_t packed[16];
                    const bool rank_ok = y < p.board_h;  // boards smaller than 8x8: squares outside stay zero (padding)
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
                        if (!rank_ok || ((2 * j) & 7) >= p.board_w) f0 = 0.0f;
                        if (!rank_ok || ((2 * j + 1) & 7) >= p.board_w) f1 = 0.0f;
                        packed[j] = pack_bf16(f0, f1);
                    }
                    if (has_res && live && yy < kChunks - 1) {  // next rank's residual
                        ldg256(xt + ((y + 1) * 2 + 0) * 2048, res);
                        ldg256(xt + ((y + 1) * 2 + 1) * 2048, res + 8);
                    }
                    // the staging tile was last read by the TMA store of the previous chunk
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                    if (live) {
                        if (rowmajor) {
                            // [position j][32 channels]: the 32 lanes write 64 contiguous bytes per j
                            uint8_t* sp = stage + lane * 2;
#pragma unroll
                            for (int j = 0; j < 16; j++) {
                                if ((j >> 2) < nb) {
                                    *reinterpret_cast<uint16_t*>(sp + (2 * j) * 64) = uint16_t(packed[j] & 0xffffu);
                                    *reinterpret_cast<uint16_t*>(sp + (2 * j + 1) * 64) = uint16_t(packed[j] >> 16);
                                }
                            }
                        } else {
                            // [board][kc][80 elements]: position j = board*8 + file -> board*640 + kc*160 + file*16; the 160-byte
                            // kc pitch spreads the four 8-channel groups of the warp over all 32 banks
                            // lane pairs (channels c, c+1) swap halves so that every lane stores whole 32-bit words: the even lane
                            // writes (c, c+1) of the even position 2j, the odd lane of position 2j+1 -- one conflict-free
                            // wavefront per instruction (shared-memory bandwidth is what the MMA operand fetch competes for)
                            const bool odd = lane & 1;
                            uint8_t* sp = stage + (lane >> 3) * 160 + (lane & 6) * 2 + (odd ? 16 : 0);
#pragma unroll
                            for (int j = 0; j < 16; j++) {
                                const uint32_t other = __shfl_xor_sync(0xffffffffu, packed[j], 1);
                                const uint32_t word = odd ? __byte_perm(other, packed[j], 0x7632) : __byte_perm(packed[j], other, 0x5410);
                                if ((j >> 2) < nb) *reinterpret_cast<uint32_t*>(sp + ((2 * j) >> 3) * 640 + ((2 * j) & 7) * 16) = word;
                            }
                        }
                        if (to_x) {  // channel-major copy of the residual stream
                            stg256(xt + (y * 2 + 0) * 2048, packed);
                            stg256(xt + (y * 2 + 1) * 2048, packed + 8);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0 && live && !(p.debug & 8)) {
                        if (rowmajor) tma_store_4d(omap, stage, quarter * 32, 0, un.board0, y);   // (c, x, board, y)
                        else tma_store_4d(omap, stage, 0, quarter * 4, un.board0, y);             // (x*8+c8, kc, board, y)
                        tma_store_commit();
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&sm.tmem_empty[buf]);
                    tma_store_wait_all();  // this warp's part of the unit is in global memory (async proxy, like the loads)
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

size_t tower8k_smem_bytes(int w_slots) {
    return 1024 + size_t(w_slots) * kWBytes + size_t(kXSlots) * kHalfBytes + kStageBytes + (2 * kXSlots + 2 * w_slots + 4 + kMaxLocalUnits) * 8 + 16;
}

int tower8k_pick_b_slots() {
    int slots = 8;
    while (slots > 3 && tower8k_smem_bytes(slots) > 227 * 1024) slots--;
    return slots;
}

int tower8k_max_local_units() { return kMaxLocalUnits; }

void tower8k_prepare() {
    cudaFuncSetAttribute(tower8k_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(tower8k_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

// cluster == 2 requires an even p.num_units (see launch_tower8)
void launch_tower8k(const Tower8kMaps& maps, const Tower8Params& p, int grid, cudaStream_t s) {
    if (p.num_units <= 0 || p.num_layers <= 0) return;
    if (p.cluster == 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(unsigned(p.balanced ? p.bal_grid : std::min(grid & ~1, p.num_units)));
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = tower8k_smem_bytes(p.b_slots);
        cfg.stream = s;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = p.pdl ? 2 : 1;
        cudaLaunchKernelEx(&cfg, tower8k_kernel<2>, maps, p);
    } else {
        tower8k_kernel<1><<<p.balanced ? p.bal_grid : std::min(grid, p.num_units), kThreads, tower8k_smem_bytes(p.b_slots), s>>>(maps, p);
    }
}

}  // namespace kzb
