// K1p: conv_tch.cu on the CTA-pair MMA (tcgen05.mma.cta_group::2), the 3x3 layers of boards larger than 8x8 (go).
//
// Two CTAs of a cluster own 256 consecutive padded rows (128 each).  Each stages its own activation tile (with halo, as in
// conv_tch.cu) and only HALF of every weight tile (n/2 output channels); the leader issues tcgen05.mma.cta_group::2 with
// M = 256 and both CTAs' TMEM receive their own 128 rows x n columns.  Per SM that is n/2 * 128 bytes of weights written
// into shared memory per (tap, k-block) instead of n * 128, and 8 KB instead of 12 KB read per MMA at n = 256.  The
// primitives were verified in isolation by scripts/micro/mma2_bench.cu (profiles/r01d_mma2_bench.txt): operand /
// accumulator placement of the pair MMA, the unswizzled halo-tile A operand in pair mode, full rate at N = 256 and N = 128.
//
// Synchronisation (L = leader = cluster rank 0, P = peer).  Only L's MMA thread issues, so its per-stage work must stay well
// below the 512 cycles four MMAs take: it waits on ONE barrier per stage.  (The first version relayed P's "stage landed"
// through a second barrier with an acquire.cluster wait per stage; it ran at 0.65x of conv_tch -- profiles/r02_go_pair.md.)
//   full[s] / a_full[a] (L's copy)  armed by L's producer with the bytes of BOTH CTAs; P's TMA loads complete their bytes on L's
//                                   barrier (cp.async.bulk.tensor .cta_group::2 with the peer bit of the barrier address cleared)
//   empty[s] / a_empty[a]           L's commit, multicast to both CTAs: the MMAs that read the slot are done
//   tmem_full[b]                    L's commit, multicast: accumulator b is complete in both CTAs' TMEM
//   tmem_empty[b] (on L)            8 arrivals: the 4 epilogue warps of L (local) and of P (remote) have drained accumulator b
#include "conv_epilogue.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace kzb {
namespace {

constexpr int kTileM = 128;
constexpr int kBlockK = 64;
constexpr int kThreads = 192;
constexpr int kASlots = 2;
constexpr uint16_t kPairMask = 3;

using namespace tc;

struct SmemLayout {
    uint8_t* b_base;  // ring of `stages` half weight tiles, n/2 * 128 bytes each (1024-byte aligned)
    uint8_t* a_base;  // activation slots
    uint64_t *full, *empty, *a_full, *a_empty, *tmem_full, *tmem_empty;
    uint32_t* tmem_ptr;
    float* bias;
};

__host__ __device__ inline size_t a_slot_bytes(int a_rows) { return (size_t(a_rows) * 128 + 1023) & ~size_t(1023); }
__host__ __device__ inline size_t b_stage_bytes(int n) { return (size_t(n / 2) * 128 + 1023) & ~size_t(1023); }

__device__ __forceinline__ SmemLayout carve(uint8_t* base, int n, int stages, int a_rows) {
    SmemLayout s;
    s.b_base = base;
    s.a_base = base + size_t(stages) * b_stage_bytes(n);
    uint8_t* p = s.a_base + kASlots * a_slot_bytes(a_rows);
    s.full = reinterpret_cast<uint64_t*>(p);
    s.empty = s.full + stages;
    s.a_full = s.empty + stages;
    s.a_empty = s.a_full + kASlots;
    s.tmem_full = s.a_empty + kASlots;
    s.tmem_empty = s.tmem_full + 2;
    s.tmem_ptr = reinterpret_cast<uint32_t*>(s.tmem_empty + 2);
    s.bias = reinterpret_cast<float*>(s.tmem_ptr + 4);
    return s;
}

__device__ __forceinline__ uint64_t umma_desc_nosw_hi(uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t(lbo_bytes >> 4) << 16) | (uint64_t(sbo_bytes >> 4) << 32) | (uint64_t(1) << 46);
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// TMA load issued by either CTA of the pair; the bytes are completed on the LEADER's barrier (same offset, peer bit cleared)
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {  // arrives on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(kPairMask)
                 : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
    conv_tchp_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_bh, const ConvTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const SmemLayout sm = carve(smem, p.n, p.stages, p.a_rows);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t b_stage = uint32_t(b_stage_bytes(p.n));
    const uint32_t b_bytes = uint32_t(p.n / 2) * 128u;
    const uint32_t a_bytes = uint32_t(a_slot_bytes(p.a_rows));
    const uint32_t chunk_bytes = uint32_t(p.a_rows) * 16u;
    const int num_pairs = (p.num_tiles + 1) / 2;          // pair tiles of 256 rows
    const int cluster_id = int(blockIdx.x) / 2, num_clusters = int(gridDim.x) / 2;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_bh)) : "memory");
        for (int i = 0; i < p.stages; i++) {
            mbar_init(&sm.full[i], 1);
            mbar_init(&sm.empty[i], 1);
        }
        for (int i = 0; i < kASlots; i++) {
            mbar_init(&sm.a_full[i], 1);
            mbar_init(&sm.a_empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&sm.tmem_full[i], 1);
            mbar_init(&sm.tmem_empty[i], 8);  // 4 epilogue warps of each CTA (only the leader's copy is used)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.tmem_ptr)),
                     "r"(uint32_t(p.tmem_cols))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.n; i += kThreads) sm.bias[i] = p.bias[i];
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers exist before anything is committed to / arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_ptr;
    const uint32_t acc_stride = uint32_t(p.tmem_cols / 2);
    if (p.pdl) grid_dep_launch_dependents();  // the next layer may set itself up while this one computes

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs, own rows / own weight half)
        if (lane == 0) {
            int stage = 0, a_slot = 0;
            uint32_t phase = 0, a_phase = 0;
            int pre = 0;  // weight tiles issued before the previous layer had finished (they do not depend on it)
            if (p.pdl) {
                if (cluster_id < num_pairs) {
                    for (; pre < p.stages && pre < 9 * p.kblocks; pre++) {
                        if (leader) mbar_expect_tx(&sm.full[pre], 2 * b_bytes);
                        tma2_load_2d(&tmap_bh, &sm.full[pre], sm.b_base + size_t(pre) * b_stage, (pre % 9) * p.cin_pad + (pre / 9) * kBlockK,
                                     int(rank) * (p.n / 2));
                    }
                }
                grid_dep_wait();
            }
            for (int pt = cluster_id; pt < num_pairs; pt += num_clusters) {
                const int tile = 2 * pt + int(rank);  // may be one past the last tile: its rows are zero-filled and never stored
                for (int kb = 0; kb < p.kblocks; kb++) {
                    mbar_wait(&sm.a_empty[a_slot], a_phase ^ 1);
                    if (leader) mbar_expect_tx(&sm.a_full[a_slot], 2u * 8u * chunk_bytes);
                    uint8_t* a_dst = sm.a_base + size_t(a_slot) * a_bytes;
#pragma unroll
                    for (int kc = 0; kc < 8; kc++)
                        tma2_load_2d(&tmap_a, &sm.a_full[a_slot], a_dst + size_t(kc) * chunk_bytes, kb * kBlockK + kc * 8,
                                     tile * kTileM - p.halo);
                    if (++a_slot == kASlots) {
                        a_slot = 0;
                        a_phase ^= 1;
                    }
                    for (int tap = 0; tap < 9; tap++) {
                        if (pre > 0) {
                            pre--;  // already in flight
                        } else {
                            mbar_wait(&sm.empty[stage], phase ^ 1);
                            if (leader) mbar_expect_tx(&sm.full[stage], 2 * b_bytes);
                            tma2_load_2d(&tmap_bh, &sm.full[stage], sm.b_base + size_t(stage) * b_stage, tap * p.cin_pad + kb * kBlockK,
                                         int(rank) * (p.n / 2));
                        }
                        if (++stage == p.stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA warp: only the leader's issues
        if (leader) {
            const uint32_t idesc = umma_idesc_bf16(2 * kTileM, p.n);
            const uint64_t b_hi = umma_desc_sw128_hi();
            const uint64_t a_hi = umma_desc_nosw_hi(chunk_bytes, 128);
            int stage = 0, a_slot = 0;
            uint32_t phase = 0, a_phase = 0;
            int local = 0;
            for (int pt = cluster_id; pt < num_pairs; pt += num_clusters, local++) {
                const int buf = local & 1;
                const uint32_t buf_phase = (local >> 1) & 1;
                mbar_wait_cluster(&sm.tmem_empty[buf], buf_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * acc_stride;
                bool first = true;
                for (int kb = 0; kb < p.kblocks; kb++) {
                    mbar_wait(&sm.a_full[a_slot], a_phase);  // both CTAs' activation tiles
                    const uint32_t a_lo = umma_desc_lo(smem_u32(sm.a_base + size_t(a_slot) * a_bytes));
                    for (int tap = 0; tap < 9; tap++) {
                        mbar_wait(&sm.full[stage], phase);  // both halves of the weight tile
                        tc_fence_after();
                        const uint32_t b_lo = umma_desc_lo(smem_u32(sm.b_base + size_t(stage) * b_stage));
                        const uint32_t a_t = a_lo + uint32_t(p.halo + (tap / 3 - 1) * p.lay.rank_pitch + (tap % 3 - 1));
                        if (lane == 0) {
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; k++)
                                umma2_bf16(tmem_d, a_hi | uint64_t(a_t + uint32_t(k) * (2u * chunk_bytes >> 4)), b_hi | uint64_t(b_lo + 2 * k), idesc,
                                           (!first || k != 0) ? 1u : 0u);
                            umma2_commit(&sm.empty[stage]);
                            if (tap == 8) umma2_commit(&sm.a_empty[a_slot]);
                        }
                        __syncwarp();
                        first = false;
                        if (++stage == p.stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    if (++a_slot == kASlots) {
                        a_slot = 0;
                        a_phase ^= 1;
                    }
                }
                if (lane == 0) umma2_commit(&sm.tmem_full[buf]);  // accumulator complete in both CTAs -> both epilogues
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5), each CTA drains its own 128 rows
        const int quarter = warp % 4;
        int local = 0;
        if (p.pdl) grid_dep_wait();  // the residual rows are the previous layers' output
        for (int pt = cluster_id; pt < num_pairs; pt += num_clusters, local++) {
            const int buf = local & 1;
            const uint32_t buf_phase = (local >> 1) & 1;
            const int tile = 2 * pt + int(rank);
            conv_epilogue_tile(p, sm.bias, tile, quarter, lane, tmem_base + buf * acc_stride, &sm.tmem_full[buf], buf_phase, nullptr, 0, p.n_store);
            if (lane == 0) mbar_arrive_remote(&sm.tmem_empty[buf], 0);  // the leader's barrier, from either CTA
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // no CTA leaves while the pair may still read its shared memory or signal its barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(p.tmem_cols))
                     : "memory");
    }
}

}  // namespace

size_t conv_tchp_smem_bytes(int n, int stages, int a_rows) {
    return 1024 /*alignment slack*/ + size_t(stages) * b_stage_bytes(n) + kASlots * a_slot_bytes(a_rows) + (2 * stages + 2 * kASlots + 4) * 8 + 16 +
           size_t(n) * 4;
}

int conv_tchp_pick_stages(int n, int a_rows) {
    const size_t budget = 227 * 1024;
    int stages = 12;
    while (stages > 2 && conv_tchp_smem_bytes(n, stages, a_rows) > budget) stages--;
    return stages;
}

void conv_tchp_prepare() { cudaFuncSetAttribute(conv_tchp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }

// tmap_a: unswizzled (8 channels, a_rows rows) boxes over the channels-last rows; tmap_bh: weight map with a box of n/2 rows
void launch_conv_tchp(const CUtensorMap& tmap_a, const CUtensorMap& tmap_bh, const ConvTcParams& p, int grid, cudaStream_t s) {
    if (p.num_tiles <= 0) return;
    const int pairs = (p.num_tiles + 1) / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(2 * std::min(grid / 2, pairs)));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = conv_tchp_smem_bytes(p.n, p.stages, p.a_rows);
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
    cudaLaunchKernelEx(&cfg, conv_tchp_kernel, tmap_a, tmap_bh, p);
}

}  // namespace kzb
