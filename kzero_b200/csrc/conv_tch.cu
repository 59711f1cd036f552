// K1h: conv3x3 on padded rows with the activation tile loaded ONCE per k-block (conv_tc.cu re-loads it for each tap).
//
// Same GEMM view, tile shape and epilogue as conv_tc.cu (D[128 rows, cout] += A_tap[128, 64] * W[cout, 64] per tap and
// k-block), but the A operand lives in shared memory in the SWIZZLE_NONE K-major layout: 8 boxes of (8 channels, R rows)
// land as [k-chunk][row][16 bytes], R = 128 + 2 * halo rows (rounded up to 8) with halo = rank_pitch + 1.  Rows are then 16 bytes apart
// (SBO = 128 bytes per 8-row group, LBO = R * 16 bytes per k-chunk), and a descriptor may start at any row: tap (dy, dx)
// is the same tile read from row halo + dy * rank_pitch + dx.  Per 128-row tile and k-block the SM receives R * 128 bytes
// of activations instead of 9 * 16 KB, the weight stream (cout * 128 bytes per tap and k-block) is unchanged: 27 % fewer
// bytes into shared memory per MMA at cout = 256.  profiles/r01d_go9_conv_tc_ncu.md has the measurement that motivates
// it.  Weights keep the SWIZZLE_128B layout and an own ring.  p.n_split = 2 (small batches: fewer tiles than half the SMs)
// makes a work item one tile x one HALF of the output channels, so that twice as many SMs share the layer.  Default for 3x3 layers in mode 0 (KZB_CONV_HALO=0: conv_tc.cu).
#include "conv_epilogue.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace kzb {
namespace {

constexpr int kTileM = 128;
constexpr int kBlockK = 64;
constexpr int kThreads = 192;
constexpr int kASlots = 2;

using namespace tc;

struct SmemLayout {
    uint8_t* b_base;  // weight ring, `stages` slots of n * 128 bytes (1024-byte aligned)
    uint8_t* a_base;  // activation slots, kASlots of a_bytes
    uint64_t* full;
    uint64_t* empty;
    uint64_t* a_full;
    uint64_t* a_empty;
    uint64_t* tmem_full;
    uint64_t* tmem_empty;
    uint32_t* tmem_ptr;
    float* bias;
};

__host__ __device__ inline size_t a_slot_bytes(int a_rows) { return (size_t(a_rows) * 128 + 1023) & ~size_t(1023); }

__device__ __forceinline__ SmemLayout carve(uint8_t* base, int n, int stages, int a_rows) {
    SmemLayout s;
    s.b_base = base;
    s.a_base = base + size_t(stages) * n * 128;
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

// SWIZZLE_NONE K-major descriptor, address-independent part: LBO = k-chunk pitch, SBO = 8-row group pitch, version 1
__device__ __forceinline__ uint64_t umma_desc_nosw_hi(uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t(lbo_bytes >> 4) << 16) | (uint64_t(sbo_bytes >> 4) << 32) | (uint64_t(1) << 46);
}

__global__ void __launch_bounds__(kThreads, 1)
    conv_tch_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const ConvTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const SmemLayout sm = carve(smem, p.n, p.stages, p.a_rows);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int n_eff = p.n / p.n_split;  // output channels per work item
    const int num_items = p.num_tiles * p.n_split;
    const uint32_t b_slot = uint32_t(p.n) * 128u;     // ring pitch (sized for the whole n)
    const uint32_t b_bytes = uint32_t(n_eff) * 128u;  // bytes one weight tile brings
    const uint32_t a_bytes = uint32_t(a_slot_bytes(p.a_rows));
    const uint32_t chunk_bytes = uint32_t(p.a_rows) * 16u;  // one k-chunk (8 channels) of the activation tile

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
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
            mbar_init(&sm.tmem_empty[i], 4);  // one arrive per epilogue warp
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
    const uint32_t acc_stride = uint32_t(p.tmem_cols / 2);
    if (p.pdl) grid_dep_launch_dependents();  // the next layer may set itself up while this one computes

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0, a_slot = 0;
            uint32_t phase = 0, a_phase = 0;
            int pre = 0;  // weight tiles issued before the previous layer had finished (they do not depend on it)
            if (p.pdl) {
                if (int(blockIdx.x) < num_items) {
                    const int n0 = (int(blockIdx.x) % p.n_split) * n_eff;
                    for (; pre < p.stages && pre < 9 * p.kblocks; pre++) {
                        mbar_expect_tx(&sm.full[pre], b_bytes);
                        tma_load_2d(&tmap_b, &sm.full[pre], sm.b_base + size_t(pre) * b_slot, (pre % 9) * p.cin_pad + (pre / 9) * kBlockK, n0);
                    }
                }
                grid_dep_wait();
            }
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int tile = item / p.n_split, n0 = (item % p.n_split) * n_eff;
                for (int kb = 0; kb < p.kblocks; kb++) {
                    // the activation tile of this k-block, halo rows included (rows off either end are zero-filled)
                    mbar_wait(&sm.a_empty[a_slot], a_phase ^ 1);
                    mbar_expect_tx(&sm.a_full[a_slot], 8u * chunk_bytes);
                    uint8_t* a_dst = sm.a_base + size_t(a_slot) * a_bytes;
#pragma unroll
                    for (int kc = 0; kc < 8; kc++)
                        tma_load_2d(&tmap_a, &sm.a_full[a_slot], a_dst + size_t(kc) * chunk_bytes, kb * kBlockK + kc * 8,
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
                            mbar_expect_tx(&sm.full[stage], b_bytes);
                            tma_load_2d(&tmap_b, &sm.full[stage], sm.b_base + size_t(stage) * b_slot, tap * p.cin_pad + kb * kBlockK, n0);
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
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = umma_idesc_bf16(kTileM, n_eff);
        const uint64_t b_hi = umma_desc_sw128_hi();
        const uint64_t a_hi = umma_desc_nosw_hi(chunk_bytes, 128);
        int stage = 0, a_slot = 0;
        uint32_t phase = 0, a_phase = 0;
        int local = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x, local++) {
            const int buf = local & 1;
            const uint32_t buf_phase = (local >> 1) & 1;
            mbar_wait(&sm.tmem_empty[buf], buf_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + buf * acc_stride;
            bool first = true;
            for (int kb = 0; kb < p.kblocks; kb++) {
                mbar_wait(&sm.a_full[a_slot], a_phase);
                const uint32_t a_lo = umma_desc_lo(smem_u32(sm.a_base + size_t(a_slot) * a_bytes));
                for (int tap = 0; tap < 9; tap++) {
                    mbar_wait(&sm.full[stage], phase);
                    tc_fence_after();
                    const uint32_t b_lo = umma_desc_lo(smem_u32(sm.b_base + size_t(stage) * b_slot));
                    // tap (dy, dx) = the same tile read from another row; one row is 16 bytes = one descriptor unit
                    const uint32_t a_t = a_lo + uint32_t(p.halo + (tap / 3 - 1) * p.lay.rank_pitch + (tap % 3 - 1));
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; k++)  // 16 channels = two k-chunks of the activation tile
                            umma_bf16(tmem_d, a_hi | uint64_t(a_t + uint32_t(k) * (2u * chunk_bytes >> 4)), b_hi | uint64_t(b_lo + 2 * k), idesc,
                                      (!first || k != 0) ? 1u : 0u);
                        umma_commit(&sm.empty[stage]);
                        if (tap == 8) umma_commit(&sm.a_empty[a_slot]);
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
            if (lane == 0) umma_commit(&sm.tmem_full[buf]);  // accumulator complete -> epilogue
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int quarter = warp % 4;  // TMEM lanes [32*quarter, 32*quarter+32) are the ones this warp may read
        if (p.pdl) grid_dep_wait();    // the residual rows are the previous layers' output
        for (int item = blockIdx.x, local = 0; item < num_items; item += gridDim.x, local++) {
            const int buf = local & 1;
            const uint32_t buf_phase = (local >> 1) & 1;
            const int tile = item / p.n_split, n0 = (item % p.n_split) * n_eff;
            conv_epilogue_tile(p, sm.bias, tile, quarter, lane, tmem_base + buf * acc_stride, &sm.tmem_full[buf], buf_phase, &sm.tmem_empty[buf], n0,
                               min(n_eff, p.n_store - n0));
        }
    }


    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(p.tmem_cols))
                     : "memory");
    }
}

}  // namespace

size_t conv_tch_smem_bytes(int n, int stages, int a_rows) {
    return 1024 /*alignment slack*/ + size_t(stages) * n * 128 + kASlots * a_slot_bytes(a_rows) + (2 * stages + 2 * kASlots + 4) * 8 + 16 +
           size_t(n) * 4;
}

int conv_tch_pick_stages(int n, int a_rows) {
    const size_t budget = 227 * 1024;
    int stages = 9;
    while (stages > 2 && conv_tch_smem_bytes(n, stages, a_rows) > budget) stages--;
    return stages;
}

void conv_tch_prepare() { cudaFuncSetAttribute(conv_tch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }

// tmap_a: 2-D map over the channels-last rows with an UNSWIZZLED box of (8 channels, p.a_rows rows); tmap_b: the weight map
// whose box holds p.n / p.n_split rows
void launch_conv_tch(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const ConvTcParams& p, int grid, cudaStream_t s) {
    if (p.num_tiles <= 0) return;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(std::min(grid, p.num_tiles * p.n_split)));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = conv_tch_smem_bytes(p.n, p.stages, p.a_rows);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = p.pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, conv_tch_kernel, tmap_a, tmap_b, p);
}

}  // namespace kzb
