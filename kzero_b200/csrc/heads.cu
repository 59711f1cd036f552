// K3: everything after the head convolutions, one warp per position, warp-shuffle reductions.
//
// Replaces, for this path:
//   - the scalar head's Flatten/Gemm/Relu/Gemm steps (python/lib/model/post_act.py:15-19), which the
//     reference executes as separate cuBLAS launches,
//   - the policy Flatten/Gather/Concat plumbing (post_act.py:76-88, 102-112) and the attention head's
//     slice / reshape / bmm / scale / Gather chain (post_act.py:130-141), evaluated only for the requested indices,
//   - and, in packed mode, the CPU-side decode_output (rust/kz-core/src/network/common.rs:16-100):
//     value = tanh(s0), wdl = softmax(s1..3), moves_left = s4, policy = softmax over the LEGAL moves only
//     (gathered by the host-supplied move_to_index list), so only n_legal probabilities per position cross
//     PCIe instead of the full [B, P] logit matrix.
// HBM/latency-bound: algorithmic bytes per position = hc*A*4 + (n_legal or P)*4 read, (5 + n_legal or P)*4 written.
#include "kernels.cuh"

namespace kzb {
namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kMaxHiddenPerLane = 4;  // hs <= 128

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <bool PACKED, int HPL>  // HPL = hidden units per lane = ceil(hs / 32)
__global__ void __launch_bounds__(kWarpsPerBlock * 32) heads_tail_kernel(HeadsTailParams p) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int b = blockIdx.x * kWarpsPerBlock + warp;
    if (b >= p.batch) return;  // whole warp exits together, no block-level sync below

    const int area = p.lay.W * p.lay.H;
    const int n_in = p.hc * area;
    float* s_in = smem + size_t(warp) * n_in;  // flatten order (channel, y, x), post_act.py:16

    for (int i = lane; i < n_in; i += 32) {
        int c = i / area, sq = i % area;
        s_in[i] = p.s1[size_t(p.lay.row(b, sq)) * p.s1_stride + c];
    }
    float extra = 0.0f;
    if (p.extra_w) {
        float e = 0.0f;
        for (int sq = lane; sq < area; sq += 32) e += p.extra_w[sq] * p.s1[size_t(p.lay.row(b, sq)) * p.s1_stride + p.hc];
        extra = warp_sum(e) + p.extra_b;
    }
    __syncwarp();

    // fc1 + relu: lane owns hidden units lane, lane+32, ...; 8 inputs per step so 8*HPL independent
    // weight loads are in flight (the loop is latency-bound otherwise); accumulation order stays i-ascending
    float h[HPL];
#pragma unroll
    for (int u = 0; u < HPL; u++) h[u] = 0.0f;
    for (int i0 = 0; i0 < n_in; i0 += 8) {
        float w[8][HPL], v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = i0 + k;
            const bool ok = i < n_in;
            v[k] = ok ? s_in[i] : 0.0f;
#pragma unroll
            for (int u = 0; u < HPL; u++) {
                const int j = lane + 32 * u;
                w[k][u] = (ok && j < p.hs) ? __ldg(p.fc1_t + size_t(i) * p.hs + j) : 0.0f;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++)
#pragma unroll
            for (int u = 0; u < HPL; u++) h[u] = fmaf(w[k][u], v[k], h[u]);
    }
#pragma unroll
    for (int u = 0; u < HPL; u++) {
        const int j = lane + 32 * u;
        float t = j < p.hs ? h[u] + p.fc1_b[j] : 0.0f;
        h[u] = t < 0.0f ? 0.0f : t;
    }
    // fc2: 5 outputs
    float sc[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        float part = 0.0f;
#pragma unroll
        for (int u = 0; u < HPL; u++) {
            int j = lane + 32 * u;
            if (j < p.hs) part = fmaf(p.fc2_w[k * p.hs + j], h[u], part);
        }
        sc[k] = warp_sum(part) + p.fc2_b[k];
    }

    auto logit = [&](int i) -> float {
        if (p.att_entries) {  // attention head: q_from(from square) . q_to(to square / under-promotion slot) / sqrt(Q)
            const AttEntryDev e = p.att_entries[i];
            const float* a = p.att + size_t(p.lay.row(b, e.a_sq)) * p.att_stride + e.a_chan;
            const float* c = p.att + size_t(p.lay.row(b, e.b_sq)) * p.att_stride + e.b_chan;
            float acc = 0.0f;
            for (int q = 0; q < p.att_q; q++) acc = fmaf(a[q * e.a_stride], c[q * e.b_stride], acc);
            return acc / p.att_div;
        }
        int src = p.policy_src[i];
        if (src >= 0) return p.pm[size_t(p.lay.row(b, src % area)) * p.pm_stride + src / area];
        return src == -1 ? 0.0f : extra;
    };

    if (!PACKED) {
        if (lane < 5) p.out_scalars[size_t(b) * 5 + lane] = sc[lane];
        for (int i = lane; i < p.policy_len; i += 32) p.out_logits[size_t(b) * p.policy_len + i] = logit(i);
        return;
    }

    // decode_output, common.rs:59-74
    if (lane == 0) {
        float* ov = p.out_values + size_t(b) * 5;
        ov[0] = tanhf(sc[0]);
        float mx = fmaxf(sc[1], fmaxf(sc[2], sc[3]));
        float e0 = expf(sc[1] - mx), e1 = expf(sc[2] - mx), e2 = expf(sc[3] - mx);
        float sum = (e0 + e1) + e2;
        if (!(sum > 0.0f)) *reinterpret_cast<volatile int*>(p.err_flag) = 1 + b;
        ov[1] = e0 / sum;
        ov[2] = e1 / sum;
        ov[3] = e2 / sum;
        ov[4] = sc[4];
    }
    // masked softmax over the legal moves only, common.rs:76-86 + softmax_in_place :102-114
    const uint32_t o0 = p.mv_off[b], o1 = p.mv_off[b + 1];
    const int n = int(o1 - o0);
    if (n <= 0) return;  // terminal board: empty policy (common.rs:77)
    const int32_t* pmap = p.sym ? p.policy_map + size_t(p.sym[b]) * p.policy_len : nullptr;
    constexpr int kRegMoves = 12;  // up to 384 legal moves stay in registers (chess <= 218, go-19 <= 362)
    if (n <= 32 * kRegMoves) {
        float l[kRegMoves];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < kRegMoves; k++) {
            const int j = lane + 32 * k;
            l[k] = -INFINITY;
            if (j < n) {
                uint32_t idx = p.mv_idx[o0 + j];
                if (pmap && idx < uint32_t(p.policy_len)) idx = uint32_t(pmap[idx]);
                l[k] = idx < uint32_t(p.policy_len) ? logit(int(idx)) : NAN;
                mx = fmaxf(mx, l[k]);
            }
        }
        mx = warp_max(mx);
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < kRegMoves; k++) {
            const int j = lane + 32 * k;
            if (j < n) {
                l[k] = expf(l[k] - mx);
                sum += l[k];
            }
        }
        sum = warp_sum(sum);
        if (!(sum > 0.0f)) {  // NaN logits: the reference panics here (common.rs:110)
            if (lane == 0) *reinterpret_cast<volatile int*>(p.err_flag) = 1 + b;
        }
#pragma unroll
        for (int k = 0; k < kRegMoves; k++) {
            const int j = lane + 32 * k;
            if (j < n) p.out_probs[o0 + j] = l[k] / sum;
        }
        return;
    }
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) {
        uint32_t idx = p.mv_idx[o0 + j];
        if (pmap && idx < uint32_t(p.policy_len)) idx = uint32_t(pmap[idx]);
        float l = idx < uint32_t(p.policy_len) ? logit(int(idx)) : NAN;
        p.out_probs[o0 + j] = l;  // stage the gathered logit; each lane re-reads only its own entries
        mx = fmaxf(mx, l);
    }
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int j = lane; j < n; j += 32) {
        float e = expf(p.out_probs[o0 + j] - mx);
        p.out_probs[o0 + j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (!(sum > 0.0f)) {
        if (lane == 0) *reinterpret_cast<volatile int*>(p.err_flag) = 1 + b;
    }
    for (int j = lane; j < n; j += 32) p.out_probs[o0 + j] /= sum;
}

}  // namespace

void launch_heads_tail(const HeadsTailParams& p, bool packed, cudaStream_t s) {
    if (p.batch <= 0) return;
    int blocks = (p.batch + kWarpsPerBlock - 1) / kWarpsPerBlock;
    size_t smem = size_t(kWarpsPerBlock) * p.hc * p.lay.W * p.lay.H * sizeof(float);
    const int hpl = (p.hs + 31) / 32;
#define KZB_TAIL(PACKED, HPL) heads_tail_kernel<PACKED, HPL><<<blocks, kWarpsPerBlock * 32, smem, s>>>(p)
    if (packed) {
        if (hpl <= 1) KZB_TAIL(true, 1);
        else if (hpl == 2) KZB_TAIL(true, 2);
        else KZB_TAIL(true, 4);
    } else {
        if (hpl <= 1) KZB_TAIL(false, 1);
        else if (hpl == 2) KZB_TAIL(false, 2);
        else KZB_TAIL(false, 4);
    }
#undef KZB_TAIL
}

}  // namespace kzb
