/*
 * kz_oracle.c -- CPU restatement of kZero's self-play inference hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (kzero_b200/csrc, libkzb200.so) never links, loads or calls anything in oracle/.
 *
 * What it restates (all citations relative to /root/reference):
 *   - plane expansion:  rust/kz-core/src/mapping/mod.rs:40-63 (encode_input_full)
 *                       rust/kz-core/src/mapping/bit_buffer.rs:27-35,51-55,73-75 (LSB-first bit order)
 *   - network arithmetic: the nn-graph CPU executor (kn-graph 0.7.3, crates.io, NOT vendored in the
 *     reference tree; call site rust/kz-core/src/network/cpu.rs:50).  Its published algorithm is
 *     straight-loop f32 evaluation of the ONNX graph; the ops restated here are the ones the
 *     reference nets contain (python/lib/model/post_act.py:10-23,54-141,187-239).
 *   - output decode:    rust/kz-core/src/network/common.rs:16-114 (tanh, wdl softmax,
 *                       legal-move gather + softmax_in_place with sequential f32 sum)
 *
 * Parity pinning: plane expansion is pinned by the reference's own bit_buffer unit tests
 * (bit_buffer.rs:112-164) and by python/lib/data/position.py:94-98,267-271 (imported when the
 * golden fixtures are generated).  Network numerics are pinned against PyTorch fp32 running the
 * reference's own model classes (the reference's own cross-check mechanism,
 * python/lib/save_onnx.py:94-102); fixtures under tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KZO_API __attribute__((visibility("default")))

KZO_API int kzo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

KZO_API void kzo_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n < 1 ? 1 : n);
#else
    (void)n;
#endif
}

/* mapping/mod.rs:40-63: for each board, scalar planes first (each scalar repeated w*h times),
 * then one f32 per bool, bool i = bit (i%8) of byte (i/8) (bit_buffer.rs:73-75, Index impl :79-89).
 * bits:    n * bits_stride bytes, bits_stride = ceil(bool_count/8) (BitBuffer::new, bit_buffer.rs:11-17)
 * scalars: n * scalar_count f32
 * out:     n * (scalar_count*area + bool_count) f32  == [n, Cs+Cb, H, W] NCHW                      */
KZO_API void kzo_expand_planes(const uint8_t *bits, const float *scalars, int64_t n, int64_t bool_count,
                               int64_t scalar_count, int64_t area, float *out) {
    int64_t bits_stride = (bool_count + 7) / 8;
    int64_t full_len = scalar_count * area + bool_count;
    for (int64_t b = 0; b < n; b++) {
        float *o = out + b * full_len;
        const uint8_t *bb = bits + b * bits_stride;
        const float *ss = scalars + b * scalar_count;
        for (int64_t s = 0; s < scalar_count; s++)
            for (int64_t a = 0; a < area; a++)
                *o++ = ss[s];
        for (int64_t i = 0; i < bool_count; i++)
            *o++ = (float)((bb[i / 8] >> (i % 8)) & 1);
    }
}

/* ONNX Conv, group 1, stride 1, dilation 1, symmetric zero padding `pad`, square kernel k.
 * x [N,Ci,H,W], w [Co,Ci,k,k], bias [Co] or NULL, y [N,Co,H,W].  Straight f32 loops with
 * the accumulation order (ci, ky, kx); relu fused only when asked (the graph has it separate). */
KZO_API void kzo_conv2d(const float *x, const float *w, const float *bias, float *y, int64_t N, int64_t Ci,
                        int64_t H, int64_t W, int64_t Co, int64_t k, int64_t pad, int relu) {
    int64_t Ho = H + 2 * pad - k + 1, Wo = W + 2 * pad - k + 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t n = 0; n < N; n++) {
        for (int64_t co = 0; co < Co; co++) {
            float *yo = y + (n * Co + co) * Ho * Wo;
            float b0 = bias ? bias[co] : 0.0f;
            for (int64_t i = 0; i < Ho * Wo; i++)
                yo[i] = b0;
            for (int64_t ci = 0; ci < Ci; ci++) {
                const float *xi = x + (n * Ci + ci) * H * W;
                const float *wk = w + (co * Ci + ci) * k * k;
                for (int64_t ky = 0; ky < k; ky++) {
                    for (int64_t kx = 0; kx < k; kx++) {
                        float wv = wk[ky * k + kx];
                        int64_t oy0 = pad - ky > 0 ? pad - ky : 0;
                        int64_t oy1 = H + pad - ky < Ho ? H + pad - ky : Ho;
                        int64_t ox0 = pad - kx > 0 ? pad - kx : 0;
                        int64_t ox1 = W + pad - kx < Wo ? W + pad - kx : Wo;
                        for (int64_t oy = oy0; oy < oy1; oy++) {
                            const float *xr = xi + (oy + ky - pad) * W + (kx - pad);
                            float *yr = yo + oy * Wo;
                            for (int64_t ox = ox0; ox < ox1; ox++)
                                yr[ox] += wv * xr[ox];
                        }
                    }
                }
            }
            if (relu)
                for (int64_t i = 0; i < Ho * Wo; i++)
                    yo[i] = yo[i] > 0.0f ? yo[i] : 0.0f;
        }
    }
}

/* ONNX Gemm: y[M,N] = alpha * x[M,K] * (transB ? w[N,K]^T : w[K,N]) + beta * c[N] */
KZO_API void kzo_gemm(const float *x, const float *w, const float *c, float *y, int64_t M, int64_t K, int64_t N,
                      int transB, float alpha, float beta) {
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < M; m++) {
        for (int64_t n = 0; n < N; n++) {
            float acc = 0.0f;
            for (int64_t kk = 0; kk < K; kk++)
                acc += x[m * K + kk] * (transB ? w[n * K + kk] : w[kk * N + n]);
            y[m * N + n] = alpha * acc + (c ? beta * c[n] : 0.0f);
        }
    }
}

/* network/common.rs:102-114 softmax_in_place: max-fold from -inf, exp(v-max), sequential f32 sum,
 * assert sum > 0 (returns -1 instead of panicking), divide. */
static int softmax_in_place(float *v, int64_t n) {
    float mx = -INFINITY;
    for (int64_t i = 0; i < n; i++)
        mx = fmaxf(mx, v[i]); /* f32::max: NaN-ignoring like fmaxf */
    float sum = 0.0f;
    for (int64_t i = 0; i < n; i++) {
        v[i] = expf(v[i] - mx);
        sum += v[i];
    }
    if (!(sum > 0.0f))
        return -1;
    for (int64_t i = 0; i < n; i++)
        v[i] /= sum;
    return 0;
}

/* network/common.rs:16-100 decode_output, 2-output form (scalars [B,5], policy [B,P]).
 * mv_idx/mv_off: CSR list of move_to_index(available_moves) per board (binary_output.rs:299-315).
 * out_values [B,5] = value(tanh), wdl w/d/l (softmax), moves_left; out_policy CSR-aligned.
 * A board with zero moves (terminal, common.rs:77 map_or(vec![])) gets an empty policy.
 * Returns 0, or -(1+board) when the reference would have panicked in the softmax assert. */
KZO_API int64_t kzo_decode_output(const float *scalars, const float *policy_logits, int64_t B, int64_t P,
                                  const uint32_t *mv_idx, const uint32_t *mv_off, float *out_values,
                                  float *out_policy) {
    for (int64_t b = 0; b < B; b++) {
        const float *s = scalars + b * 5;
        float *ov = out_values + b * 5;
        ov[0] = tanhf(s[0]);
        float wdl[3] = {s[1], s[2], s[3]};
        if (softmax_in_place(wdl, 3))
            return -(1 + b);
        ov[1] = wdl[0];
        ov[2] = wdl[1];
        ov[3] = wdl[2];
        ov[4] = s[4];
        int64_t o0 = mv_off[b], o1 = mv_off[b + 1];
        if (o1 > o0) {
            for (int64_t j = o0; j < o1; j++)
                out_policy[j] = policy_logits[b * P + mv_idx[j]];
            if (softmax_in_place(out_policy + o0, o1 - o0))
                return -(1 + b);
        }
    }
    return 0;
}
