// Device-side parameter blocks and host launchers for the hot path's kernels (sm_100a only).
//   K2  encode      packed bitboards + scalars -> input planes          (encode.cu)      HBM-bound
//   K1  conv tower  implicit-GEMM conv3x3/1x1 on tcgen05 + TMA + TMEM   (conv_tc.cu)     tensor-bound
//       conv fp32   CUDA-core fp32 implicit GEMM, the <=1e-4 parity mode (conv_fp32.cu)
//   K3  heads tail  scalar-head FCs, policy gather, masked softmax      (heads.cu)       HBM/latency-bound
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

namespace kzb {

// Row index of square (y, x) of board b inside an activation matrix [rows][channels]:
//   row = b * board_pitch + y * rank_pitch + x
// dense  layout: rank_pitch = W,   board_pitch = H*W        (no padding rows)
// padded layout: rank_pitch = W+1, board_pitch = (H+1)*(W+1): one zero column after every rank and one
//   zero rank after every board, so every 3x3 tap is a pure row shift dy*rank_pitch+dx that lands on a
//   zero row whenever it leaves the board.
struct RowLayout {
    int W, H, rank_pitch, board_pitch;
    __host__ __device__ int row(int b, int sq) const { return b * board_pitch + (sq / W) * rank_pitch + (sq % W); }
    __host__ __device__ bool padded() const { return rank_pitch != W; }
};

// ---------------------------------------------------------------------------------------------- K2
struct EncodeParams {
    const uint8_t* bits;   // [batch][bits_stride]   LSB-first (bit_buffer.rs:27-35)
    const float* scalars;  // [batch][scalar_count]
    int batch, bits_stride, scalar_count, bool_channels;
    RowLayout lay;
    int c_pad;  // channels per row in the output (>= scalar_count + bool_channels, multiple of 8)
    void* out;  // [batch*board_pitch][c_pad] bf16 or f32
    // optional board symmetry (RandomSymmetryNetwork on the GPU, network/symmetry.rs:41-67): plane square sq of board b is
    // read from square square_src[sym[b]*A + sq] of the record; null = identity
    const uint8_t* sym;
    const int32_t* square_src;
    // k-chunk-major output only: the record's own board (<= 8x8), embedded top-left in the 8x8 grid the tower works on
    int rec_w, rec_h;
};
void launch_encode_nhwc(const EncodeParams& p, bool out_bf16, cudaStream_t s);
// exact twin of InputMapper::encode_input_full (mapping/mod.rs:40-63): out [batch][Cs+Cb][H*W] f32
void launch_encode_nchw_f32(const uint8_t* bits, const float* scalars, int batch, int bits_stride, int scalar_count,
                            int bool_channels, int area, float* out, cudaStream_t s);
// f32 NCHW [batch][C][H*W] -> rows [batch*board_pitch][c_pad] (bf16 or f32), zero pad rows / channels
void launch_nchw_to_rows(const float* in, int batch, int channels, RowLayout lay, int c_pad, void* out, bool out_bf16,
                         cudaStream_t s);

// ---------------------------------------------------------------------------------------------- conv fp32
struct ConvF32Params {
    const float* in;  // [rows][in_stride]
    int in_stride;
    const float* w;     // [taps][cin][cout]
    const float* bias;  // [cout]
    const float* res;   // optional [rows][res_stride], added AFTER the relu (post_act.py:227-228)
    int res_stride;
    float* out;  // [rows][out_stride]
    int out_stride;
    int cin, cout, taps;  // taps 1 or 9
    int relu_n;           // relu on output channels < relu_n
    int batch;
    RowLayout lay;  // dense
};
void launch_conv_fp32(const ConvF32Params& p, cudaStream_t s);

// ---------------------------------------------------------------------------------------------- K1
struct ConvTcParams {
    int num_tiles;  // 128-row M tiles
    int taps;       // 1 or 9
    int kblocks;    // cin_pad / 64
    int cin_pad;
    int n;     // UMMA N = padded cout, multiple of 16, <= 256
    int mode;  // 0: rows are a padded linear layout, 2-D TMA; 1: dense 8x8 boards, 4-D TMA box (c,x,y,b)
    int boards_per_tile;
    RowLayout lay;
    int valid_rows;  // rows >= valid_rows are not stored
    const float* bias;  // [n]
    int relu_n;
    const __nv_bfloat16* res;  // optional [rows][res_stride]
    int res_stride;
    void* out;  // [rows][out_stride] bf16 or f32
    int out_stride;
    int out_f32;
    int n_store;  // channels written per row (multiple of 16, <= n)
    int stages;
    int tmem_cols;
    int n_split;  // conv_i2c: 1, or 2 = a work item is one 256-pixel tile x one half of the output channels
    int pdl;      // conv_i2c: launched with programmatic stream serialization (set-up and the first weight tiles overlap the previous layer's tail)
    // development aid: when non-null, each CTA writes 16 clock64() stamps (see conv_tc8.cu) -- KZB_TIMELINE=1
    unsigned long long* timeline;
};
void launch_conv_tc(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const ConvTcParams& p, int grid, cudaStream_t s);
// conv_i2c.cu: conv3x3 / conv1x1 (p.taps) on DENSE rows, activation tiles by TMA im2col, CTA-pair MMA (tmap_bh: weight box of p.n / p.n_split / 2 rows)
void launch_conv_i2c(const CUtensorMap& tmap_a_im2col, const CUtensorMap& tmap_bh, const CUtensorMap& tmap_out, const CUtensorMap& tmap_res,
                     const ConvTcParams& p, int grid, cudaStream_t s);
size_t conv_i2c_smem_bytes(int n, int stages);
int conv_i2c_pick_stages(int n);
void conv_i2c_prepare();
size_t conv_tc_smem_bytes(int n, int stages);
int conv_tc_pick_stages(int n);
void conv_tc_prepare();  // per-device: opt in to 227 KB dynamic shared memory

// 8x8-board specialisation (conv_tc8.cu): p.num_tiles counts 4-board work units, p.stages is the number
// of weight-ring slots, tmap_a is the (c, x, board, y)-ordered map with box (64, 8, 4, 10).
void launch_conv_tc8(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const ConvTcParams& p, int grid,
                     cudaStream_t s);
size_t conv_tc8_smem_bytes(int n, int b_slots);
int conv_tc8_pick_b_slots(int n);
void conv_tc8_prepare();

// whole-tower persistent kernel for 8x8 boards (tower8k.cu)
struct TowerLayerDev {
    int a_map;   // input activations: 0 = encoded planes, 1 = X (residual stream), 2 = T (block-internal)
    int w_map;   // 0 = first-layer weights, 1 = concatenated block weights
    int w_row0;  // first row of this layer inside its weight matrix
    int kblocks, cin_pad;
    int relu_n;
    int has_res;  // add X (same rows) after the relu
    int out_buf;  // 1 = X, 2 = T
    const float* bias;
    // second-generation kernel (tower8k.cu)
    int kchunks;       // 8-channel chunks staged per k-block (8, fewer for a narrow first layer)
    int ksteps;        // K=16 MMA steps per k-block and tap (kchunks / 2)
    int out_rowmajor;  // last layer: store [position][C] rows (what the head kernels read) instead of k-chunk-major
};
struct Tower8Params {
    int num_layers;
    const TowerLayerDev* layers;  // device memory
    int num_units;                // 4-board work units
    int valid_rows;
    int n, n_store;
    __nv_bfloat16* x;
    __nv_bfloat16* t;
    __nv_bfloat16* xt;  // channel-major copy of the residual stream: [unit][128 channels][256 positions]
    int stride;  // elements per row of X / T
    int b_slots, tmem_cols;
    int cluster;  // 1: every CTA streams its own weight tiles; 2: CTA pairs, each loads half of every tile and multicasts it
    // tower8k only: the real board inside the 8x8 grid (ataxx 7x7, ...): outputs on squares outside it are forced to zero after
    // every layer, so that they keep acting as the convolution's zero padding
    int board_w, board_h;
    // tower8k only: balanced board assignment (CTA c owns bal_base + (c < bal_rem) contiguous boards as two units of 4 / 3)
    int balanced, bal_base, bal_rem, bal_grid;
    int pdl;  // launched with programmatic stream serialization: set-up overlaps the encode kernel, which runs right before it
    unsigned long long* timeline;
    int debug;  // development aid (KZB_DEBUG): 1 = skip A loads, 2 = skip B loads, 4 = skip epilogue memory traffic
};
// second generation (tower8k.cu): k-chunk-major activations A[kc][board][y][x][8], every tap = a descriptor offset
struct Tower8kMaps {
    // second index: boards per unit - 3 (units of 3 or 4 boards)
    CUtensorMap a[3][2];    // loads: encoded planes, X, T -- dims (x*8+c8: 64, board, y: 8, kc), box (72, nb, 9, 1), no swizzle
    CUtensorMap w[2];       // loads: first-layer weights, concatenated block weights -- box (64, n / cluster), SWIZZLE_128B
    CUtensorMap out[3][2];  // stores: X, T k-chunk-major -- dims (64, kc, board, y), box (80, 4, nb, 1); [2] = X row-major (c, x, board, y), box (32, 8, nb, 1)
};
void launch_tower8k(const Tower8kMaps& maps, const Tower8Params& p, int grid, cudaStream_t s);
size_t tower8k_smem_bytes(int w_slots);
int tower8k_pick_b_slots();
int tower8k_max_local_units();
void tower8k_prepare();

// K2 / layout twins for the k-chunk-major tower input: out[kc][boards_total][64 squares][8 channels] bf16
void launch_encode_kc(const EncodeParams& p, int kc_total, int boards_total, cudaStream_t s);
void launch_nchw_to_kc(const float* in, int batch, int channels, int rec_w, int rec_h, int kc_total, int boards_total, void* out,
                       cudaStream_t s);

// ---------------------------------------------------------------------------------------------- K3
struct AttEntryDev {  // same layout as NetSpec::AttEntry
    int32_t a_chan, a_stride, a_sq, b_chan, b_stride, b_sq;
};
struct HeadsTailParams {
    int batch;
    RowLayout lay;
    const float* s1;  // [rows][s1_stride]: ch 0..hc-1 relu(scalar conv), ch hc = extra policy conv (no relu)
    int s1_stride, hc;
    const float* pm;  // [rows][pm_stride]: policy conv map, channel pc at [row][pc]
    int pm_stride;
    const float* fc1_t;  // [hc*A][hs]  (transposed fc1 weight)
    const float* fc1_b;  // [hs]
    const float* fc2_w;  // [5][hs]
    const float* fc2_b;  // [5]
    int hs;
    const float* extra_w;  // [A] or null
    float extra_b;
    const int32_t* policy_src;  // [P]; null for the attention head
    int policy_len;
    // attention policy head (post_act.py:115-141): logit[i] = dot_q(E[a.sq][a.chan + q*a.stride], E[b.sq][b.chan + q*b.stride]) / att_div
    const float* att;  // [rows][att_stride]: concatenated outputs of the head's 1x1 convs
    int att_stride, att_q;
    float att_div;
    const AttEntryDev* att_entries;  // [P] or null
    // planes mode (twin of CudaExecutor::evaluate, network/cudnn.rs:73)
    float* out_scalars;  // [batch][5] raw
    float* out_logits;   // [batch][P]
    // packed mode (fused decode_output, network/common.rs:16-100)
    const uint32_t* mv_idx;
    const uint32_t* mv_off;  // [batch+1]
    // optional symmetry: legal index i of board b is looked up at policy_map[sym[b]*P + i] (unmap_eval, symmetry.rs:126-148)
    const uint8_t* sym;
    const int32_t* policy_map;
    float* out_values;       // [batch][5]: tanh(v), softmax(wdl), moves_left
    float* out_probs;        // CSR-aligned with mv_idx
    int* err_flag;           // set to 1+board (any failing board) when a softmax sum is not > 0 (common.rs:110); may be host memory
};
void launch_heads_tail(const HeadsTailParams& p, bool packed, cudaStream_t s);

// ---------------------------------------------------------------------------------------------- K1' + K3 fused (heads8.cu)
struct Heads8Maps {
    CUtensorMap x;   // tower output rows [rows][c_pad] bf16, box (64, 128), SWIZZLE_128B
    CUtensorMap w1;  // policy conv1 [n1][c_pad], box (64, n1)
    CUtensorMap w2;  // policy conv2 [n2][n1],    box (64, n2)
    CUtensorMap ws;  // scalar conv  [16][c_pad], box (64, 16)
};
struct Heads8Params {
    int num_tiles;  // 128-row tiles = pairs of 8x8 boards
    int batch;
    int kblocks;    // c_pad / 64
    int n1, n2;     // padded output channels of policy conv1 (multiple of 64) / conv2 (multiple of 16)
    int pc;         // real output channels of policy conv2
    int hc, hs;     // scalar head: conv channels, hidden size (<= 32)
    const float *b1, *b2, *bs;  // [n1], [n2], [16]
    const float* fc1_t;         // [hc*64][hs]
    const float* fc1_b;         // [hs]
    const float* fc2_w;         // [5][hs]
    const float* fc2_b;         // [5]
    const int32_t* policy_src;  // [P]: pc*64 + sq, or -1 for a constant-zero logit
    int policy_len;
    int packed;
    const uint32_t* mv_idx;
    const uint32_t* mv_off;
    const uint8_t* sym;         // optional symmetry, as in HeadsTailParams
    const int32_t* policy_map;
    float* out_values;
    float* out_probs;
    int* err_flag;
    float* out_scalars;
    float* out_logits;
    int pdl;  // launched with programmatic stream serialization: set-up and the resident weights' loads overlap the tower's tail
    unsigned long long* timeline;  // development aid (KZB_TIMELINE=heads8): per-CTA clock64() stamps
};
size_t heads8_smem_bytes(const Heads8Params& p);
bool heads8_supported(const Heads8Params& p);
void heads8_prepare();
void launch_heads8(const Heads8Maps& maps, const Heads8Params& p, int grid, cudaStream_t s);

}  // namespace kzb
