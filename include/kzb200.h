/*
 * kzb200.h -- C ABI of libkzb200.so: B200-native batched evaluation of kZero's AlphaZero ResNet.
 *
 * This is the drop-in boundary for ONE path of the reference: what `CudaNetwork` does between
 * `Network::evaluate_batch` and the GPU (rust/kz-core/src/network/cudnn.rs:18-88).  It replaces the
 * kn-cuda-eval / kn-cuda-sys / cuDNN dependency for that path; everything above the `Network` trait
 * (kz-selfplay executor threads, RandomSymmetryNetwork, MCTS) is an unchanged consumer.
 * Plain pointers and sizes only; all host pointers are owned by the caller and only used during the call.
 *
 * Threading (mirrors `evaluate_batch(&mut self)`, network/mod.rs:52-63): a kzb_net handle is NOT
 * thread-safe; different handles are independent (own stream, own buffers) and may be created, used and
 * destroyed concurrently from different threads, on the same or different devices.
 *
 * Errors: every int-returning function returns 0 on success and non-zero on failure; the message is
 * available (per thread) from kzb_last_error().  The reference panics in the same situations
 * (cudnn.rs:58 batch too large, common.rs:165-198 shape mismatch, common.rs:110 softmax sum not > 0);
 * the Rust shim turns a non-zero return into a panic to preserve that behaviour.
 * There is no CPU fallback: without a CUDA device every entry point that needs one fails.
 */
#ifndef KZB200_H
#define KZB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kzb_net kzb_net;

#define KZB_PRECISION_FP32 0 /* CUDA-core fp32, <= 1e-4 max-abs vs the CPU executor            */
#define KZB_PRECISION_BF16 1 /* tcgen05 tensor cores: bf16 operands, fp32 accumulate (the fast path) */

/* Replaces CudaDevice::all() (rust/kz-selfplay/src/server/server.rs:49-51). */
int kzb_device_count(void);

/* Thread-local message of the last failure on this thread ("" if none). */
const char* kzb_last_error(void);

/* Shapes of a loaded network, for the caller's own checks. */
typedef struct kzb_net_info {
    int32_t input_channels, board_h, board_w; /* graph input [BATCH, C, H, W]                 */
    int32_t policy_len;                       /* product of the policy output's non-batch dims */
    int32_t channels, depth;                  /* tower width / residual blocks                 */
    int32_t max_batch, precision, device;
    int32_t conv_mode;      /* 0: padded-row 2-D TMA im2col, 1: 8x8-board 4-D TMA box, -1: fp32 path */
    double flops_per_position; /* algorithmic FLOPs (no padding/halo), SURVEY.md 8(d)          */
} kzb_net_info;

/* Replaces `CudaNetwork::new(mapper, &graph, max_batch_size, device)` (network/cudnn.rs:29-43) together
 * with `load_graph_from_onnx_path` + `optimize_graph` (rust/kz-selfplay/src/server/server_alphazero.rs:126-128):
 * parses the ONNX bytes, recognises the ResNet tower + heads, folds every BatchNorm, packs weights,
 * allocates device buffers for `max_batch` positions.  Fails with a message for graphs that are not the
 * reference architecture (python/lib/model/post_act.py). */
int kzb_net_create_from_onnx(int device, const void* onnx_bytes, size_t onnx_len, int max_batch, int precision,
                             kzb_net** out);

/* Twin of check_graph_shapes (network/common.rs:165-198): verifies the graph input is
 * [BATCH, scalar_count + bool_channels, board_h, board_w] and the policy has `policy_len` entries, and
 * tells the network how a packed record splits into scalar and bool planes (InputMapper::input_bool_shape /
 * input_scalar_count, rust/kz-core/src/mapping/mod.rs:20-22).  Required before the *_packed / encode calls. */
int kzb_net_bind_mapper(kzb_net* net, int scalar_count, int bool_channels, int board_h, int board_w, int policy_len);

int kzb_net_get_info(const kzb_net* net, kzb_net_info* out);

/* Host-only half of kzb_net_create_from_onnx: parse + recognise + fold, no device touched (max_batch,
 * precision, device, conv_mode are reported as -1).  Lets a caller validate a network file, and lets the
 * host logic be tested on a machine without a GPU. */
int kzb_onnx_inspect(const void* onnx_bytes, size_t onnx_len, kzb_net_info* out);

/* Replaces dropping the CudaNetwork (executor.rs:326-331 drops the old net before loading the next). */
void kzb_net_destroy(kzb_net* net);

/* Exact semantic twin of `CudaExecutor::evaluate(&[DTensor::F32(input)])` (network/cudnn.rs:73):
 * nchw_in [batch, C, H, W] f32 -> out_scalars [batch, 5] raw head outputs, out_policy_logits [batch, policy_len].
 * Computes exactly `batch` rows (1 <= batch <= max_batch); rows are independent of each other. */
int kzb_eval_planes(kzb_net* net, const float* nchw_in, int batch, float* out_scalars, float* out_policy_logits);

/* The fused fast path: everything `CudaNetwork::evaluate_batch` does (cudnn.rs:55-87) in one call.
 *   bits    [batch, ceil(bool_channels*H*W/8)]  BitBuffer::storage() per board (mapping/bit_buffer.rs:68-70)
 *   scalars [batch, scalar_count] f32           as pushed by InputMapper::encode_input (mapping/mod.rs:38)
 *   mv_idx / mv_off: CSR list of PolicyMapper::move_to_index over board.available_moves(), in iteration
 *                    order (same as collect_policy_indices, rust/kz-selfplay/src/binary_output.rs:299-315);
 *                    mv_off has batch+1 entries, mv_off[0] == 0; a terminal board has an empty range.
 *   out_values [batch, 5]: value = tanh(s0), wdl = softmax(s1..s3), moves_left = s4   (network/common.rs:59-74)
 *   out_policy [mv_off[batch]]: softmax over each board's legal moves only            (network/common.rs:76-86)
 * Returns non-zero (and writes no NaN probabilities silently) if any softmax sum is not > 0 (common.rs:110). */
int kzb_eval_packed(kzb_net* net, const uint8_t* bits, const float* scalars, int batch, const uint32_t* mv_idx,
                    const uint32_t* mv_off, float* out_values, float* out_policy);

/* K2 alone, for bit-exact parity checks: twin of InputMapper::encode_input_full (mapping/mod.rs:40-63),
 * out_nchw [batch, scalar_count + bool_channels, H, W] f32, produced on the GPU. */
int kzb_encode_planes(kzb_net* net, const uint8_t* bits, const float* scalars, int batch, float* out_nchw);

/* ---- measurement hooks (bench.py); not part of the reference's interface ------------------------- */

/* Upload one packed batch into the network's device buffers (inputs resident in HBM afterwards). */
int kzb_stage_packed(kzb_net* net, const uint8_t* bits, const float* scalars, int batch, const uint32_t* mv_idx,
                     const uint32_t* mv_off);
/* Run the device side of kzb_eval_packed on the staged batch `iters` times, each iteration timed with
 * CUDA events on the network's own stream; ms_out[iters].  If flush_l2 != 0 a 256 MiB scratch buffer is
 * overwritten before every iteration, outside the timed region. */
int kzb_time_staged(kzb_net* net, int iters, int flush_l2, float* ms_out);
/* Same, but per kernel launch of ONE iteration: names_out receives '\n'-separated step names,
 * ms_out[*n_steps] the CUDA-event duration of each launch.  Returns non-zero if the buffers are too small. */
int kzb_profile_staged(kzb_net* net, int flush_l2, char* names_out, size_t names_cap, float* ms_out, int ms_cap,
                       int* n_steps);
/* Number of kernel launches one kzb_eval_packed call issues. */
int kzb_launches_per_eval(const kzb_net* net);

#ifdef __cplusplus
}
#endif
#endif /* KZB200_H */
