// NetSpec: the AlphaZero ResNet (tower + heads) recognised inside an ONNX graph, with every
// BatchNorm folded away.  This is the B200 build's replacement for kn-graph's `optimize_graph`
// + kn-cuda-eval's planner for this one architecture (call sites:
// rust/kz-selfplay/src/server/server_alphazero.rs:126-128, rust/kz-core/src/network/cudnn.rs:29-43).
//
// Architecture recognised (python/lib/model/post_act.py):
//   tower   :201-228   x0 = conv3x3(in)+b ; D x [ x <- x + relu(bn(conv3x3(relu(bn(conv3x3(x)))))) ] ; t = bn(x)
//   scalars :10-23     conv1x1(t)->relu->flatten->fc->relu->fc(5)
//   policy  :54-112    conv1x1(t)->relu->conv1x1(Pc) then Flatten [+ Gather(const) | ++zeros | ++(conv1x1(t,1)->flatten->fc)]
//   policy  :115-141   attention head: conv1x1 "bulk"/"under" -> slices/reshapes -> bmm(q_from^T, q_to)/sqrt(Q) -> Gather(const)
// BN folding: a = gamma/sqrt(var+eps), b = beta - a*mean;
//   conv followed by BN:   W' = a (.) W,  bias' = a (.) bias + b
//   BN followed by 1x1 conv (final BN into each head conv): W' = W diag(a), bias' = bias + W b   (exact, no padding)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "onnx_reader.hpp"

namespace kzb {

struct ConvParams {
    int cin = 0, cout = 0, ksize = 0;  // ksize 1 or 3, stride 1, pad ksize/2
    std::vector<float> w;              // [cout][cin][k][k]  (ONNX order)
    std::vector<float> b;              // [cout]
};

struct FcParams {
    int in = 0, out = 0;
    std::vector<float> w;  // [out][in]
    std::vector<float> b;  // [out]
};

// Where policy logit i comes from.
enum : int32_t {
    kPolicySrcZero = -1,   // constant 0 (ataxx "pass" logit, post_act.py:108-110)
    kPolicySrcExtra = -2,  // -2 - e : e-th output of the extra fc (go pass move, post_act.py:63-67)
};

struct NetSpec {
    int cin = 0, board_h = 0, board_w = 0;
    int channels = 0, depth = 0;
    ConvParams first;                // conv3x3, bias, no bn, no relu
    std::vector<ConvParams> blocks;  // 2*depth conv3x3 with BN folded; relu after each

    ConvParams scalar_conv;  // 1x1 C->hc (+relu), final BN folded in
    FcParams fc1, fc2;       // fc1: hc*A -> hs (+relu); fc2: hs -> 5

    ConvParams policy_conv1;  // 1x1 C->Cp (+relu), final BN folded in
    ConvParams policy_conv2;  // 1x1 Cp->Pc
    bool has_extra = false;
    ConvParams extra_conv;  // 1x1 C->1, final BN folded in (no relu)
    FcParams extra_fc;      // A -> E

    // attention policy head (post_act.py:115-141): logit[i] = (sum_q E[a.sq][a.chan + q*a.chan_stride] *
    // E[b.sq][b.chan + q*b.chan_stride]) / att_div, where E[sq][chan] is the concatenation (channel offsets
    // att_chan_base, each conv padded to a multiple of 16 channels) of the 1x1 convs in att_convs over the tower output.
    struct AttOperand {
        int32_t chan, chan_stride, sq;
    };
    struct AttEntry {
        AttOperand a, b;
    };
    bool has_attention = false;
    int att_q = 0;
    float att_div = 1.0f;
    std::vector<ConvParams> att_convs;  // final BN folded in, no relu
    std::vector<int> att_chan_base;     // [att_convs.size() + 1]
    std::vector<AttEntry> att_entries;  // [policy_len]

    int policy_len = 0;
    // policy_src[i] >= 0: pc * A + sq into policy_conv2's output (channel-major, as Flatten sees NCHW)
    std::vector<int32_t> policy_src;
    std::vector<int64_t> policy_shape;  // as declared by the graph output (may be 1-D or 3-D)

    int area() const { return board_h * board_w; }
    double flops_per_position() const;
};

// Throws std::runtime_error describing the first thing that does not match.
NetSpec build_net_spec(const OnnxGraph& g);

// The same NetSpec from weights the caller already holds (kzb_net_create, include/kzb200.h): convs with their in-block BatchNorm
// already folded, the tower's trailing BatchNormalization as a per-channel affine y = scale * x + shift (folded into the head
// convs here, like build_net_spec does).  Checks every shape and throws on the first mismatch.
struct RawConv {
    int cin, cout, ksize;
    const float *w, *b;
};
struct RawFc {
    int in, out;
    const float *w, *b;
};
struct RawNet {
    int cin, board_h, board_w, channels, depth;
    RawConv first;
    const RawConv* blocks;  // 2 * depth
    const float *final_scale, *final_shift;  // [channels] or both null
    RawConv scalar_conv;
    RawFc fc1, fc2;
    RawConv policy_conv1, policy_conv2;
    bool has_extra;
    RawConv extra_conv;
    RawFc extra_fc;
    int policy_len;
    const int32_t* policy_src;  // [policy_len]
};
NetSpec net_spec_from_raw(const RawNet& r);

}  // namespace kzb
