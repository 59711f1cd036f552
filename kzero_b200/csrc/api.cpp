// extern "C" surface of libkzb200.so (include/kzb200.h).  Every entry point catches C++ exceptions and
// turns them into a non-zero return + thread-local message; nothing throws across the ABI.
#include <cuda_runtime.h>

#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/kzb200.h"
#include "executor.hpp"

#define KZB_API extern "C" __attribute__((visibility("default")))

struct kzb_net {
    kzb::Net impl;
    kzb_net(int device, const void* onnx, size_t len, int max_batch, int precision)
        : impl(device, onnx, len, max_batch, precision) {}
    kzb_net(int device, kzb::NetSpec spec, int max_batch, int precision) : impl(device, std::move(spec), max_batch, precision) {}
};

namespace {
template <typename F>
int guarded(F&& f) {
    try {
        f();
        kzb::set_last_error("");
        return 0;
    } catch (const std::exception& e) {
        kzb::set_last_error(e.what());
        return 1;
    } catch (...) {
        kzb::set_last_error("unknown error");
        return 1;
    }
}
void need(const void* p, const char* what) {
    if (!p) throw std::runtime_error(std::string(what) + " must not be NULL");
}
}  // namespace

KZB_API int kzb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

KZB_API const char* kzb_last_error(void) { return kzb::last_error(); }

KZB_API int kzb_net_create_from_onnx(int device, const void* onnx_bytes, size_t onnx_len, int max_batch, int precision,
                                     kzb_net** out) {
    return guarded([&] {
        need(out, "out");
        *out = nullptr;
        need(onnx_bytes, "onnx_bytes");
        *out = new kzb_net(device, onnx_bytes, onnx_len, max_batch, precision);
    });
}

KZB_API int kzb_net_create(int device, const kzb_net_spec* spec, int max_batch, int precision, kzb_net** out) {
    return guarded([&] {
        need(out, "out");
        *out = nullptr;
        need(spec, "spec");
        auto conv = [](const kzb_conv_weights& c) { return kzb::RawConv{c.cin, c.cout, c.ksize, c.w, c.b}; };
        auto fc = [](const kzb_fc_weights& f) { return kzb::RawFc{f.in, f.out, f.w, f.b}; };
        std::vector<kzb::RawConv> blocks;
        if (spec->depth > 0) need(spec->blocks, "spec->blocks");
        for (int i = 0; i < 2 * spec->depth; i++) blocks.push_back(conv(spec->blocks[i]));
        kzb::RawNet r{};
        r.cin = spec->input_channels, r.board_h = spec->board_h, r.board_w = spec->board_w, r.channels = spec->channels, r.depth = spec->depth;
        r.first = conv(spec->first);
        r.blocks = blocks.data();
        r.final_scale = spec->final_scale, r.final_shift = spec->final_shift;
        r.scalar_conv = conv(spec->scalar_conv), r.fc1 = fc(spec->fc1), r.fc2 = fc(spec->fc2);
        r.policy_conv1 = conv(spec->policy_conv1), r.policy_conv2 = conv(spec->policy_conv2);
        r.has_extra = spec->has_extra != 0;
        r.extra_conv = conv(spec->extra_conv), r.extra_fc = fc(spec->extra_fc);
        r.policy_len = spec->policy_len, r.policy_src = spec->policy_src;
        *out = new kzb_net(device, kzb::net_spec_from_raw(r), max_batch, precision);
    });
}

KZB_API int kzb_net_bind_mapper(kzb_net* net, int scalar_count, int bool_channels, int board_h, int board_w, int policy_len) {
    return guarded([&] {
        need(net, "net");
        net->impl.bind_mapper(scalar_count, bool_channels, board_h, board_w, policy_len);
    });
}

KZB_API int kzb_net_get_info(const kzb_net* net, kzb_net_info* out) {
    return guarded([&] {
        need(net, "net");
        need(out, "out");
        const kzb::NetSpec& s = net->impl.spec();
        out->input_channels = s.cin;
        out->board_h = s.board_h;
        out->board_w = s.board_w;
        out->policy_len = s.policy_len;
        out->channels = s.channels;
        out->depth = s.depth;
        out->max_batch = net->impl.max_batch();
        out->precision = net->impl.precision();
        out->device = net->impl.device();
        out->conv_mode = net->impl.conv_mode();
        out->flops_per_position = s.flops_per_position();
    });
}

KZB_API int kzb_onnx_inspect(const void* onnx_bytes, size_t onnx_len, kzb_net_info* out) {
    return guarded([&] {
        need(onnx_bytes, "onnx_bytes");
        need(out, "out");
        kzb::NetSpec s = kzb::build_net_spec(kzb::parse_onnx(onnx_bytes, onnx_len));
        out->input_channels = s.cin;
        out->board_h = s.board_h;
        out->board_w = s.board_w;
        out->policy_len = s.policy_len;
        out->channels = s.channels;
        out->depth = s.depth;
        out->max_batch = out->precision = out->device = out->conv_mode = -1;
        out->flops_per_position = s.flops_per_position();
    });
}

KZB_API void kzb_net_destroy(kzb_net* net) {
    try {
        delete net;
    } catch (...) {
    }
}

KZB_API int kzb_eval_planes(kzb_net* net, const float* nchw_in, int batch, float* out_scalars, float* out_policy_logits) {
    return guarded([&] {
        need(net, "net");
        need(nchw_in, "nchw_in");
        need(out_scalars, "out_scalars");
        need(out_policy_logits, "out_policy_logits");
        net->impl.eval_planes(nchw_in, batch, out_scalars, out_policy_logits);
    });
}

KZB_API int kzb_eval_packed(kzb_net* net, const uint8_t* bits, const float* scalars, int batch, const uint32_t* mv_idx,
                            const uint32_t* mv_off, float* out_values, float* out_policy) {
    return guarded([&] {
        need(net, "net");
        need(bits, "bits");
        need(mv_off, "mv_off");
        need(out_values, "out_values");
        if (batch > 0 && mv_off[batch] > 0) {
            need(mv_idx, "mv_idx");
            need(out_policy, "out_policy");
        }
        net->impl.eval_packed(bits, scalars, batch, mv_idx, mv_off, out_values, out_policy);
    });
}

KZB_API int kzb_net_set_symmetries(kzb_net* net, int n_sym, const int32_t* square_src, const int32_t* policy_map) {
    return guarded([&] {
        need(net, "net");
        need(square_src, "square_src");
        need(policy_map, "policy_map");
        net->impl.set_symmetries(n_sym, square_src, policy_map);
    });
}

KZB_API int kzb_eval_packed_sym(kzb_net* net, const uint8_t* bits, const float* scalars, const uint8_t* sym, int batch,
                                const uint32_t* mv_idx, const uint32_t* mv_off, float* out_values, float* out_policy) {
    return guarded([&] {
        need(net, "net");
        need(bits, "bits");
        need(sym, "sym");
        need(mv_off, "mv_off");
        need(out_values, "out_values");
        if (batch > 0 && mv_off[batch] > 0) {
            need(mv_idx, "mv_idx");
            need(out_policy, "out_policy");
        }
        net->impl.eval_packed(bits, scalars, batch, mv_idx, mv_off, out_values, out_policy, sym);
    });
}

KZB_API int kzb_encode_planes(kzb_net* net, const uint8_t* bits, const float* scalars, int batch, float* out_nchw) {
    return guarded([&] {
        need(net, "net");
        need(bits, "bits");
        need(out_nchw, "out_nchw");
        net->impl.encode_planes(bits, scalars, batch, out_nchw);
    });
}

KZB_API int kzb_stage_packed(kzb_net* net, const uint8_t* bits, const float* scalars, int batch, const uint32_t* mv_idx,
                             const uint32_t* mv_off) {
    return guarded([&] {
        need(net, "net");
        need(bits, "bits");
        need(mv_off, "mv_off");
        net->impl.stage_packed(bits, scalars, batch, mv_idx, mv_off);
    });
}

KZB_API int kzb_time_staged(kzb_net* net, int iters, int flush_l2, float* ms_out) {
    return guarded([&] {
        need(net, "net");
        need(ms_out, "ms_out");
        if (iters < 1) throw std::runtime_error("iters must be >= 1");
        net->impl.time_staged(iters, flush_l2 != 0, ms_out);
    });
}

KZB_API int kzb_profile_staged(kzb_net* net, int flush_l2, char* names_out, size_t names_cap, float* ms_out, int ms_cap,
                               int* n_steps) {
    return guarded([&] {
        need(net, "net");
        need(names_out, "names_out");
        need(ms_out, "ms_out");
        need(n_steps, "n_steps");
        std::vector<std::string> names;
        std::vector<float> ms;
        net->impl.profile_staged(flush_l2 != 0, names, ms);
        std::string joined;
        for (size_t i = 0; i < names.size(); i++) joined += (i ? "\n" : "") + names[i];
        if (joined.size() + 1 > names_cap || int(ms.size()) > ms_cap) throw std::runtime_error("profile buffers too small");
        std::memcpy(names_out, joined.c_str(), joined.size() + 1);
        std::memcpy(ms_out, ms.data(), ms.size() * sizeof(float));
        *n_steps = int(ms.size());
    });
}

KZB_API int kzb_launches_per_eval(const kzb_net* net) { return net ? net->impl.launches_per_eval() : 0; }
