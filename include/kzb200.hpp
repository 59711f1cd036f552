// kzb200.hpp -- C++17 mirror of kz-core's `Network` interface over the C ABI of kzb200.h (header only).
//
// The reference's host side is Rust: `trait Network<B>` (rust/kz-core/src/network/mod.rs:52-63) implemented by
// `CudaNetwork<B, M>` (rust/kz-core/src/network/cudnn.rs:18-88).  INTEGRATION.md holds the Rust shim a maintainer would add; this is the
// same shim in C++ for hosts that are C++ -- same names, same argument meaning, same error behaviour -- and what tests/cpp/
// network_mirror_test.cpp drives:
//
//   Rust                                                     here
//   CudaNetwork::new(mapper, &graph, max_batch, device)      B200Network<Board, Mapper>(mapper, onnx_bytes, len, max_batch, device)
//   Network::max_batch_size(&self)                           max_batch_size()
//   Network::evaluate_batch(&mut self, &[impl Borrow<B>])    evaluate_batch(boards)  -> std::vector<ZeroEvaluation>, one per board, in order
//   Network::evaluate(&mut self, &B)                         evaluate(board)
//   ZeroEvaluation { values: ZeroValuesPov, policy }         ZeroEvaluation { values, policy }: policy over the board's available moves
//                                                            only, in iteration order, sums to 1 (network/mod.rs:26-32)
//   panics (cudnn.rs:58, common.rs:110,171-196)              kzb200::Error (a std::runtime_error carrying kzb_last_error())
//
// `Mapper` plays BoardMapper<B> (rust/kz-core/src/mapping/mod.rs:9-36):
//   std::array<int, 3> input_bool_shape() const;            // [channels, height, width]
//   int input_scalar_count() const;
//   int policy_len() const;
//   void encode_input(uint8_t* bits, float* scalars, const Board&) const;   // bits: the BitBuffer's storage, LSB first, zeroed by the caller
//   void available_move_indices(const Board&, std::vector<uint32_t>& out) const;   // move_to_index over available_moves(), in order; empty
//                                                                                  // for a finished board
// A handle is single-threaded, like `&mut self`; distinct networks are independent.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <iterator>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "kzb200.h"

namespace kzb200 {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline void check(int rc) {  // the reference panics where the C ABI returns non-zero
    if (rc != 0) throw Error(kzb_last_error());
}

struct WDL {
    float win, draw, loss;
};
struct ZeroValuesPov {  // rust/kz-core/src/zero/values.rs:12-18
    float value;
    WDL wdl;
    float moves_left;
};
struct ZeroEvaluation {  // rust/kz-core/src/network/mod.rs:26-32
    ZeroValuesPov values;
    std::vector<float> policy;
};

inline int device_count() { return kzb_device_count(); }  // CudaDevice::all(), server.rs:49-51

template <typename Board, typename Mapper>
class B200Network {
public:
    // CudaNetwork::new (cudnn.rs:29-43) + check_graph_shapes (common.rs:165-198): throws when the graph does not fit the mapper
    B200Network(Mapper mapper, const void* onnx_bytes, size_t onnx_len, int max_batch_size, int device, int precision = KZB_PRECISION_BF16)
        : mapper_(std::move(mapper)), max_batch_size_(max_batch_size) {
        check(kzb_net_create_from_onnx(device, onnx_bytes, onnx_len, max_batch_size, precision, &handle_));
        bind();
    }
    // the same from an already optimised graph's constants (a host that keeps its own ONNX loading: kzb_net_create)
    B200Network(Mapper mapper, const kzb_net_spec& spec, int max_batch_size, int device, int precision = KZB_PRECISION_BF16)
        : mapper_(std::move(mapper)), max_batch_size_(max_batch_size) {
        check(kzb_net_create(device, &spec, max_batch_size, precision, &handle_));
        bind();
    }
    ~B200Network() { kzb_net_destroy(handle_); }  // executor.rs:326-331: the old network goes before the next one loads
    B200Network(const B200Network&) = delete;
    B200Network& operator=(const B200Network&) = delete;
    B200Network(B200Network&& o) noexcept
        : mapper_(std::move(o.mapper_)), handle_(o.handle_), max_batch_size_(o.max_batch_size_), bits_bytes_(o.bits_bytes_) {
        o.handle_ = nullptr;
    }

    int max_batch_size() const { return max_batch_size_; }
    const Mapper& mapper() const { return mapper_; }
    kzb_net_info info() const {
        kzb_net_info i;
        check(kzb_net_get_info(handle_, &i));
        return i;
    }

    // `boards`: any range of Board, const Board* or std::reference_wrapper<const Board> (impl Borrow<B>)
    template <typename Range>
    std::vector<ZeroEvaluation> evaluate_batch(const Range& boards) {
        size_t n = 0;
        for (auto it = std::begin(boards); it != std::end(boards); ++it) n++;
        if (n > size_t(max_batch_size_))  // cudnn.rs:58 assert!(batch_size <= max_batch_size)
            throw Error("batch size " + std::to_string(n) + " exceeds max_batch_size " + std::to_string(max_batch_size_));
        std::vector<ZeroEvaluation> out;
        if (n == 0) return out;
        const int scalar_count = mapper_.input_scalar_count();
        bits_.assign(n * bits_bytes_, 0);
        scalars_.assign(n * size_t(scalar_count), 0.0f);
        mv_idx_.clear();
        mv_off_.assign(1, 0u);
        size_t i = 0;
        for (auto it = std::begin(boards); it != std::end(boards); ++it, ++i) {
            const Board& board = deref(*it);
            // the non-`_full` half of encode_input_full (mapping/mod.rs:38,47-50): no f32 plane expansion on the CPU
            mapper_.encode_input(bits_.data() + i * bits_bytes_, scalars_.data() + i * size_t(scalar_count), board);
            // the list collect_policy_indices builds (rust/kz-selfplay/src/binary_output.rs:299-315)
            mapper_.available_move_indices(board, scratch_);
            mv_idx_.insert(mv_idx_.end(), scratch_.begin(), scratch_.end());
            mv_off_.push_back(uint32_t(mv_idx_.size()));
        }
        values_.resize(n * 5);
        probs_.resize(mv_idx_.size() + 1);
        check(kzb_eval_packed(handle_, bits_.data(), scalars_.data(), int(n), mv_idx_.data(), mv_off_.data(), values_.data(), probs_.data()));
        out.reserve(n);
        for (size_t b = 0; b < n; b++) {
            const float* v = values_.data() + b * 5;  // tanh / softmax already applied (common.rs:59-74)
            ZeroEvaluation e;
            e.values = ZeroValuesPov{v[0], WDL{v[1], v[2], v[3]}, v[4]};
            e.policy.assign(probs_.begin() + mv_off_[b], probs_.begin() + mv_off_[b + 1]);  // legal moves only, available_moves order
            out.push_back(std::move(e));
        }
        return out;
    }
    ZeroEvaluation evaluate(const Board& board) {  // Network::evaluate, network/mod.rs:57-62
        const Board* one[1] = {&board};
        return std::move(evaluate_batch(one)[0]);
    }

private:
    static const Board& deref(const Board& b) { return b; }
    static const Board& deref(const Board* b) { return *b; }
    template <typename W>
    static auto deref(const W& w) -> decltype(static_cast<const Board&>(w.get())) { return w.get(); }

    void bind() {
        const std::array<int, 3> shape = mapper_.input_bool_shape();
        bits_bytes_ = size_t(shape[0]) * size_t(shape[1]) * size_t(shape[2]);
        bits_bytes_ = (bits_bytes_ + 7) / 8;
        const int rc = kzb_net_bind_mapper(handle_, mapper_.input_scalar_count(), shape[0], shape[1], shape[2], mapper_.policy_len());
        if (rc != 0) {
            const std::string message = kzb_last_error();
            kzb_net_destroy(handle_);
            handle_ = nullptr;
            throw Error(message);
        }
    }

    Mapper mapper_;
    kzb_net* handle_ = nullptr;
    int max_batch_size_;
    size_t bits_bytes_ = 0;
    std::vector<uint8_t> bits_;
    std::vector<float> scalars_, values_, probs_;
    std::vector<uint32_t> mv_idx_, mv_off_, scratch_;
};

}  // namespace kzb200
