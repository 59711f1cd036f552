// Minimal ONNX (protobuf wire format) reader: just enough to load the nets the reference exports
// (python/lib/save_onnx.py:111-119, opset 10).  Replaces the loader half of kn-graph
// (`load_graph_from_onnx_path`, call site rust/kz-selfplay/src/server/server_alphazero.rs:126-128)
// for this path.  No protobuf library: the wire format is parsed directly.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace kzb {

struct OnnxTensor {
    std::vector<int64_t> dims;
    int dtype = 0;  // 1 = f32, 6 = i32, 7 = i64
    std::vector<float> f32;    // filled when dtype == 1
    std::vector<int64_t> i64;  // filled when dtype is an integer type

    int64_t numel() const {
        int64_t n = 1;
        for (auto d : dims) n *= d;
        return n;
    }
};

struct OnnxAttr {
    int64_t i = 0;
    float f = 0;
    std::vector<int64_t> ints;
    OnnxTensor t;
    bool has_t = false;
};

struct OnnxNode {
    std::string op, name;
    std::vector<std::string> inputs, outputs;
    std::map<std::string, OnnxAttr> attrs;

    int64_t attr_i(const std::string& k, int64_t dflt) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? dflt : it->second.i;
    }
    float attr_f(const std::string& k, float dflt) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? dflt : it->second.f;
    }
    std::vector<int64_t> attr_ints(const std::string& k) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? std::vector<int64_t>{} : it->second.ints;
    }
};

struct OnnxValueInfo {
    std::string name;
    std::vector<int64_t> dims;  // -1 for symbolic (batch) dims
};

struct OnnxGraph {
    std::vector<OnnxNode> nodes;
    std::map<std::string, OnnxTensor> initializers;
    std::vector<OnnxValueInfo> inputs;   // initializers filtered out
    std::vector<OnnxValueInfo> outputs;
};

// Throws std::runtime_error on malformed input.
OnnxGraph parse_onnx(const void* data, size_t size);

}  // namespace kzb
