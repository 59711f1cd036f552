#include "onnx_reader.hpp"

#include <cstring>
#include <stdexcept>

namespace kzb {
namespace {

struct Cursor {
    const uint8_t* p;
    const uint8_t* end;

    bool done() const { return p >= end; }

    uint64_t varint() {
        uint64_t result = 0;
        int shift = 0;
        while (true) {
            if (p >= end) throw std::runtime_error("onnx: truncated varint");
            uint8_t b = *p++;
            result |= uint64_t(b & 0x7F) << shift;
            if (!(b & 0x80)) return result;
            shift += 7;
            if (shift > 63) throw std::runtime_error("onnx: varint too long");
        }
    }

    Cursor sub(uint64_t len) {
        if (uint64_t(end - p) < len) throw std::runtime_error("onnx: truncated length-delimited field");
        Cursor c{p, p + len};
        p += len;
        return c;
    }
};

struct Field {
    int number;
    int wire;
    uint64_t value;  // varint, or raw bits for fixed32/fixed64
    Cursor bytes;    // for wire type 2
};

bool next_field(Cursor& c, Field& f) {
    if (c.done()) return false;
    uint64_t key = c.varint();
    f.number = int(key >> 3);
    f.wire = int(key & 7);
    f.bytes = Cursor{nullptr, nullptr};
    switch (f.wire) {
        case 0: f.value = c.varint(); break;
        case 1: {
            Cursor s = c.sub(8);
            std::memcpy(&f.value, s.p, 8);
            break;
        }
        case 2: f.bytes = c.sub(c.varint()); break;
        case 5: {
            Cursor s = c.sub(4);
            uint32_t v;
            std::memcpy(&v, s.p, 4);
            f.value = v;
            break;
        }
        default: throw std::runtime_error("onnx: unsupported wire type");
    }
    return true;
}

std::string str(const Cursor& c) { return std::string(reinterpret_cast<const char*>(c.p), c.end - c.p); }

void read_varints(const Field& f, std::vector<int64_t>& out) {
    if (f.wire == 0) {
        out.push_back(int64_t(f.value));
    } else {
        Cursor c = f.bytes;
        while (!c.done()) out.push_back(int64_t(c.varint()));
    }
}

float bits_to_float(uint32_t v) {
    float f;
    std::memcpy(&f, &v, 4);
    return f;
}

OnnxTensor parse_tensor(Cursor c, std::string* name_out) {
    OnnxTensor t;
    t.dtype = 1;
    Cursor raw{nullptr, nullptr};
    bool has_raw = false;
    std::vector<float> floats;
    std::vector<int64_t> ints;
    Field f;
    while (next_field(c, f)) {
        switch (f.number) {
            case 1: read_varints(f, t.dims); break;
            case 2: t.dtype = int(f.value); break;
            case 4:
                if (f.wire == 5) {
                    floats.push_back(bits_to_float(uint32_t(f.value)));
                } else {
                    size_t n = (f.bytes.end - f.bytes.p) / 4;
                    size_t at = floats.size();
                    floats.resize(at + n);
                    std::memcpy(floats.data() + at, f.bytes.p, n * 4);
                }
                break;
            case 5:
            case 7: read_varints(f, ints); break;
            case 8:
                if (name_out) *name_out = str(f.bytes);
                break;
            case 9:
                raw = f.bytes;
                has_raw = true;
                break;
            default: break;
        }
    }
    int64_t n = t.numel();
    if (t.dtype == 1) {
        if (has_raw) {
            if (raw.end - raw.p != n * 4) throw std::runtime_error("onnx: f32 raw_data size mismatch");
            t.f32.resize(n);
            std::memcpy(t.f32.data(), raw.p, n * 4);
        } else {
            t.f32 = std::move(floats);
        }
        if (int64_t(t.f32.size()) != n) throw std::runtime_error("onnx: f32 tensor element count mismatch");
    } else if (t.dtype == 7 || t.dtype == 6) {
        if (has_raw) {
            size_t w = t.dtype == 7 ? 8 : 4;
            if (raw.end - raw.p != int64_t(n * w)) throw std::runtime_error("onnx: int raw_data size mismatch");
            t.i64.resize(n);
            for (int64_t i = 0; i < n; i++) {
                if (w == 8) {
                    int64_t v;
                    std::memcpy(&v, raw.p + i * 8, 8);
                    t.i64[i] = v;
                } else {
                    int32_t v;
                    std::memcpy(&v, raw.p + i * 4, 4);
                    t.i64[i] = v;
                }
            }
        } else {
            t.i64 = std::move(ints);
        }
        if (int64_t(t.i64.size()) != n) throw std::runtime_error("onnx: int tensor element count mismatch");
    } else {
        throw std::runtime_error("onnx: unsupported tensor data type " + std::to_string(t.dtype));
    }
    return t;
}

void parse_attr(Cursor c, OnnxNode& node) {
    std::string name;
    OnnxAttr a;
    Field f;
    while (next_field(c, f)) {
        switch (f.number) {
            case 1: name = str(f.bytes); break;
            case 2: a.f = bits_to_float(uint32_t(f.value)); break;
            case 3: a.i = int64_t(f.value); break;
            case 5:
                a.t = parse_tensor(f.bytes, nullptr);
                a.has_t = true;
                break;
            case 8: read_varints(f, a.ints); break;
            default: break;
        }
    }
    node.attrs[name] = std::move(a);
}

OnnxNode parse_node(Cursor c) {
    OnnxNode n;
    Field f;
    while (next_field(c, f)) {
        switch (f.number) {
            case 1: n.inputs.push_back(str(f.bytes)); break;
            case 2: n.outputs.push_back(str(f.bytes)); break;
            case 3: n.name = str(f.bytes); break;
            case 4: n.op = str(f.bytes); break;
            case 5: parse_attr(f.bytes, n); break;
            default: break;
        }
    }
    return n;
}

OnnxValueInfo parse_value_info(Cursor c) {
    OnnxValueInfo vi;
    Field f;
    while (next_field(c, f)) {
        if (f.number == 1) {
            vi.name = str(f.bytes);
        } else if (f.number == 2) {  // TypeProto
            Cursor ty = f.bytes;
            Field f2;
            while (next_field(ty, f2)) {
                if (f2.number != 1) continue;  // tensor_type
                Cursor tt = f2.bytes;
                Field f3;
                while (next_field(tt, f3)) {
                    if (f3.number != 2) continue;  // shape
                    Cursor sh = f3.bytes;
                    Field f4;
                    while (next_field(sh, f4)) {
                        if (f4.number != 1) continue;  // dim
                        int64_t d = -1;
                        Cursor dm = f4.bytes;
                        Field f5;
                        while (next_field(dm, f5))
                            if (f5.number == 1) d = int64_t(f5.value);
                        vi.dims.push_back(d);
                    }
                }
            }
        }
    }
    return vi;
}

}  // namespace

OnnxGraph parse_onnx(const void* data, size_t size) {
    Cursor model{static_cast<const uint8_t*>(data), static_cast<const uint8_t*>(data) + size};
    Cursor graph{nullptr, nullptr};
    bool have_graph = false;
    Field f;
    while (next_field(model, f)) {
        if (f.number == 7 && f.wire == 2) {
            graph = f.bytes;
            have_graph = true;
        }
    }
    if (!have_graph) throw std::runtime_error("onnx: ModelProto has no graph");

    OnnxGraph g;
    std::vector<OnnxValueInfo> raw_inputs;
    while (next_field(graph, f)) {
        switch (f.number) {
            case 1: g.nodes.push_back(parse_node(f.bytes)); break;
            case 5: {
                std::string name;
                OnnxTensor t = parse_tensor(f.bytes, &name);
                g.initializers[name] = std::move(t);
                break;
            }
            case 11: raw_inputs.push_back(parse_value_info(f.bytes)); break;
            case 12: g.outputs.push_back(parse_value_info(f.bytes)); break;
            default: break;
        }
    }
    for (auto& vi : raw_inputs)
        if (!g.initializers.count(vi.name)) g.inputs.push_back(vi);
    return g;
}

}  // namespace kzb
