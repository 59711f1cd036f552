#include "net_spec.hpp"

#include <algorithm>
#include <cmath>
#include <map>
#include <stdexcept>

namespace kzb {
namespace {

[[noreturn]] void fail(const std::string& msg) { throw std::runtime_error("net_spec: " + msg); }

struct Matcher {
    const OnnxGraph& g;
    std::map<std::string, int> producer;
    std::map<std::string, std::vector<int>> consumers;
    NetSpec spec;
    std::string tower_out;         // output of the final BN (input of every head conv)
    std::vector<double> bn_a, bn_b;  // final BN as y = a*x + b

    explicit Matcher(const OnnxGraph& graph) : g(graph) {
        for (int i = 0; i < int(g.nodes.size()); i++) {
            for (auto& o : g.nodes[i].outputs) producer[o] = i;
            for (auto& in : g.nodes[i].inputs) consumers[in].push_back(i);
        }
    }

    const OnnxNode* prod(const std::string& name) const {
        auto it = producer.find(name);
        return it == producer.end() ? nullptr : &g.nodes[it->second];
    }

    // initializer, Constant node, or Identity chain of those (the exporter de-duplicates identical
    // tensors through Identity nodes)
    const OnnxTensor* constant(const std::string& name) const {
        auto it = g.initializers.find(name);
        if (it != g.initializers.end()) return &it->second;
        const OnnxNode* n = prod(name);
        if (!n) return nullptr;
        if (n->op == "Identity") return constant(n->inputs.at(0));
        if (n->op == "Constant") {
            auto a = n->attrs.find("value");
            if (a != n->attrs.end() && a->second.has_t) return &a->second.t;
        }
        return nullptr;
    }

    const OnnxTensor& f32_const(const std::string& name, const char* what) const {
        const OnnxTensor* t = constant(name);
        if (!t || t->dtype != 1) fail(std::string(what) + " '" + name + "' is not a constant f32 tensor");
        return *t;
    }

    const OnnxNode& sole_consumer(const std::string& name, const char* expect_op) const {
        auto it = consumers.find(name);
        if (it == consumers.end() || it->second.size() != 1)
            fail("expected exactly one consumer (" + std::string(expect_op) + ") of '" + name + "'");
        const OnnxNode& n = g.nodes[it->second[0]];
        if (n.op != expect_op) fail("expected " + std::string(expect_op) + " after '" + name + "', found " + n.op);
        return n;
    }

    ConvParams read_conv(const OnnxNode& n, int expect_k) const {
        if (n.op != "Conv") fail("expected Conv, found " + n.op);
        auto ks = n.attr_ints("kernel_shape");
        auto pads = n.attr_ints("pads");
        auto strides = n.attr_ints("strides");
        auto dil = n.attr_ints("dilations");
        if (n.attr_i("group", 1) != 1) fail("Conv group != 1");
        for (auto s : strides)
            if (s != 1) fail("Conv stride != 1");
        for (auto d : dil)
            if (d != 1) fail("Conv dilation != 1");
        const OnnxTensor& w = f32_const(n.inputs.at(1), "Conv weight");
        if (w.dims.size() != 4 || w.dims[2] != w.dims[3]) fail("Conv weight must be [co,ci,k,k]");
        int k = int(w.dims[2]);
        if (!ks.empty() && (ks[0] != k || ks[1] != k)) fail("Conv kernel_shape disagrees with weight");
        if (k != expect_k) fail("expected " + std::to_string(expect_k) + "x" + std::to_string(expect_k) + " Conv, found k=" + std::to_string(k));
        for (auto p : pads)
            if (p != k / 2) fail("Conv padding must be k/2 ('same')");
        if (pads.empty() && k != 1) fail("Conv without pads");
        ConvParams c;
        c.cout = int(w.dims[0]);
        c.cin = int(w.dims[1]);
        c.ksize = k;
        c.w = w.f32;
        if (n.inputs.size() > 2 && !n.inputs[2].empty()) {
            const OnnxTensor& b = f32_const(n.inputs[2], "Conv bias");
            if (b.numel() != c.cout) fail("Conv bias size mismatch");
            c.b = b.f32;
        } else {
            c.b.assign(c.cout, 0.0f);
        }
        return c;
    }

    // y = a*x + b per channel
    void read_bn(const OnnxNode& n, std::vector<double>& a, std::vector<double>& b) const {
        double eps = n.attr_f("epsilon", 1e-5f);
        const OnnxTensor& gamma = f32_const(n.inputs.at(1), "BN weight");
        const OnnxTensor& beta = f32_const(n.inputs.at(2), "BN bias");
        const OnnxTensor& mean = f32_const(n.inputs.at(3), "BN running_mean");
        const OnnxTensor& var = f32_const(n.inputs.at(4), "BN running_var");
        size_t c = gamma.f32.size();
        if (beta.f32.size() != c || mean.f32.size() != c || var.f32.size() != c) fail("BN parameter size mismatch");
        a.resize(c);
        b.resize(c);
        for (size_t i = 0; i < c; i++) {
            a[i] = double(gamma.f32[i]) / std::sqrt(double(var.f32[i]) + eps);
            b[i] = double(beta.f32[i]) - a[i] * double(mean.f32[i]);
        }
    }

    // If `cur` feeds exactly one BatchNormalization, fold it into `c` (conv -> bn) and advance `cur`.
    void fold_following_bn(ConvParams& c, std::string& cur) const {
        auto it = consumers.find(cur);
        if (it == consumers.end() || it->second.size() != 1) return;
        const OnnxNode& n = g.nodes[it->second[0]];
        if (n.op != "BatchNormalization") return;
        std::vector<double> a, b;
        read_bn(n, a, b);
        if (int(a.size()) != c.cout) fail("BN channel count does not match preceding Conv");
        size_t per = size_t(c.cin) * c.ksize * c.ksize;
        for (int o = 0; o < c.cout; o++) {
            for (size_t i = 0; i < per; i++) c.w[o * per + i] = float(a[o] * double(c.w[o * per + i]));
            c.b[o] = float(a[o] * double(c.b[o]) + b[o]);
        }
        cur = n.outputs.at(0);
    }

    // fold the final BN (applied to the conv's input) into a 1x1 conv
    void fold_preceding_bn(ConvParams& c) const {
        if (bn_a.empty()) return;
        if (c.ksize != 1 || c.cin != int(bn_a.size())) fail("head conv after the tower must be 1x1 over C channels");
        for (int o = 0; o < c.cout; o++) {
            double extra = 0;
            for (int i = 0; i < c.cin; i++) {
                double w = c.w[size_t(o) * c.cin + i];
                extra += w * bn_b[i];
                c.w[size_t(o) * c.cin + i] = float(w * bn_a[i]);
            }
            c.b[o] = float(double(c.b[o]) + extra);
        }
    }

    FcParams read_gemm(const OnnxNode& n) const {
        if (n.op != "Gemm") fail("expected Gemm, found " + n.op);
        if (n.attr_i("transA", 0) != 0 || n.attr_i("transB", 0) != 1) fail("Gemm must have transA=0, transB=1");
        if (n.attr_f("alpha", 1.0f) != 1.0f || n.attr_f("beta", 1.0f) != 1.0f) fail("Gemm alpha/beta must be 1");
        const OnnxTensor& w = f32_const(n.inputs.at(1), "Gemm weight");
        if (w.dims.size() != 2) fail("Gemm weight must be 2-D");
        FcParams f;
        f.out = int(w.dims[0]);
        f.in = int(w.dims[1]);
        f.w = w.f32;
        if (n.inputs.size() > 2 && !n.inputs[2].empty()) {
            const OnnxTensor& b = f32_const(n.inputs[2], "Gemm bias");
            if (b.numel() != f.out) fail("Gemm bias size mismatch");
            f.b = b.f32;
        } else {
            f.b.assign(f.out, 0.0f);
        }
        return f;
    }

    const OnnxNode& expect_prod(const std::string& name, const char* op) const {
        const OnnxNode* n = prod(name);
        if (!n || n->op != op)
            fail("expected '" + name + "' to be produced by " + op + (n ? ", found " + n->op : ", found a graph input/constant"));
        return *n;
    }

    void match_tower() {
        if (g.inputs.size() != 1) fail("Wrong number of inputs");  // network/common.rs:167-168
        const OnnxValueInfo& in = g.inputs[0];
        if (in.dims.size() != 4 || in.dims[0] != -1 || in.dims[1] <= 0 || in.dims[2] <= 0 || in.dims[3] <= 0)
            fail("input must be [BATCH, C, H, W]");
        spec.cin = int(in.dims[1]);
        spec.board_h = int(in.dims[2]);
        spec.board_w = int(in.dims[3]);

        std::string cur = in.name;
        spec.first = read_conv(sole_consumer(cur, "Conv"), 3);
        if (spec.first.cin != spec.cin) fail("first conv input channels do not match the graph input");
        spec.channels = spec.first.cout;
        cur = sole_consumer(cur, "Conv").outputs.at(0);
        fold_following_bn(spec.first, cur);

        while (true) {
            auto it = consumers.find(cur);
            if (it == consumers.end() || it->second.size() != 2) break;
            const OnnxNode* conv = nullptr;
            const OnnxNode* add = nullptr;
            for (int idx : it->second) {
                const OnnxNode& n = g.nodes[idx];
                if (n.op == "Conv") conv = &n;
                if (n.op == "Add") add = &n;
            }
            if (!conv || !add) break;
            std::string y = cur;
            for (int j = 0; j < 2; j++) {
                const OnnxNode& cn = j == 0 ? *conv : sole_consumer(y, "Conv");
                ConvParams c = read_conv(cn, 3);
                if (c.cin != spec.channels || c.cout != spec.channels) fail("tower conv must be C->C");
                y = cn.outputs.at(0);
                fold_following_bn(c, y);
                y = sole_consumer(y, "Relu").outputs.at(0);
                spec.blocks.push_back(std::move(c));
            }
            const OnnxNode& a = sole_consumer(y, "Add");
            if (&a != add) fail("residual Add does not close the block");
            bool ok = (a.inputs.at(0) == cur && a.inputs.at(1) == y) || (a.inputs.at(1) == cur && a.inputs.at(0) == y);
            if (!ok) fail("residual Add inputs are not (block input, block output)");
            cur = a.outputs.at(0);
            spec.depth++;
        }

        // final BN (post_act.py:207); folded into the head 1x1 convs below
        auto it = consumers.find(cur);
        if (it != consumers.end() && it->second.size() == 1 && g.nodes[it->second[0]].op == "BatchNormalization") {
            const OnnxNode& bn = g.nodes[it->second[0]];
            read_bn(bn, bn_a, bn_b);
            if (int(bn_a.size()) != spec.channels) fail("final BN channel count mismatch");
            cur = bn.outputs.at(0);
        }
        tower_out = cur;
    }

    ConvParams head_conv_from_tower(const OnnxNode& n) const {
        if (n.inputs.at(0) != tower_out) fail("head conv does not read the tower output");
        ConvParams c = read_conv(n, 1);
        if (c.cin != spec.channels) fail("head conv input channels mismatch");
        fold_preceding_bn(c);
        return c;
    }

    void match_scalar_head(const std::string& out) {
        const OnnxNode& g2 = expect_prod(out, "Gemm");
        spec.fc2 = read_gemm(g2);
        const OnnxNode& r2 = expect_prod(g2.inputs.at(0), "Relu");
        const OnnxNode& g1 = expect_prod(r2.inputs.at(0), "Gemm");
        spec.fc1 = read_gemm(g1);
        const OnnxNode& fl = expect_prod(g1.inputs.at(0), "Flatten");
        const OnnxNode& r1 = expect_prod(fl.inputs.at(0), "Relu");
        const OnnxNode& cv = expect_prod(r1.inputs.at(0), "Conv");
        spec.scalar_conv = head_conv_from_tower(cv);
        if (spec.fc2.out != 5) fail("Wrong scalars shape");  // network/common.rs:181
        if (spec.fc1.in != spec.scalar_conv.cout * spec.area()) fail("scalar head fc1 input size mismatch");
        if (spec.fc2.in != spec.fc1.out) fail("scalar head fc2 input size mismatch");
    }

    std::vector<int32_t> provenance(const std::string& name, int depth_guard = 0) {
        if (depth_guard > 16) fail("policy head too deep");
        const OnnxNode* n = prod(name);
        if (!n) fail("policy output depends on non-node value '" + name + "'");
        const std::string& op = n->op;
        if (op == "Identity" || op == "Flatten" || op == "Reshape") return provenance(n->inputs.at(0), depth_guard + 1);
        if (op == "Gather") {
            if (n->attr_i("axis", 0) != 1) fail("policy Gather must be over axis 1");
            std::vector<int32_t> data = provenance(n->inputs.at(0), depth_guard + 1);
            const OnnxTensor* idx = constant(n->inputs.at(1));
            if (!idx || idx->dtype == 1) fail("policy Gather indices must be a constant integer tensor");
            std::vector<int32_t> out(idx->i64.size());
            for (size_t i = 0; i < out.size(); i++) {
                int64_t j = idx->i64[i];
                if (j < 0) j += int64_t(data.size());
                if (j < 0 || j >= int64_t(data.size())) fail("policy Gather index out of range");
                out[i] = data[size_t(j)];
            }
            return out;
        }
        if (op == "Concat") {
            if (n->attr_i("axis", 0) != 1) fail("policy Concat must be over axis 1");
            std::vector<int32_t> out;
            for (auto& in : n->inputs) {
                std::vector<int32_t> part = provenance(in, depth_guard + 1);
                out.insert(out.end(), part.begin(), part.end());
            }
            return out;
        }
        if (op == "ConstantOfShape") {
            auto a = n->attrs.find("value");
            if (a != n->attrs.end() && a->second.has_t) {
                const OnnxTensor& v = a->second.t;
                bool zero = v.dtype == 1 ? (v.f32.size() == 1 && v.f32[0] == 0.0f) : (v.i64.size() == 1 && v.i64[0] == 0);
                if (!zero) fail("policy ConstantOfShape value must be 0");
            }
            // shape = Concat(batch (dynamic), static dims...): count = product of the static parts
            const OnnxNode& sh = expect_prod(n->inputs.at(0), "Concat");
            int64_t count = 1;
            int dynamic = 0;
            for (auto& in : sh.inputs) {
                const OnnxTensor* c = constant(in);
                if (c && c->dtype != 1) {
                    for (auto v : c->i64) count *= v;
                } else {
                    dynamic++;
                }
            }
            if (dynamic != 1 || count <= 0 || count > (1 << 20)) fail("cannot size the policy ConstantOfShape");
            return std::vector<int32_t>(size_t(count), kPolicySrcZero);
        }
        if (op == "Gemm") {  // go: conv1x1(C->1) -> Flatten -> Linear(A -> extra)   post_act.py:63-67
            if (spec.has_extra) fail("more than one extra policy branch");
            spec.extra_fc = read_gemm(*n);
            const OnnxNode& fl = expect_prod(n->inputs.at(0), "Flatten");
            const OnnxNode& cv = expect_prod(fl.inputs.at(0), "Conv");
            spec.extra_conv = head_conv_from_tower(cv);
            if (spec.extra_conv.cout != 1) fail("extra policy conv must have 1 output channel");
            if (spec.extra_fc.in != spec.area()) fail("extra policy fc input size mismatch");
            spec.has_extra = true;
            std::vector<int32_t> out(spec.extra_fc.out);
            for (int e = 0; e < spec.extra_fc.out; e++) out[e] = kPolicySrcExtra - e;
            return out;
        }
        if (op == "Conv") {  // the policy map: conv1x1(t)->relu->conv1x1(Pc)
            if (!spec.policy_conv2.w.empty()) fail("more than one policy conv map");
            spec.policy_conv2 = read_conv(*n, 1);
            const OnnxNode& r = expect_prod(n->inputs.at(0), "Relu");
            const OnnxNode& c1 = expect_prod(r.inputs.at(0), "Conv");
            spec.policy_conv1 = head_conv_from_tower(c1);
            if (spec.policy_conv2.cin != spec.policy_conv1.cout) fail("policy conv2 input channels mismatch");
            int a = spec.area();
            std::vector<int32_t> out(size_t(spec.policy_conv2.cout) * a);
            for (size_t i = 0; i < out.size(); i++) out[i] = int32_t(i);
            return out;
        }
        fail("unsupported op in policy head: " + op +
             " (supported: the conv policy heads of post_act.py:54-112 and the attention head :115-141)");
    }

    // ------------------------------------------------------------------------------------------------
    // Attention policy head (post_act.py:115-141).  The exporter turns its slicing / reshaping into a long
    // chain of Slice / Shape / Concat / Reshape / Transpose / Gather / Unsqueeze nodes.
    // Instead of matching that chain node by node, the head's index plumbing is EVALUATED symbolically at
    // batch size 1 on integer tensors of element ids: id = (src << 32) | (channel * A + square), src 0 = the
    // tower output, src k+1 = output of the k-th 1x1 head conv.  Pure index ops just shuffle ids, so at the
    // MatMul both operands say exactly which conv outputs are multiplied.
    struct Sym {
        std::vector<int64_t> shape;
        std::vector<int64_t> v;
        bool elem = false;
        int64_t numel() const {
            int64_t n = 1;
            for (auto d : shape) n *= d;
            return n;
        }
    };
    std::map<std::string, Sym> sym_memo;

    static std::vector<int64_t> strides_of(const std::vector<int64_t>& shape) {
        std::vector<int64_t> st(shape.size(), 1);
        for (int i = int(shape.size()) - 2; i >= 0; i--) st[i] = st[i + 1] * shape[i + 1];
        return st;
    }

    std::vector<int64_t> sym_ints(const std::string& name, const char* what) {
        const Sym& t = sym_eval(name);
        if (t.elem) fail(std::string(what) + " must be an integer tensor");
        return t.v;
    }

    const Sym& sym_eval(const std::string& name, int guard = 0) {
        auto memo = sym_memo.find(name);
        if (memo != sym_memo.end()) return memo->second;
        if (guard > 64) fail("policy head too deep");
        Sym out;
        const int area = spec.area();
        if (name == tower_out) {
            out.shape = {1, spec.channels, spec.board_h, spec.board_w};
            out.elem = true;
            out.v.resize(size_t(spec.channels) * area);
            for (size_t i = 0; i < out.v.size(); i++) out.v[i] = int64_t(i);
            return sym_memo[name] = std::move(out);
        }
        if (const OnnxTensor* c = constant(name)) {
            if (c->dtype == 1) fail("unexpected float constant '" + name + "' in the policy head's index plumbing");
            out.shape = c->dims;
            out.v = c->i64;
            return sym_memo[name] = std::move(out);
        }
        const OnnxNode* n = prod(name);
        if (!n) fail("policy output depends on non-node value '" + name + "'");
        const std::string& op = n->op;
        auto in = [&](size_t i) -> const Sym& { return sym_eval(n->inputs.at(i), guard + 1); };
        auto norm_axis = [&](int64_t ax, size_t rank) {
            if (ax < 0) ax += int64_t(rank);
            if (ax < 0 || ax >= int64_t(rank)) fail(op + " axis out of range");
            return size_t(ax);
        };
        if (op == "Identity") {
            out = in(0);
        } else if (op == "Shape") {
            const Sym& x = in(0);
            out.shape = {int64_t(x.shape.size())};
            out.v = x.shape;
        } else if (op == "Conv") {
            const Sym& x = in(0);
            if (!x.elem || x.shape.size() != 4 || x.shape[0] != 1 || x.shape[1] != spec.channels)
                fail("attention head conv must read (a spatial slice of) the tower output");
            const int64_t hw = x.shape[2] * x.shape[3];
            std::vector<int64_t> sq(size_t(hw), 0);
            for (int64_t c = 0; c < spec.channels; c++)
                for (int64_t i = 0; i < hw; i++) {
                    int64_t id = x.v[size_t(c * hw + i)];
                    if ((id >> 32) != 0 || (id & 0xffffffff) / area != c) fail("attention head conv input permutes tower channels");
                    int64_t s = (id & 0xffffffff) % area;
                    if (c == 0) sq[size_t(i)] = s;
                    else if (sq[size_t(i)] != s) fail("attention head conv input mixes squares");
                }
            ConvParams cp = read_conv(*n, 1);
            if (cp.cin != spec.channels) fail("head conv input channels mismatch");
            fold_preceding_bn(cp);
            const int64_t k = int64_t(spec.att_convs.size());
            out.shape = {1, cp.cout, x.shape[2], x.shape[3]};
            out.elem = true;
            out.v.resize(size_t(cp.cout * hw));
            for (int64_t co = 0; co < cp.cout; co++)
                for (int64_t i = 0; i < hw; i++) out.v[size_t(co * hw + i)] = ((k + 1) << 32) | (co * area + sq[size_t(i)]);
            spec.att_convs.push_back(std::move(cp));
        } else if (op == "Slice") {
            const Sym& x = in(0);
            std::vector<int64_t> starts, ends, axes, steps;
            if (n->inputs.size() >= 3) {  // opset >= 10: tensors
                starts = sym_ints(n->inputs[1], "Slice starts");
                ends = sym_ints(n->inputs[2], "Slice ends");
                if (n->inputs.size() > 3 && !n->inputs[3].empty()) axes = sym_ints(n->inputs[3], "Slice axes");
                if (n->inputs.size() > 4 && !n->inputs[4].empty()) steps = sym_ints(n->inputs[4], "Slice steps");
            } else {
                starts = n->attr_ints("starts");
                ends = n->attr_ints("ends");
                axes = n->attr_ints("axes");
            }
            if (axes.empty())
                for (size_t i = 0; i < starts.size(); i++) axes.push_back(int64_t(i));
            if (steps.empty()) steps.assign(starts.size(), 1);
            if (starts.size() != ends.size() || starts.size() != axes.size() || starts.size() != steps.size()) fail("malformed Slice");
            std::vector<int64_t> lo(x.shape.size(), 0), cnt = x.shape;
            for (size_t i = 0; i < axes.size(); i++) {
                size_t ax = norm_axis(axes[i], x.shape.size());
                if (steps[i] != 1) fail("Slice step != 1 is not supported");
                int64_t d = x.shape[ax], s0 = starts[i], e0 = ends[i];
                if (s0 < 0) s0 += d;
                if (e0 < 0) e0 += d;
                s0 = std::min(std::max<int64_t>(s0, 0), d);
                e0 = std::min(std::max<int64_t>(e0, 0), d);
                lo[ax] = s0;
                cnt[ax] = std::max<int64_t>(e0 - s0, 0);
            }
            out.shape = cnt;
            out.elem = x.elem;
            out.v.resize(size_t(out.numel()));
            auto xs = strides_of(x.shape), os = strides_of(out.shape);
            for (int64_t f = 0; f < out.numel(); f++) {
                int64_t src = 0, r = f;
                for (size_t d = 0; d < out.shape.size(); d++) {
                    int64_t i = r / os[d];
                    r %= os[d];
                    src += (i + lo[d]) * xs[d];
                }
                out.v[size_t(f)] = x.v[size_t(src)];
            }
        } else if (op == "Concat") {
            const Sym& first = in(0);
            size_t ax = norm_axis(n->attr_i("axis", 0), first.shape.size());
            out.shape = first.shape;
            out.elem = first.elem;
            out.shape[ax] = 0;
            std::vector<const Sym*> parts;
            for (size_t i = 0; i < n->inputs.size(); i++) {
                const Sym& p = in(i);
                if (p.shape.size() != first.shape.size() || p.elem != first.elem) fail("Concat of mismatched tensors");
                out.shape[ax] += p.shape[ax];
                parts.push_back(&p);
            }
            int64_t outer = 1, inner = 1;
            for (size_t d = 0; d < ax; d++) outer *= first.shape[d];
            for (size_t d = ax + 1; d < first.shape.size(); d++) inner *= first.shape[d];
            for (int64_t o = 0; o < outer; o++)
                for (const Sym* p : parts) {
                    int64_t len = p->shape[ax] * inner;
                    out.v.insert(out.v.end(), p->v.begin() + o * len, p->v.begin() + (o + 1) * len);
                }
        } else if (op == "Reshape" || op == "Flatten" || op == "Unsqueeze" || op == "Squeeze") {
            const Sym& x = in(0);
            out = x;
            if (op == "Reshape") {
                std::vector<int64_t> target = sym_ints(n->inputs.at(1), "Reshape shape");
                int64_t known = 1, infer = -1;
                for (size_t i = 0; i < target.size(); i++) {
                    if (target[i] == 0) {
                        if (i >= x.shape.size()) fail("Reshape 0 beyond input rank");
                        target[i] = x.shape[i];
                    }
                    if (target[i] == -1) {
                        if (infer >= 0) fail("Reshape with two -1");
                        infer = int64_t(i);
                    } else {
                        known *= target[i];
                    }
                }
                if (infer >= 0) {
                    if (known == 0 || x.numel() % known) fail("Reshape cannot infer -1");
                    target[size_t(infer)] = x.numel() / known;
                }
                out.shape = target;
            } else if (op == "Flatten") {
                size_t ax = size_t(n->attr_i("axis", 1));
                if (ax > x.shape.size()) fail("Flatten axis out of range");
                int64_t a0 = 1, a1 = 1;
                for (size_t d = 0; d < x.shape.size(); d++) (d < ax ? a0 : a1) *= x.shape[d];
                out.shape = {a0, a1};
            } else if (op == "Unsqueeze") {
                std::vector<int64_t> axes = n->attr_ints("axes");
                if (axes.empty() && n->inputs.size() > 1) axes = sym_ints(n->inputs[1], "Unsqueeze axes");
                std::vector<int64_t> shp = x.shape;
                std::sort(axes.begin(), axes.end());
                for (auto a : axes) {
                    if (a < 0) a += int64_t(shp.size()) + 1;
                    if (a < 0 || a > int64_t(shp.size())) fail("Unsqueeze axis out of range");
                    shp.insert(shp.begin() + a, 1);
                }
                out.shape = shp;
            } else {
                std::vector<int64_t> axes = n->attr_ints("axes");
                std::vector<int64_t> shp;
                for (size_t d = 0; d < x.shape.size(); d++) {
                    bool drop = axes.empty() ? x.shape[d] == 1 : false;
                    for (auto a : axes)
                        if (norm_axis(a, x.shape.size()) == d) drop = true;
                    if (drop && x.shape[d] != 1) fail("Squeeze of a non-unit dimension");
                    if (!drop) shp.push_back(x.shape[d]);
                }
                out.shape = shp;
            }
            if (out.numel() != x.numel()) fail(op + " changes the element count");
        } else if (op == "Transpose") {
            const Sym& x = in(0);
            std::vector<int64_t> perm = n->attr_ints("perm");
            if (perm.empty())
                for (size_t d = x.shape.size(); d-- > 0;) perm.push_back(int64_t(d));
            if (perm.size() != x.shape.size()) fail("Transpose perm rank mismatch");
            out.elem = x.elem;
            for (auto p : perm) out.shape.push_back(x.shape.at(size_t(p)));
            out.v.resize(x.v.size());
            auto xs = strides_of(x.shape), os = strides_of(out.shape);
            for (int64_t f = 0; f < out.numel(); f++) {
                int64_t src = 0, r = f;
                for (size_t d = 0; d < out.shape.size(); d++) {
                    src += (r / os[d]) * xs[size_t(perm[d])];
                    r %= os[d];
                }
                out.v[size_t(f)] = x.v[size_t(src)];
            }
        } else if (op == "Gather") {
            const Sym& x = in(0);
            const Sym& idx = in(1);
            if (idx.elem) fail("Gather indices must be integers");
            size_t ax = norm_axis(n->attr_i("axis", 0), x.shape.size());
            out.elem = x.elem;
            int64_t outer = 1, inner = 1;
            for (size_t d = 0; d < ax; d++) {
                outer *= x.shape[d];
                out.shape.push_back(x.shape[d]);
            }
            for (auto d : idx.shape) out.shape.push_back(d);
            for (size_t d = ax + 1; d < x.shape.size(); d++) {
                inner *= x.shape[d];
                out.shape.push_back(x.shape[d]);
            }
            const int64_t dim = x.shape[ax];
            for (int64_t o = 0; o < outer; o++)
                for (int64_t j : idx.v) {
                    if (j < 0) j += dim;
                    if (j < 0 || j >= dim) fail("Gather index out of range");
                    auto b = x.v.begin() + (o * dim + j) * inner;
                    out.v.insert(out.v.end(), b, b + inner);
                }
        } else {
            fail("unsupported op in the attention policy head: " + op);
        }
        return sym_memo[name] = std::move(out);
    }

    // one MatMul operand row/column -> (channel0, channel stride per q, square) inside the concatenated conv outputs
    NetSpec::AttOperand att_operand(const std::vector<int64_t>& ids) const {
        const int area = spec.area();
        NetSpec::AttOperand o{0, 0, 0};
        int64_t src0 = -1, chan0 = 0, sq0 = 0;
        for (size_t q = 0; q < ids.size(); q++) {
            int64_t src = ids[q] >> 32, rest = ids[q] & 0xffffffff;
            int64_t chan = rest / area, sq = rest % area;
            if (src < 1) fail("attention MatMul operand is not a head conv output");
            if (q == 0) {
                src0 = src;
                chan0 = chan;
                sq0 = sq;
            } else {
                if (src != src0 || sq != sq0) fail("attention MatMul operand mixes convs or squares along the query axis");
                if (q == 1) o.chan_stride = int32_t(chan - chan0);
                if (chan != chan0 + int64_t(q) * o.chan_stride) fail("attention MatMul operand is not affine in the query index");
            }
        }
        o.chan = int32_t(spec.att_chan_base[size_t(src0 - 1)] + chan0);
        o.sq = int32_t(sq0);
        return o;
    }

    // policy = [Gather(const, axis 1)] o Flatten o [Div|Mul scalar] o MatMul(X, Y); returns false if there is no MatMul
    bool match_attention_head(const std::string& policy_out) {
        std::string cur = policy_out;
        const OnnxTensor* gather_idx = nullptr;
        double div = 1.0;
        const OnnxNode* mm = nullptr;
        for (int guard = 0; guard < 16 && !mm; guard++) {
            const OnnxNode* n = prod(cur);
            if (!n) return false;
            if (n->op == "MatMul") {
                mm = n;
            } else if (n->op == "Identity" || (n->op == "Flatten" && n->attr_i("axis", 1) == 1)) {
                cur = n->inputs.at(0);
            } else if (n->op == "Gather" && !gather_idx && n->attr_i("axis", 0) == 1) {
                gather_idx = constant(n->inputs.at(1));
                if (!gather_idx || gather_idx->dtype == 1) return false;
                cur = n->inputs.at(0);
            } else if (n->op == "Div" || n->op == "Mul") {
                const OnnxTensor* c = constant(n->inputs.at(1));
                if (!c || c->dtype != 1 || c->f32.size() != 1) return false;
                div = n->op == "Div" ? div * double(c->f32[0]) : div / double(c->f32[0]);
                cur = n->inputs.at(0);
            } else {
                return false;
            }
        }
        if (!mm) return false;
        const Sym x = sym_eval(mm->inputs.at(0));
        const Sym y = sym_eval(mm->inputs.at(1));
        if (!x.elem || !y.elem || x.shape.size() != 3 || y.shape.size() != 3 || x.shape[0] != 1 || y.shape[0] != 1 ||
            x.shape[2] != y.shape[1])
            fail("attention MatMul operands must be [BATCH, M, Q] x [BATCH, Q, N]");
        const int64_t M = x.shape[1], Q = x.shape[2], N = y.shape[2];
        spec.att_chan_base.assign(1, 0);
        for (auto& c : spec.att_convs) spec.att_chan_base.push_back(spec.att_chan_base.back() + (c.cout + 15) / 16 * 16);
        std::vector<NetSpec::AttOperand> rows, cols;
        rows.resize(size_t(M));
        cols.resize(size_t(N));
        std::vector<int64_t> ids(size_t(Q), 0);
        for (int64_t i = 0; i < M; i++) {
            for (int64_t q = 0; q < Q; q++) ids[size_t(q)] = x.v[size_t(i * Q + q)];
            rows[size_t(i)] = att_operand(ids);
        }
        for (int64_t j = 0; j < N; j++) {
            for (int64_t q = 0; q < Q; q++) ids[size_t(q)] = y.v[size_t(q * N + j)];
            cols[size_t(j)] = att_operand(ids);
        }
        const int64_t count = gather_idx ? int64_t(gather_idx->i64.size()) : M * N;
        spec.att_entries.resize(size_t(count));
        for (int64_t p = 0; p < count; p++) {
            int64_t f = gather_idx ? gather_idx->i64[size_t(p)] : p;
            if (f < 0) f += M * N;
            if (f < 0 || f >= M * N) fail("policy Gather index out of range");
            spec.att_entries[size_t(p)] = NetSpec::AttEntry{rows[size_t(f / N)], cols[size_t(f % N)]};
        }
        spec.has_attention = true;
        spec.att_q = int(Q);
        spec.att_div = float(div);
        spec.policy_len = int(count);
        return true;
    }

    void match_heads() {
        // network/common.rs:176-196: (scalars [B,5], policy [B]+policy_shape); the 3-output legacy form
        // (value, wdl, policy) is not produced by the reference's current exporter
        if (g.outputs.size() == 3)
            fail("3-output graphs (value, wdl, policy: the legacy form of rust/kz-core/src/network/common.rs:43-50) are not supported: no model class "
                 "of the reference exports that form any more (save_onnx.py writes scalars, policy), so there is no architecture to recognise");
        if (g.outputs.size() != 2)
            fail("Wrong number of outputs, expected (scalars, policy), got " + std::to_string(g.outputs.size()));
        match_scalar_head(g.outputs[0].name);
        if (!match_attention_head(g.outputs[1].name)) {
            spec.policy_src = provenance(g.outputs[1].name);
            spec.policy_len = int(spec.policy_src.size());
            if (spec.policy_conv2.w.empty()) fail("policy head has no conv map");
        }
        for (size_t i = 1; i < g.outputs[1].dims.size(); i++) spec.policy_shape.push_back(g.outputs[1].dims[i]);
    }
};

}  // namespace

double NetSpec::flops_per_position() const {
    double a = area();
    double f = 2.0 * a * 9 * cin * channels + double(depth) * 2 * (2.0 * a * 9 * channels * channels);
    f += 2.0 * a * channels * (scalar_conv.cout + policy_conv1.cout) + 2.0 * a * policy_conv1.cout * policy_conv2.cout;
    if (has_attention) {
        // conv_bulk on every square, conv_under on one rank (post_act.py:131-132), then the [A, Q] x [Q, N] product
        for (auto& c : att_convs) f += 2.0 * channels * c.cout * (c.cout == 3 * att_q ? board_w : a);
        f += 2.0 * a * att_q * (a + 3.0 * board_w);
    }
    f += 2.0 * fc1.in * fc1.out + 2.0 * fc2.in * fc2.out;
    if (has_extra) f += 2.0 * a * channels + 2.0 * extra_fc.in * extra_fc.out;
    return f;
}

namespace {
[[noreturn]] void raw_fail(const std::string& what) { throw std::runtime_error("kzb_net_create: " + what); }
ConvParams raw_conv(const RawConv& c, int cin, int cout, int ksize, const char* name) {
    if (!c.w || !c.b) raw_fail(std::string(name) + ": weight / bias pointer is NULL");
    if (c.cin != cin || c.ksize != ksize || (cout > 0 && c.cout != cout) || c.cout < 1)
        raw_fail(std::string(name) + ": expected a " + std::to_string(ksize) + "x" + std::to_string(ksize) + " conv over " + std::to_string(cin) +
                 " channels" + (cout > 0 ? " with " + std::to_string(cout) + " outputs" : "") + ", got cin " + std::to_string(c.cin) + ", cout " +
                 std::to_string(c.cout) + ", ksize " + std::to_string(c.ksize));
    ConvParams p;
    p.cin = c.cin, p.cout = c.cout, p.ksize = c.ksize;
    p.w.assign(c.w, c.w + size_t(c.cout) * c.cin * c.ksize * c.ksize);
    p.b.assign(c.b, c.b + c.cout);
    return p;
}
FcParams raw_fc(const RawFc& f, int in, int out, const char* name) {
    if (!f.w || !f.b) raw_fail(std::string(name) + ": weight / bias pointer is NULL");
    if (f.in != in || (out > 0 && f.out != out) || f.out < 1)
        raw_fail(std::string(name) + ": expected " + std::to_string(in) + " inputs" + (out > 0 ? ", " + std::to_string(out) + " outputs" : "") + ", got " +
                 std::to_string(f.in) + " -> " + std::to_string(f.out));
    FcParams p;
    p.in = f.in, p.out = f.out;
    p.w.assign(f.w, f.w + size_t(f.in) * f.out);
    p.b.assign(f.b, f.b + f.out);
    return p;
}
// y = conv1x1(scale * x + shift): W' = W diag(scale), bias' = bias + W shift (exact: a 1x1 conv has no padding)
void fold_affine_into(ConvParams& c, const float* scale, const float* shift) {
    for (int o = 0; o < c.cout; o++) {
        double extra = 0;
        for (int i = 0; i < c.cin; i++) {
            const double w = c.w[size_t(o) * c.cin + i];
            extra += w * shift[i];
            c.w[size_t(o) * c.cin + i] = float(w * scale[i]);
        }
        c.b[size_t(o)] = float(double(c.b[size_t(o)]) + extra);
    }
}
}  // namespace

NetSpec net_spec_from_raw(const RawNet& r) {
    if (r.cin < 1 || r.board_h < 1 || r.board_w < 1 || r.channels < 1 || r.depth < 0) raw_fail("invalid shape");
    if (r.depth > 0 && !r.blocks) raw_fail("blocks is NULL");
    if ((r.final_scale == nullptr) != (r.final_shift == nullptr)) raw_fail("final_scale and final_shift must both be given or both be NULL");
    NetSpec s;
    s.cin = r.cin, s.board_h = r.board_h, s.board_w = r.board_w, s.channels = r.channels, s.depth = r.depth;
    const int C = r.channels, A = r.board_h * r.board_w;
    s.first = raw_conv(r.first, r.cin, C, 3, "first conv");
    for (int i = 0; i < 2 * r.depth; i++) s.blocks.push_back(raw_conv(r.blocks[i], C, C, 3, "block conv"));
    s.scalar_conv = raw_conv(r.scalar_conv, C, 0, 1, "scalar head conv");
    s.fc1 = raw_fc(r.fc1, s.scalar_conv.cout * A, 0, "scalar head fc1");
    s.fc2 = raw_fc(r.fc2, s.fc1.out, 5, "scalar head fc2");
    s.policy_conv1 = raw_conv(r.policy_conv1, C, 0, 1, "policy conv1");
    s.policy_conv2 = raw_conv(r.policy_conv2, s.policy_conv1.cout, 0, 1, "policy conv2");
    s.has_extra = r.has_extra;
    if (r.has_extra) {
        s.extra_conv = raw_conv(r.extra_conv, C, 1, 1, "extra policy conv");
        s.extra_fc = raw_fc(r.extra_fc, A, 0, "extra policy fc");
    }
    if (r.final_scale) {
        fold_affine_into(s.scalar_conv, r.final_scale, r.final_shift);
        fold_affine_into(s.policy_conv1, r.final_scale, r.final_shift);
        if (r.has_extra) fold_affine_into(s.extra_conv, r.final_scale, r.final_shift);
    }
    if (r.policy_len < 1 || !r.policy_src) raw_fail("policy_src is NULL or policy_len < 1");
    s.policy_len = r.policy_len;
    s.policy_src.assign(r.policy_src, r.policy_src + r.policy_len);
    for (int32_t v : s.policy_src) {
        const bool conv_ok = v >= 0 && v < s.policy_conv2.cout * A;
        const bool extra_ok = r.has_extra && v <= kPolicySrcExtra && v > kPolicySrcExtra - s.extra_fc.out;
        if (!conv_ok && v != kPolicySrcZero && !extra_ok) raw_fail("policy_src entry " + std::to_string(v) + " is out of range");
    }
    s.policy_shape = {int64_t(r.policy_len)};
    return s;
}

NetSpec build_net_spec(const OnnxGraph& g) {
    Matcher m(g);
    m.match_tower();
    m.match_heads();
    return std::move(m.spec);
}

}  // namespace kzb
