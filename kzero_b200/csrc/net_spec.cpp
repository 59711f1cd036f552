#include "net_spec.hpp"

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
             " (supported: conv policy heads of post_act.py:54-112; the attention head is not built yet)");
    }

    void match_heads() {
        // network/common.rs:176-196: (scalars [B,5], policy [B]+policy_shape); the 3-output legacy form
        // (value, wdl, policy) is not produced by the reference's current exporter
        if (g.outputs.size() != 2)
            fail("Wrong number of outputs, expected (scalars, policy), got " + std::to_string(g.outputs.size()));
        match_scalar_head(g.outputs[0].name);
        spec.policy_src = provenance(g.outputs[1].name);
        spec.policy_len = int(spec.policy_src.size());
        if (spec.policy_conv2.w.empty()) fail("policy head has no conv map");
        for (size_t i = 1; i < g.outputs[1].dims.size(); i++) spec.policy_shape.push_back(g.outputs[1].dims[i]);
    }
};

}  // namespace

double NetSpec::flops_per_position() const {
    double a = area();
    double f = 2.0 * a * 9 * cin * channels + double(depth) * 2 * (2.0 * a * 9 * channels * channels);
    f += 2.0 * a * channels * (scalar_conv.cout + policy_conv1.cout) + 2.0 * a * policy_conv1.cout * policy_conv2.cout;
    f += 2.0 * fc1.in * fc1.out + 2.0 * fc2.in * fc2.out;
    if (has_extra) f += 2.0 * a * channels + 2.0 * extra_fc.in * extra_fc.out;
    return f;
}

NetSpec build_net_spec(const OnnxGraph& g) {
    Matcher m(g);
    m.match_tower();
    m.match_heads();
    return std::move(m.spec);
}

}  // namespace kzb
