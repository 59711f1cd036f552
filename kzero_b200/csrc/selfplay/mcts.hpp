// AlphaZero tree search, the producer of the hot path's evaluation requests ("next" row N1 of SURVEY.md 8(f)).
//
// Restates the semantics (not the code) of the reference's search so that a self-play driver can feed the B200
// evaluator with realistic, ragged batches:
//   Node / Uct / UctWeights / UctContext      rust/kz-core/src/zero/node.rs:11-206
//   ZeroValuesAbs / pov / parent              rust/kz-core/src/zero/values.rs:6-70
//   zero_step_gather / zero_step_apply / tree_propagate_values   rust/kz-core/src/zero/step.rs:61-188
//   Tree::uct_context / policy / best_child   rust/kz-core/src/zero/tree.rs:49-146
//   choose_max_by_key (random tie break)      rust/kz-util/src/sequence.rs:11-41
// Header-only and generic over the game, like the Rust generics.  A Game provides:
//   bool done() const; int outcome() const (+1 player A won, 0 draw, -1 player B won; only if done);
//   int next_player() const (0 = A, 1 = B); void moves(std::vector<uint32_t>&) const; void play(uint32_t move);
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace kzb {
namespace selfplay {

// xorshift64* -- the oracle (oracle/mcts_oracle.py) implements the same generator so that tie breaks agree
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull) {
        if (s == 0) s = 0x2545F4914F6CDD1Dull;
    }
    uint64_t next_u64() {
        s ^= s >> 12;
        s ^= s << 25;
        s ^= s >> 27;
        return s * 0x2545F4914F6CDD1Dull;
    }
    uint32_t gen_range(uint32_t n) { return uint32_t((next_u64() >> 32) % n); }
    double uniform() { return double(next_u64() >> 11) * (1.0 / 9007199254740992.0); }  // [0, 1)
    bool gen_bool(double p) { return uniform() < p; }
};

// values.rs:6-10; "abs" = from player A's point of view
struct ValuesAbs {
    float value = 0, win_a = 0, draw = 0, win_b = 0, moves_left = 0;
    void add(const ValuesAbs& o) {
        value += o.value;
        win_a += o.win_a;
        draw += o.draw;
        win_b += o.win_b;
        moves_left += o.moves_left;
    }
    ValuesAbs div(float d) const { return {value / d, win_a / d, draw / d, win_b / d, moves_left / d}; }
    ValuesAbs parent() const { return {value, win_a, draw, win_b, moves_left + 1.0f}; }  // values.rs:62-68
    static ValuesAbs from_outcome(int outcome, float moves_left) {                         // values.rs:44-50
        return {float(outcome), outcome > 0 ? 1.0f : 0.0f, outcome == 0 ? 1.0f : 0.0f, outcome < 0 ? 1.0f : 0.0f, moves_left};
    }
};
// values.rs:12-16; from the point of view of `player`
struct ValuesPov {
    float value = 0, win = 0, draw = 0, loss = 0, moves_left = 0;
};
inline ValuesPov pov(const ValuesAbs& v, int player) {  // values.rs:21-28
    return player == 0 ? ValuesPov{v.value, v.win_a, v.draw, v.win_b, v.moves_left}
                       : ValuesPov{-v.value, v.win_b, v.draw, v.win_a, v.moves_left};
}
inline ValuesAbs un_pov(const ValuesPov& v, int player) {  // values.rs:33-40
    return player == 0 ? ValuesAbs{v.value, v.win, v.draw, v.loss, v.moves_left}
                       : ValuesAbs{-v.value, v.loss, v.draw, v.win, v.moves_left};
}

struct UctWeights {  // node.rs:47-53, defaults :66-75
    float exploration_weight = 2.0f, moves_left_weight = 0.03f, moves_left_clip = 20.0f, moves_left_sharpness = 0.5f;
};
struct FpuMode {  // step.rs:35-41
    bool relative = false;
    float value = 0.0f;
};
struct QMode {  // step.rs:43-51
    bool wdl = true;
    float draw_score = 0.0f;
    float select(const ValuesPov& v) const { return wdl ? v.win + draw_score * v.draw - v.loss : v.value; }  // step.rs:237-242
};

struct SearchSettings {
    UctWeights weights;
    QMode q_mode;
    FpuMode fpu_root{false, 0.1f}, fpu_child{true, 0.0f};  // python/main/loop_main_alpha.py:42-43
    float virtual_loss = 1.0f;
};

// node.rs:11-34.  One cache line per node: a search touches every child of every node on its path, and with hundreds of
// concurrent 800-visit trees the nodes do not stay in cache.  (`net_values` is only kept as a flag: the reference stores
// the network's own values to write them into game records, which this driver does not produce.)
struct alignas(64) Node {
    int32_t parent = -1;
    uint32_t last_move = 0;
    int32_t child_start = -1, child_count = 0;  // children == None  <=>  child_start < 0
    uint64_t complete_visits = 0, virtual_visits = 0;
    ValuesAbs sum_values;
    float net_policy = NAN;
    bool has_net_values = false;
    uint64_t total_visits() const { return complete_visits + virtual_visits; }
    ValuesAbs values() const { return sum_values.div(float(complete_visits)); }  // node.rs:126-128
};

struct UctContext {  // node.rs:55-64
    uint64_t total_visits;
    ValuesAbs values;
    float visited_policy_mass;
};

template <typename Game>
struct Tree {
    Game root_board;
    std::vector<Node> nodes;

    explicit Tree(const Game& root) : root_board(root) {  // tree.rs:31-39
        if (root.done()) throw std::runtime_error("Cannot build tree for done board");
        nodes.emplace_back();
    }
    uint64_t root_visits() const { return nodes[0].complete_visits; }

    UctContext uct_context(int node) const {  // tree.rs:49-66, node.rs:153-161
        const Node& n = nodes[node];
        float mass = 0.0f;
        for (int c = n.child_start; c < n.child_start + n.child_count; c++)
            if (nodes[c].total_visits() > 0) mass += nodes[c].net_policy;
        return {n.total_visits(), n.values(), mass};
    }

    // node.rs:163-206 + Uct::total :87-98
    float uct_total(const Node& child, const UctContext& parent, FpuMode fpu_mode, const SearchSettings& s, int player) const {
        if (parent.total_visits == 0) return NAN;
        float fpu;
        if (fpu_mode.relative) {
            float parent_value = s.q_mode.select(pov(parent.values, player));
            fpu = parent_value - fpu_mode.value * std::sqrt(parent.visited_policy_mass);
        } else {
            fpu = fpu_mode.value;
        }
        const float vl = s.virtual_loss;
        const float total_visits_virtual = float(child.complete_visits) + vl * float(child.virtual_visits);
        float q;
        if (total_visits_virtual == 0.0f) {
            q = fpu;
        } else {
            float total_value = s.q_mode.select(pov(child.sum_values, player));
            float total_value_virtual = total_value - vl * float(child.virtual_visits);
            q = total_value_virtual / total_visits_virtual;
        }
        const float u = child.net_policy * std::sqrt(float(parent.total_visits - 1)) / float(1 + child.total_visits());
        const float m = child.complete_visits == 0 ? 0.0f : child.values().moves_left - (parent.values.moves_left - 1.0f);
        const UctWeights& w = s.weights;
        float m_unit = 0.0f;
        if (w.moves_left_weight != 0.0f) {
            float m_clipped = std::fmin(std::fmax(m, -w.moves_left_clip), w.moves_left_clip);
            m_unit = std::fmin(std::fmax(w.moves_left_sharpness * m_clipped * -q, -1.0f), 1.0f);
        }
        return q + w.exploration_weight * u + w.moves_left_weight * m_unit;
    }

    // The per-parent part of uct_total, computed once per selection step instead of once per child (same float
    // operations in the same order, so the result is bit-identical to uct_total).
    struct UctParent {
        float fpu, sqrt_visits, moves_left_m1;
    };
    UctParent uct_parent(const UctContext& parent, FpuMode fpu_mode, const SearchSettings& s, int player) const {
        UctParent u;
        if (fpu_mode.relative) {
            float parent_value = s.q_mode.select(pov(parent.values, player));
            u.fpu = parent_value - fpu_mode.value * std::sqrt(parent.visited_policy_mass);
        } else {
            u.fpu = fpu_mode.value;
        }
        u.sqrt_visits = std::sqrt(float(parent.total_visits - 1));
        u.moves_left_m1 = parent.values.moves_left - 1.0f;
        return u;
    }
    float uct_total_fast(const Node& child, const UctParent& up, const SearchSettings& s, int player) const {
        const float vl = s.virtual_loss;
        const float total_visits_virtual = float(child.complete_visits) + vl * float(child.virtual_visits);
        float q;
        if (total_visits_virtual == 0.0f) {
            q = up.fpu;
        } else {
            float total_value = s.q_mode.select(pov(child.sum_values, player));
            float total_value_virtual = total_value - vl * float(child.virtual_visits);
            q = total_value_virtual / total_visits_virtual;
        }
        const float u = child.net_policy * up.sqrt_visits / float(1 + child.total_visits());
        const UctWeights& w = s.weights;
        float m_unit = 0.0f;
        if (w.moves_left_weight != 0.0f) {
            const float m = child.complete_visits == 0 ? 0.0f : child.sum_values.moves_left / float(child.complete_visits) - up.moves_left_m1;
            float m_clipped = std::fmin(std::fmax(m, -w.moves_left_clip), w.moves_left_clip);
            m_unit = std::fmin(std::fmax(w.moves_left_sharpness * m_clipped * -q, -1.0f), 1.0f);
        }
        return q + w.exploration_weight * u + w.moves_left_weight * m_unit;
    }

    void propagate(int node, ValuesAbs values) {  // step.rs:171-188
        int cur = node;
        while (true) {
            Node& n = nodes[cur];
            if (n.virtual_visits == 0) throw std::logic_error("propagate: node has no virtual visit");
            n.complete_visits += 1;
            n.virtual_visits -= 1;
            n.sum_values.add(values);
            if (n.parent < 0) break;
            cur = n.parent;
            values = values.parent();
        }
    }

    // tree.rs:132-141: visit distribution over the root's children
    void policy(std::vector<float>& out) const {
        const Node& r = nodes[0];
        out.resize(size_t(r.child_count));
        const float denom = std::fmax(float(r.complete_visits) - 1.0f, 0.0f);
        for (int i = 0; i < r.child_count; i++) out[size_t(i)] = float(nodes[r.child_start + i].complete_visits) / denom;
    }
};

template <typename Game>
struct Request {
    int node = -1;
    Game board;
};

// step.rs:61-135.  Returns true and fills `req` when an un-evaluated node was reached; false when a terminal node
// was reached (its outcome has been propagated).
template <typename Game>
bool zero_step_gather(Tree<Game>& tree, const SearchSettings& s, Rng& rng, Request<Game>& req, std::vector<uint32_t>& scratch) {
    int cur = 0;
    Game board = tree.root_board;
    while (true) {
        tree.nodes[cur].virtual_visits += 1;
        if (board.done()) {
            tree.propagate(cur, ValuesAbs::from_outcome(board.outcome(), 0.0f));
            return false;
        }
        if (tree.nodes[cur].child_start < 0) {
            // initialise the children with a uniform policy, step.rs:84-103
            board.moves(scratch);
            const float p = 1.0f / float(scratch.size());
            const int start = int(tree.nodes.size());
            for (uint32_t mv : scratch) {
                Node c;
                c.parent = cur;
                c.last_move = mv;
                c.net_policy = p;
                tree.nodes.push_back(c);
            }
            tree.nodes[cur].child_start = start;
            tree.nodes[cur].child_count = int(scratch.size());
            tree.nodes[cur].has_net_values = false;
            req.node = cur;
            req.board = board;
            return true;
        }
        const Node& n = tree.nodes[cur];
        const int player = board.next_player();
        int selected = -1;
        uint32_t ties = 0;
        if (n.complete_visits == 0) {
            // a random least-visited child, step.rs:112-114 (choose_max_by_key over Reverse(total_visits))
            uint64_t best = 0;
            for (int c = n.child_start; c < n.child_start + n.child_count; c++) {
                const uint64_t v = tree.nodes[c].total_visits();
                if (selected < 0 || v < best) {
                    selected = c;
                    best = v;
                    ties = 1;
                } else if (v == best) {
                    ties++;
                    if (rng.gen_range(ties) == 0) selected = c;
                }
            }
        } else {
            const FpuMode fpu = cur == 0 ? s.fpu_root : s.fpu_child;
            const UctContext ctx = tree.uct_context(cur);
            if (ctx.total_visits == 0) throw std::runtime_error("uct is NaN");  // node.rs:171-173
            const auto up = tree.uct_parent(ctx, fpu, s, player);
            float best = 0.0f;
            for (int c = n.child_start; c < n.child_start + n.child_count; c++) {
                const float u = tree.uct_total_fast(tree.nodes[c], up, s, player);
                if (std::isnan(u)) throw std::runtime_error("uct is NaN");  // N32::from_inner panics on NaN
                if (selected < 0 || u > best) {
                    selected = c;
                    best = u;
                    ties = 1;
                } else if (u == best) {
                    ties++;
                    if (rng.gen_range(ties) == 0) selected = c;
                }
            }
        }
        if (selected < 0) throw std::logic_error("Board is not done, this node should have a child");
        cur = selected;
        board.play(tree.nodes[cur].last_move);
    }
}

// step.rs:140-167.  `policy` has one entry per child, in available_moves order.
template <typename Game>
void zero_step_apply(Tree<Game>& tree, int node, int next_player, const ValuesPov& values, const float* policy, size_t n_policy) {
    Node& n = tree.nodes[node];
    if (n.has_net_values) throw std::logic_error("Node was already evaluated by the network");
    const ValuesAbs abs = un_pov(values, next_player);
    n.has_net_values = true;
    tree.propagate(node, abs);
    if (n.child_start < 0) throw std::logic_error("Applied node should have initialized children");
    if (size_t(n.child_count) != n_policy) throw std::logic_error("Wrong children length");
    for (int i = 0; i < n.child_count; i++) tree.nodes[n.child_start + i].net_policy = policy[i];
}

// rust/kz-core/src/network/common.rs:133-163
inline void policy_softmax_temperature_in_place(float* p, size_t n, float temperature) {
    if (temperature == 1.0f) return;
    if (!(temperature > 0.0f) || !std::isfinite(temperature)) throw std::runtime_error("Temperature must be finite and positive");
    float sum = 0.0f;
    for (size_t i = 0; i < n; i++) {
        p[i] = std::pow(p[i], 1.0f / temperature);
        sum += p[i];
    }
    for (size_t i = 0; i < n; i++) p[i] /= sum;
}

// Marsaglia-Tsang gamma(alpha, 1) on top of Rng; alpha < 1 through the alpha+1 boost
inline double sample_gamma(Rng& rng, double alpha) {
    if (alpha < 1.0) {
        const double u = rng.uniform();
        return sample_gamma(rng, alpha + 1.0) * std::pow(u > 0 ? u : 1e-300, 1.0 / alpha);
    }
    const double d = alpha - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
    while (true) {
        double x, v;
        do {
            // Box-Muller normal
            const double u1 = rng.uniform(), u2 = rng.uniform();
            x = std::sqrt(-2.0 * std::log(u1 > 0 ? u1 : 1e-300)) * std::cos(6.283185307179586 * u2);
            v = 1.0 + c * x;
        } while (v <= 0.0);
        v = v * v * v;
        const double u = rng.uniform();
        if (u < 1.0 - 0.0331 * x * x * x * x) return d * v;
        if (std::log(u > 0 ? u : 1e-300) < 0.5 * x * x + d * (1.0 - v + std::log(v))) return d * v;
    }
}

// rust/kz-selfplay/src/server/generator_alphazero.rs:247-259 + kz-util/src/stable_dirichlet.rs:30-66
inline void add_dirichlet_noise(float* policy, size_t n, float alpha, float eps, Rng& rng) {
    if (n <= 1 || eps == 0.0f) return;
    std::vector<float> noise(n, 0.0f);
    bool ok = false;
    if (alpha > 0.1f) {
        double sum = 0.0;
        for (size_t i = 0; i < n; i++) {
            noise[i] = float(sample_gamma(rng, alpha));
            sum += noise[i];
        }
        if (sum > 1e-8) {
            for (auto& v : noise) v = float(v / sum);
            ok = true;
        }
    }
    if (!ok) {  // maximally concentrated sample
        std::fill(noise.begin(), noise.end(), 0.0f);
        noise[rng.gen_range(uint32_t(n))] = 1.0f;
    }
    for (size_t i = 0; i < n; i++) policy[i] = (1.0f - eps) * policy[i] + eps * noise[i];
}

// rust/kz-selfplay/src/move_selector.rs:39-60
inline size_t select_move(const float* policy, size_t n, uint32_t move_count, float temperature, uint32_t zero_temp_move_count, Rng& rng) {
    const float t = move_count >= zero_temp_move_count ? 0.0f : temperature;
    if (t == 0.0f) {
        size_t best = 0;
        for (size_t i = 1; i < n; i++)
            if (policy[i] > policy[best]) best = i;
        return best;
    }
    if (std::isinf(t)) return rng.gen_range(uint32_t(n));
    double total = 0.0;
    std::vector<double> w(n);
    for (size_t i = 0; i < n; i++) {
        w[i] = std::pow(double(policy[i]), 1.0 / t);
        total += w[i];
    }
    double r = rng.uniform() * total;
    for (size_t i = 0; i < n; i++) {
        r -= w[i];
        if (r < 0) return i;
    }
    return n - 1;
}

}  // namespace selfplay
}  // namespace kzb
