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
#if defined(__x86_64__)
#include <immintrin.h>
#endif
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

// Node storage (node.rs:11-34) is a structure of arrays indexed by node id: the children of a node have consecutive
// ids, so a selection step reads a handful of contiguous float / u32 slices instead of one 64..88-byte struct per
// child, and the UCT scan over them vectorises (AVX2, 8 children per step; same IEEE operations in the same order as
// the scalar code, so trees stay bit-identical to the oracle's).  `net_values` is only kept as a flag: the reference
// stores the network's own values to write them into game records, which this driver does not produce.
struct UctContext {  // node.rs:55-64
    uint64_t total_visits;
    ValuesAbs values;
    float visited_policy_mass;
};

namespace detail {
struct UctParent {  // the per-parent part of Node::uct, computed once per selection step
    float fpu, sqrt_visits, moves_left_m1;
};
struct UctArrays {
    const uint32_t *complete, *virt;
    const float *value, *win_a, *draw, *win_b, *ml, *policy;
};
// node.rs:163-206 + Uct::total :87-98 for child i (scalar reference form)
inline float uct_one(const UctArrays& a, int i, const UctParent& up, const SearchSettings& s, int player) {
    const float vl = s.virtual_loss;
    const float cv = float(a.complete[i]), vv = float(a.virt[i]);
    const float tvv = cv + vl * vv;
    float q;
    if (tvv == 0.0f) {
        q = up.fpu;
    } else {
        float total_value;
        if (s.q_mode.wdl) total_value = player == 0 ? a.win_a[i] + s.q_mode.draw_score * a.draw[i] - a.win_b[i]
                                                    : a.win_b[i] + s.q_mode.draw_score * a.draw[i] - a.win_a[i];
        else total_value = player == 0 ? a.value[i] : -a.value[i];
        q = (total_value - vl * vv) / tvv;
    }
    const float u = a.policy[i] * up.sqrt_visits / float(1u + a.complete[i] + a.virt[i]);
    const UctWeights& w = s.weights;
    float m_unit = 0.0f;
    if (w.moves_left_weight != 0.0f) {
        const float m = a.complete[i] == 0 ? 0.0f : a.ml[i] / cv - up.moves_left_m1;
        const float m_clipped = std::fmin(std::fmax(m, -w.moves_left_clip), w.moves_left_clip);
        m_unit = std::fmin(std::fmax(w.moves_left_sharpness * m_clipped * -q, -1.0f), 1.0f);
    }
    return q + w.exploration_weight * u + w.moves_left_weight * m_unit;
}
#if defined(__x86_64__)
__attribute__((target("avx2"))) inline void uct_many_avx2(const UctArrays& a, int n, const UctParent& up, const SearchSettings& s,
                                                          int player, float* out) {
    const __m256 vl = _mm256_set1_ps(s.virtual_loss), fpu = _mm256_set1_ps(up.fpu), sq = _mm256_set1_ps(up.sqrt_visits);
    const __m256 mlm1 = _mm256_set1_ps(up.moves_left_m1), ds = _mm256_set1_ps(s.q_mode.draw_score), zero = _mm256_setzero_ps();
    const UctWeights& w = s.weights;
    const __m256 ew = _mm256_set1_ps(w.exploration_weight), mw = _mm256_set1_ps(w.moves_left_weight);
    const __m256 clip = _mm256_set1_ps(w.moves_left_clip), nclip = _mm256_set1_ps(-w.moves_left_clip), sharp = _mm256_set1_ps(w.moves_left_sharpness);
    const __m256 one = _mm256_set1_ps(1.0f), none = _mm256_set1_ps(-1.0f);
    const float* own = player == 0 ? a.win_a : a.win_b;
    const float* opp = player == 0 ? a.win_b : a.win_a;
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        const __m256i cvi = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(a.complete + i));
        const __m256i vvi = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(a.virt + i));
        const __m256 cv = _mm256_cvtepi32_ps(cvi), vv = _mm256_cvtepi32_ps(vvi);
        const __m256 vlvv = _mm256_mul_ps(vl, vv);
        const __m256 tvv = _mm256_add_ps(cv, vlvv);
        __m256 total_value;
        if (s.q_mode.wdl) {
            total_value = _mm256_sub_ps(_mm256_add_ps(_mm256_loadu_ps(own + i), _mm256_mul_ps(ds, _mm256_loadu_ps(a.draw + i))), _mm256_loadu_ps(opp + i));
        } else {
            total_value = _mm256_loadu_ps(a.value + i);
            if (player != 0) total_value = _mm256_sub_ps(zero, total_value);  // -x == 0 - x except for the sign of zero, which no later step sees
        }
        __m256 q = _mm256_div_ps(_mm256_sub_ps(total_value, vlvv), tvv);
        q = _mm256_blendv_ps(q, fpu, _mm256_cmp_ps(tvv, zero, _CMP_EQ_OQ));
        const __m256 denom = _mm256_cvtepi32_ps(_mm256_add_epi32(_mm256_add_epi32(cvi, vvi), _mm256_set1_epi32(1)));
        const __m256 u = _mm256_div_ps(_mm256_mul_ps(_mm256_loadu_ps(a.policy + i), sq), denom);
        __m256 total = _mm256_add_ps(q, _mm256_mul_ps(ew, u));
        if (w.moves_left_weight != 0.0f) {
            __m256 m = _mm256_sub_ps(_mm256_div_ps(_mm256_loadu_ps(a.ml + i), cv), mlm1);
            m = _mm256_blendv_ps(m, zero, _mm256_castsi256_ps(_mm256_cmpeq_epi32(cvi, _mm256_setzero_si256())));
            const __m256 m_clipped = _mm256_min_ps(_mm256_max_ps(m, nclip), clip);
            const __m256 m_unit = _mm256_min_ps(_mm256_max_ps(_mm256_mul_ps(_mm256_mul_ps(sharp, m_clipped), _mm256_sub_ps(zero, q)), none), one);
            total = _mm256_add_ps(total, _mm256_mul_ps(mw, m_unit));
        }
        _mm256_storeu_ps(out + i, total);
    }
    for (; i < n; i++) out[i] = uct_one(a, i, up, s, player);
}
#endif
inline void uct_many(const UctArrays& a, int n, const UctParent& up, const SearchSettings& s, int player, float* out) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) return uct_many_avx2(a, n, up, s, player, out);
#endif
    for (int i = 0; i < n; i++) out[i] = uct_one(a, i, up, s, player);
}
}  // namespace detail

template <typename Game>
struct Tree {
    Game root_board;
    // structure
    std::vector<int32_t> parent, child_start, child_count;  // children == None  <=>  child_start < 0
    std::vector<uint32_t> last_move;
    std::vector<uint8_t> has_net_values;
    // statistics
    std::vector<uint32_t> complete, virt;                       // complete_visits, virtual_visits
    std::vector<float> s_value, s_win_a, s_draw, s_win_b, s_ml;  // sum_values (abs)
    std::vector<float> net_policy;
    std::vector<float> uct_scratch;

    explicit Tree(const Game& root) : root_board(root) {  // tree.rs:31-39
        if (root.done()) throw std::runtime_error("Cannot build tree for done board");
        push_node(-1, 0, NAN);
    }
    size_t size() const { return parent.size(); }
    void reserve(size_t n) {
        parent.reserve(n), child_start.reserve(n), child_count.reserve(n), last_move.reserve(n), has_net_values.reserve(n);
        complete.reserve(n), virt.reserve(n), s_value.reserve(n), s_win_a.reserve(n), s_draw.reserve(n), s_win_b.reserve(n);
        s_ml.reserve(n), net_policy.reserve(n);
    }
    int push_node(int par, uint32_t mv, float p) {  // Node::new, node.rs:104-118
        parent.push_back(par), child_start.push_back(-1), child_count.push_back(0), last_move.push_back(mv), has_net_values.push_back(0);
        complete.push_back(0), virt.push_back(0);
        s_value.push_back(0), s_win_a.push_back(0), s_draw.push_back(0), s_win_b.push_back(0), s_ml.push_back(0), net_policy.push_back(p);
        return int(parent.size()) - 1;
    }
    // all children of `par` at once (step.rs:89-97): one resize per array instead of a push_back per node and field
    int push_children(int par, const std::vector<uint32_t>& moves, float p) {
        const size_t start = parent.size(), n = moves.size(), end = start + n;
        parent.resize(end, par), child_start.resize(end, -1), child_count.resize(end, 0), has_net_values.resize(end, 0);
        complete.resize(end, 0), virt.resize(end, 0);
        s_value.resize(end, 0.0f), s_win_a.resize(end, 0.0f), s_draw.resize(end, 0.0f), s_win_b.resize(end, 0.0f), s_ml.resize(end, 0.0f);
        net_policy.resize(end, p);
        last_move.insert(last_move.end(), moves.begin(), moves.end());
        return int(start);
    }
    uint64_t root_visits() const { return complete[0]; }
    uint64_t total_visits(int n) const { return uint64_t(complete[size_t(n)]) + virt[size_t(n)]; }
    ValuesAbs sum_values(int n) const { return {s_value[size_t(n)], s_win_a[size_t(n)], s_draw[size_t(n)], s_win_b[size_t(n)], s_ml[size_t(n)]}; }
    ValuesAbs values(int n) const { return sum_values(n).div(float(complete[size_t(n)])); }  // node.rs:126-128

    UctContext uct_context(int node) const {  // tree.rs:49-66, node.rs:153-161
        float mass = 0.0f;
        const int c0 = child_start[size_t(node)], c1 = c0 + child_count[size_t(node)];
        for (int c = c0; c < c1; c++)
            if (complete[size_t(c)] + virt[size_t(c)] > 0) mass += net_policy[size_t(c)];
        return {total_visits(node), values(node), mass};
    }

    detail::UctParent uct_parent(const UctContext& par, FpuMode fpu_mode, const SearchSettings& s, int player) const {
        detail::UctParent u;
        if (fpu_mode.relative) {
            float parent_value = s.q_mode.select(pov(par.values, player));
            u.fpu = parent_value - fpu_mode.value * std::sqrt(par.visited_policy_mass);
        } else {
            u.fpu = fpu_mode.value;
        }
        u.sqrt_visits = std::sqrt(float(par.total_visits - 1));
        u.moves_left_m1 = par.values.moves_left - 1.0f;
        return u;
    }
    detail::UctArrays arrays(int first_child) const {
        const size_t o = size_t(first_child);
        return {complete.data() + o, virt.data() + o, s_value.data() + o, s_win_a.data() + o, s_draw.data() + o, s_win_b.data() + o,
                s_ml.data() + o, net_policy.data() + o};
    }

    void propagate(int node, ValuesAbs v) {  // step.rs:171-188
        int cur = node;
        while (true) {
            const size_t i = size_t(cur);
            if (virt[i] == 0) throw std::logic_error("propagate: node has no virtual visit");
            complete[i] += 1;
            virt[i] -= 1;
            s_value[i] += v.value, s_win_a[i] += v.win_a, s_draw[i] += v.draw, s_win_b[i] += v.win_b, s_ml[i] += v.moves_left;
            if (parent[i] < 0) break;
            cur = parent[i];
            v = v.parent();
        }
    }

    // tree.rs:132-141: visit distribution over the root's children
    void policy(std::vector<float>& out) const {
        out.resize(size_t(child_count[0]));
        const float denom = std::fmax(float(complete[0]) - 1.0f, 0.0f);
        for (int i = 0; i < child_count[0]; i++) out[size_t(i)] = float(complete[size_t(child_start[0] + i)]) / denom;
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
        tree.virt[size_t(cur)] += 1;
        if (board.done()) {
            tree.propagate(cur, ValuesAbs::from_outcome(board.outcome(), 0.0f));
            return false;
        }
        if (tree.child_start[size_t(cur)] < 0) {
            // initialise the children with a uniform policy, step.rs:84-103
            board.moves(scratch);
            const float p = 1.0f / float(scratch.size());
            const int start = tree.push_children(cur, scratch, p);
            tree.child_start[size_t(cur)] = start;
            tree.child_count[size_t(cur)] = int(scratch.size());
            tree.has_net_values[size_t(cur)] = 0;
            req.node = cur;
            req.board = board;
            return true;
        }
        const int c0 = tree.child_start[size_t(cur)], n = tree.child_count[size_t(cur)];
        const int player = board.next_player();
        int selected = -1;
        uint32_t ties = 0;
        if (tree.complete[size_t(cur)] == 0) {
            // a random least-visited child, step.rs:112-114 (choose_max_by_key over Reverse(total_visits))
            uint64_t best = 0;
            for (int c = c0; c < c0 + n; c++) {
                const uint64_t v = tree.total_visits(c);
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
            const detail::UctParent up = tree.uct_parent(ctx, fpu, s, player);
            if (tree.uct_scratch.size() < size_t(n)) tree.uct_scratch.resize(size_t(n));
            float* u = tree.uct_scratch.data();
            detail::uct_many(tree.arrays(c0), n, up, s, player, u);
            float best = 0.0f;
            for (int i = 0; i < n; i++) {  // choose_max_by_key with random tie break, kz-util/src/sequence.rs:11-41
                if (std::isnan(u[i])) throw std::runtime_error("uct is NaN");  // N32::from_inner panics on NaN
                if (selected < 0 || u[i] > best) {
                    selected = c0 + i;
                    best = u[i];
                    ties = 1;
                } else if (u[i] == best) {
                    ties++;
                    if (rng.gen_range(ties) == 0) selected = c0 + i;
                }
            }
        }
        if (selected < 0) throw std::logic_error("Board is not done, this node should have a child");
        cur = selected;
        board.play(tree.last_move[size_t(cur)]);
    }
}

// step.rs:140-167.  `policy` has one entry per child, in available_moves order.
template <typename Game>
void zero_step_apply(Tree<Game>& tree, int node, int next_player, const ValuesPov& values, const float* policy, size_t n_policy) {
    const size_t i = size_t(node);
    if (tree.has_net_values[i]) throw std::logic_error("Node was already evaluated by the network");
    tree.has_net_values[i] = 1;
    tree.propagate(node, un_pov(values, next_player));
    if (tree.child_start[i] < 0) throw std::logic_error("Applied node should have initialized children");
    if (size_t(tree.child_count[i]) != n_policy) throw std::logic_error("Wrong children length");
    for (int c = 0; c < tree.child_count[i]; c++) tree.net_policy[size_t(tree.child_start[i] + c)] = policy[c];
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
