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
#include <cstring>
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

// Node storage.  The reference keeps one 64..88-byte Node per child (node.rs:11-34), and most of them are never visited:
// a node's children are created together (step.rs:89-97) but a search of V visits touches only about V of them.  So the
// tree is split in two:
//   child slots   one per created child, ids consecutive per parent (id 0 = the root): last_move and net_policy only
//   visited pool  one 64-byte (one cache line) entry per node that has been visited at least once: the visit counters,
//                 the value sums and the links; entry 0 is an all-zero sentinel that stands for every unvisited child
//   visited lists per pool entry, the (pool index, child position) pairs of its visited children, sorted by position,
//                 in one arena that grows by doubling
// A selection step therefore reads the children's policy slice, the parent's visited list and one line per VISITED
// child, instead of seven statistics slices over all children; pool and lists (about 100 KB for an 800-visit search)
// stay cache resident while a generator thread rotates over dozens of trees.  Unvisited children all share q = fpu, so their uct is one multiply-add
// chain over the policy slice; visited children go through the same scalar formula as before.  The IEEE operations
// and their order are those of the reference's Node::uct, so trees stay bit-identical to the oracle's.
struct UctContext {  // node.rs:55-64
    uint64_t total_visits;
    ValuesAbs values;
    float visited_policy_mass;
};

struct alignas(64) Visited {
    uint32_t complete = 0, virt = 0;                          // complete_visits, virtual_visits
    float value = 0, win_a = 0, draw = 0, win_b = 0, ml = 0;  // sum_values (abs)
    int32_t parent = -1;                                      // pool index of the parent, -1 for the root
    int32_t child_start = -1, child_count = 0;                // child slots; children == None  <=>  child_start < 0
    int32_t vis_off = 0;                                      // visited list: offset into the arena,
    uint16_t vis_count = 0, vis_cap = 0;                      // entries used / reserved
    uint8_t has_net_values = 0;
};
struct VisRef {
    int32_t idx, pos;  // pool index of a visited child, its position among the parent's children
};

namespace detail {
struct UctParent {  // the per-parent part of Node::uct, computed once per selection step
    float fpu, sqrt_visits, moves_left_m1;
};
// node.rs:163-206 + Uct::total :87-98 for one child
inline float uct_one(const Visited& v, float policy, const UctParent& up, const SearchSettings& s, int player) {
    const float vl = s.virtual_loss;
    const float cv = float(v.complete), vv = float(v.virt);
    const float tvv = cv + vl * vv;
    float q;
    if (tvv == 0.0f) {
        q = up.fpu;
    } else {
        float total_value;
        if (s.q_mode.wdl) total_value = player == 0 ? v.win_a + s.q_mode.draw_score * v.draw - v.win_b
                                                    : v.win_b + s.q_mode.draw_score * v.draw - v.win_a;
        else total_value = player == 0 ? v.value : -v.value;
        q = (total_value - vl * vv) / tvv;
    }
    const float u = policy * up.sqrt_visits / float(1u + v.complete + v.virt);
    const UctWeights& w = s.weights;
    float m_unit = 0.0f;
    if (w.moves_left_weight != 0.0f) {
        const float m = v.complete == 0 ? 0.0f : v.ml / cv - up.moves_left_m1;
        const float m_clipped = std::fmin(std::fmax(m, -w.moves_left_clip), w.moves_left_clip);
        m_unit = std::fmin(std::fmax(w.moves_left_sharpness * m_clipped * -q, -1.0f), 1.0f);
    }
    return q + w.exploration_weight * u + w.moves_left_weight * m_unit;
}
#if defined(__x86_64__)
// The same formula for 8 visited children at a time: the first 32 bytes of their pool entries (complete, virt, value,
// win_a, draw, win_b, ml, parent) are loaded as one row each and transposed into columns.  Same IEEE operations in the
// same order as uct_one.  `idx` / `policy` are padded to a multiple of 8 (index 0 = the sentinel).
__attribute__((target("avx2"))) inline void uct_visited_avx2(const Visited* pool, const int32_t* idx, const float* policy, int k,
                                                             const UctParent& up, const SearchSettings& s, int player, float* out) {
    const __m256 vl = _mm256_set1_ps(s.virtual_loss), fpu = _mm256_set1_ps(up.fpu), sq = _mm256_set1_ps(up.sqrt_visits);
    const __m256 mlm1 = _mm256_set1_ps(up.moves_left_m1), ds = _mm256_set1_ps(s.q_mode.draw_score), zero = _mm256_setzero_ps();
    const UctWeights& w = s.weights;
    const __m256 ew = _mm256_set1_ps(w.exploration_weight), mw = _mm256_set1_ps(w.moves_left_weight);
    const __m256 clip = _mm256_set1_ps(w.moves_left_clip), nclip = _mm256_set1_ps(-w.moves_left_clip), sharp = _mm256_set1_ps(w.moves_left_sharpness);
    const __m256 one = _mm256_set1_ps(1.0f), none = _mm256_set1_ps(-1.0f);
    for (int g = 0; g < k; g += 8) {
        __m256 r[8];
        for (int j = 0; j < 8; j++) r[j] = _mm256_load_ps(reinterpret_cast<const float*>(pool + idx[g + j]));
        // 8x8 transpose
        const __m256 t0 = _mm256_unpacklo_ps(r[0], r[1]), t1 = _mm256_unpackhi_ps(r[0], r[1]);
        const __m256 t2 = _mm256_unpacklo_ps(r[2], r[3]), t3 = _mm256_unpackhi_ps(r[2], r[3]);
        const __m256 t4 = _mm256_unpacklo_ps(r[4], r[5]), t5 = _mm256_unpackhi_ps(r[4], r[5]);
        const __m256 t6 = _mm256_unpacklo_ps(r[6], r[7]), t7 = _mm256_unpackhi_ps(r[6], r[7]);
        const __m256 u0 = _mm256_shuffle_ps(t0, t2, 0x44), u1 = _mm256_shuffle_ps(t0, t2, 0xEE);
        const __m256 u2 = _mm256_shuffle_ps(t1, t3, 0x44), u3 = _mm256_shuffle_ps(t1, t3, 0xEE);
        const __m256 u4 = _mm256_shuffle_ps(t4, t6, 0x44), u5 = _mm256_shuffle_ps(t4, t6, 0xEE);
        const __m256 u6 = _mm256_shuffle_ps(t5, t7, 0x44), u7 = _mm256_shuffle_ps(t5, t7, 0xEE);
        const __m256i cvi = _mm256_castps_si256(_mm256_permute2f128_ps(u0, u4, 0x20));  // complete
        const __m256i vvi = _mm256_castps_si256(_mm256_permute2f128_ps(u1, u5, 0x20));  // virt
        const __m256 value = _mm256_permute2f128_ps(u2, u6, 0x20), win_a = _mm256_permute2f128_ps(u3, u7, 0x20);
        const __m256 draw = _mm256_permute2f128_ps(u0, u4, 0x31), win_b = _mm256_permute2f128_ps(u1, u5, 0x31);
        const __m256 ml = _mm256_permute2f128_ps(u2, u6, 0x31);
        const __m256 cv = _mm256_cvtepi32_ps(cvi), vv = _mm256_cvtepi32_ps(vvi);
        const __m256 vlvv = _mm256_mul_ps(vl, vv);
        const __m256 tvv = _mm256_add_ps(cv, vlvv);
        __m256 total_value;
        if (s.q_mode.wdl) {
            const __m256 own = player == 0 ? win_a : win_b, opp = player == 0 ? win_b : win_a;
            total_value = _mm256_sub_ps(_mm256_add_ps(own, _mm256_mul_ps(ds, draw)), opp);
        } else {
            total_value = value;
            if (player != 0) total_value = _mm256_sub_ps(zero, total_value);  // -x == 0 - x except for the sign of zero, which no later step sees
        }
        __m256 q = _mm256_div_ps(_mm256_sub_ps(total_value, vlvv), tvv);
        q = _mm256_blendv_ps(q, fpu, _mm256_cmp_ps(tvv, zero, _CMP_EQ_OQ));
        const __m256 denom = _mm256_cvtepi32_ps(_mm256_add_epi32(_mm256_add_epi32(cvi, vvi), _mm256_set1_epi32(1)));
        const __m256 u = _mm256_div_ps(_mm256_mul_ps(_mm256_loadu_ps(policy + g), sq), denom);
        __m256 total = _mm256_add_ps(q, _mm256_mul_ps(ew, u));
        if (w.moves_left_weight != 0.0f) {
            __m256 m = _mm256_sub_ps(_mm256_div_ps(ml, cv), mlm1);
            m = _mm256_blendv_ps(m, zero, _mm256_castsi256_ps(_mm256_cmpeq_epi32(cvi, _mm256_setzero_si256())));
            const __m256 m_clipped = _mm256_min_ps(_mm256_max_ps(m, nclip), clip);
            const __m256 m_unit = _mm256_min_ps(_mm256_max_ps(_mm256_mul_ps(_mm256_mul_ps(sharp, m_clipped), _mm256_sub_ps(zero, q)), none), one);
            total = _mm256_add_ps(total, _mm256_mul_ps(mw, m_unit));
        }
        _mm256_storeu_ps(out + g, total);
    }
}
#endif
// uct of children that have never been visited: the formula above with zero counters, where q and the moves-left term
// are per-parent constants.  `x / 1.0f` is exact, so it is dropped.
#define KZB_UCT_UNVISITED_BODY                                                      \
    for (int i = 0; i < n; i++) out[i] = (q + ew * (policy[i] * sq)) + ml_term;
inline void uct_unvisited_generic(const float* policy, int n, float q, float ew, float sq, float ml_term, float* out) { KZB_UCT_UNVISITED_BODY }
#if defined(__x86_64__)
__attribute__((target("avx2"))) inline void uct_unvisited_avx2(const float* policy, int n, float q, float ew, float sq, float ml_term,
                                                               float* __restrict__ out) { KZB_UCT_UNVISITED_BODY }
#endif
#undef KZB_UCT_UNVISITED_BODY
inline void uct_unvisited(const float* policy, int n, const UctParent& up, const SearchSettings& s, float* out) {
    const float q = up.fpu;
    const UctWeights& w = s.weights;
    float m_unit = 0.0f;
    if (w.moves_left_weight != 0.0f) {
        const float m_clipped = std::fmin(std::fmax(0.0f, -w.moves_left_clip), w.moves_left_clip);
        m_unit = std::fmin(std::fmax(w.moves_left_sharpness * m_clipped * -q, -1.0f), 1.0f);
    }
    const float ml_term = w.moves_left_weight * m_unit;
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) return uct_unvisited_avx2(policy, n, q, w.exploration_weight, up.sqrt_visits, ml_term, out);
#endif
    uct_unvisited_generic(policy, n, q, w.exploration_weight, up.sqrt_visits, ml_term, out);
}
// The scalar rule; `best` / `arg` / `ties` carry the running state so that a caller can skip elements it knows to be
// smaller than the running best.
inline void argmax_scan(const float* u, int begin, int end, float& best, int& arg, uint32_t& ties, bool& nan, Rng& rng) {
    for (int i = begin; i < end; i++) {
        const float x = u[i];
        nan |= x != x;
        if (__builtin_expect(x == best, 0)) {
            ties++;
            if (rng.gen_range(ties) == 0) arg = i;
            continue;
        }
        const bool gt = x > best;  // selects, not a branch: "new best" is unpredictable
        best = gt ? x : best;
        arg = gt ? i : arg;
        ties = gt ? 1u : ties;
    }
}
#if defined(__x86_64__)
// 8 elements at a time: a block with no element >= the running best (and no NaN) cannot change anything and is skipped
__attribute__((target("avx2"))) inline void argmax_blocks_avx2(const float* u, int begin, int n, float& best, int& arg, uint32_t& ties, bool& nan, Rng& rng) {
    int i = begin;
    for (; i + 8 <= n; i += 8) {
        const __m256 x = _mm256_loadu_ps(u + i);
        if (_mm256_movemask_ps(_mm256_cmp_ps(x, _mm256_set1_ps(best), _CMP_NLT_UQ)) == 0) continue;  // NLT_UQ: x >= best or unordered
        argmax_scan(u, i, i + 8, best, arg, ties, nan, rng);
    }
    argmax_scan(u, i, n, best, arg, ties, nan, rng);
}
#endif
inline int argmax_random_ties(const float* u, int n, Rng& rng) {
    float best = u[0];
    int arg = 0;
    uint32_t ties = 1;
    bool nan = u[0] != u[0];
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) argmax_blocks_avx2(u, 1, n, best, arg, ties, nan, rng);
    else
#endif
        argmax_scan(u, 1, n, best, arg, ties, nan, rng);
    if (nan) throw std::runtime_error("uct is NaN");  // N32::from_inner panics on NaN
    return arg;
}
}  // namespace detail

template <typename Game>
struct Tree {
    Game root_board;
    // child slots
    std::vector<uint32_t> last_move;
    std::vector<float> net_policy;
    // visited pool; [0] is the sentinel, [1] the root
    std::vector<Visited> pool;
    std::vector<VisRef> vis_arena;
    std::vector<float> uct_scratch, vis_policy, vis_out;
    std::vector<int32_t> vis_idx, vis_pos;
    std::vector<uint32_t> visits_scratch;
    static constexpr int kRoot = 1;  // pool index of the root

    explicit Tree(const Game& root) : root_board(root) {  // tree.rs:31-39
        if (root.done()) throw std::runtime_error("Cannot build tree for done board");
        last_move.push_back(0), net_policy.push_back(NAN);
        pool.resize(2);
    }
    size_t size() const { return last_move.size(); }  // nodes in the reference's sense: the root and every created child
    void reserve(size_t slots, size_t visited) {
        last_move.reserve(slots), net_policy.reserve(slots);
        pool.reserve(visited + 2), vis_arena.reserve(4 * visited + 64);
    }
    // all children of a node at once (step.rs:89-97), with a uniform prior
    int push_children(const std::vector<uint32_t>& moves, float p) {
        const size_t start = last_move.size(), end = start + moves.size();
        net_policy.resize(end, p);
        last_move.insert(last_move.end(), moves.begin(), moves.end());
        return int(start);
    }
    // the pool entry of the child at position `pos` of `parent`, created (and entered into the parent's visited list,
    // which stays sorted by position) on its first visit
    int visit_child(int parent, int pos) {
        {
            const Visited& p = pool[size_t(parent)];
            const VisRef* list = vis_arena.data() + p.vis_off;
            for (int j = 0; j < p.vis_count; j++)
                if (list[j].pos == pos) return list[j].idx;
        }
        const int v = int(pool.size());
        pool.emplace_back();
        pool.back().parent = parent;
        Visited& p = pool[size_t(parent)];
        if (p.vis_count == p.vis_cap) {  // move the list to the end of the arena with twice the room
            const int cap = p.vis_cap ? 2 * int(p.vis_cap) : 4;
            const size_t off = vis_arena.size();
            vis_arena.resize(off + size_t(cap));
            std::memcpy(vis_arena.data() + off, vis_arena.data() + p.vis_off, size_t(p.vis_count) * sizeof(VisRef));
            p.vis_off = int32_t(off);
            p.vis_cap = uint16_t(cap);
        }
        VisRef* list = vis_arena.data() + p.vis_off;
        int j = p.vis_count;
        for (; j > 0 && list[j - 1].pos > pos; j--) list[j] = list[j - 1];
        list[j] = {v, pos};
        p.vis_count++;
        return v;
    }
    // complete visits of every child of a node, in child order
    void child_visits(int node, std::vector<uint32_t>& out) const {
        const Visited& p = pool[size_t(node)];
        out.assign(size_t(p.child_count), 0u);
        const VisRef* list = vis_arena.data() + p.vis_off;
        for (int j = 0; j < p.vis_count; j++) out[size_t(list[j].pos)] = pool[size_t(list[j].idx)].complete;
    }
    const Visited& root() const { return pool[kRoot]; }
    uint64_t root_visits() const { return pool[kRoot].complete; }
    static ValuesAbs sum_values(const Visited& v) { return {v.value, v.win_a, v.draw, v.win_b, v.ml}; }
    static ValuesAbs values(const Visited& v) { return sum_values(v).div(float(v.complete)); }  // node.rs:126-128
    ValuesAbs root_values() const { return values(pool[kRoot]); }

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

    void propagate(int node, ValuesAbs v) {  // step.rs:171-188
        int cur = node;
        while (true) {
            Visited& n = pool[size_t(cur)];
            if (n.virt == 0) throw std::logic_error("propagate: node has no virtual visit");
            n.complete += 1;
            n.virt -= 1;
            n.value += v.value, n.win_a += v.win_a, n.draw += v.draw, n.win_b += v.win_b, n.ml += v.moves_left;
            if (n.parent < 0) break;
            cur = n.parent;
            v = v.parent();
        }
    }

    // tree.rs:132-141: visit distribution over the root's children
    void policy(std::vector<float>& out) const {
        const Visited& r = pool[kRoot];
        const float denom = std::fmax(float(r.complete) - 1.0f, 0.0f);
        out.assign(size_t(r.child_count), 0.0f / denom);
        const VisRef* list = vis_arena.data() + r.vis_off;
        for (int j = 0; j < r.vis_count; j++) out[size_t(list[j].pos)] = float(pool[size_t(list[j].idx)].complete) / denom;
    }
};

template <typename Game>
struct Request {
    int node = -1;  // pool index of the node to evaluate
    int child_start = 0, child_count = 0;  // its freshly created child slots
    Game board;
    bool is_root() const { return node == Tree<Game>::kRoot; }
};

// One in-flight zero_step_gather (step.rs:61-135), advanced one tree level per descent_step call so that a caller can
// interleave the descents of several trees: each step ends by prefetching the child slices the next step will scan, and
// the cache misses of one tree overlap the arithmetic of the others.
template <typename Game>
struct Descent {
    int cur = Tree<Game>::kRoot;
    Game board;
    void begin(const Tree<Game>& tree) {
        cur = Tree<Game>::kRoot;
        board = tree.root_board;
    }
};
enum class StepResult { kDescend, kRequest, kTerminal };

inline void prefetch_span(const void* p, size_t bytes) {
    const char* c = static_cast<const char*>(p);
    for (size_t o = 0; o < bytes + 63; o += 64) __builtin_prefetch(c + o);
}

// kRequest: an un-evaluated node was reached and `req` is filled; kTerminal: a terminal node was reached and its outcome
// has been propagated; kDescend: moved one level down, call again.
struct NoLeafHook {
    template <typename Game>
    void operator()(const Game&) const {}
};
// `on_leaf(board)` runs when an un-evaluated node is reached, before its children are created: the caller's chance to
// start fetching whatever it will look up for this board (the evaluation cache) while the expansion still has work to do.
template <typename Game, typename LeafHook = NoLeafHook>
StepResult descent_step(Tree<Game>& tree, const SearchSettings& s, Rng& rng, Descent<Game>& d, Request<Game>& req, std::vector<uint32_t>& scratch,
                        LeafHook on_leaf = LeafHook()) {
    int& cur = d.cur;
    Game& board = d.board;
    {
        tree.pool[size_t(cur)].virt += 1;
        if (board.done()) {
            tree.propagate(cur, ValuesAbs::from_outcome(board.outcome(), 0.0f));
            return StepResult::kTerminal;
        }
        if (tree.pool[size_t(cur)].child_start < 0) {
            // initialise the children with a uniform policy, step.rs:84-103
            on_leaf(board);
            board.moves(scratch);
            const float p = 1.0f / float(scratch.size());
            const int start = tree.push_children(scratch, p);
            Visited& n = tree.pool[size_t(cur)];
            n.child_start = start;
            n.child_count = int(scratch.size());
            n.has_net_values = 0;
            req.node = cur;
            req.child_start = start;
            req.child_count = n.child_count;
            req.board = board;
            return StepResult::kRequest;
        }
        const Visited& pn = tree.pool[size_t(cur)];
        const int c0 = pn.child_start, n = pn.child_count;
        const VisRef* vis = tree.vis_arena.data() + pn.vis_off;
        const int k = pn.vis_count;
        const int player = board.next_player();
        int selected = -1;
        uint32_t ties = 0;
        if (pn.complete == 0) {
            // a random least-visited child, step.rs:112-114 (choose_max_by_key over Reverse(total_visits))
            tree.visits_scratch.assign(size_t(n), 0u);
            for (int j = 0; j < k; j++) tree.visits_scratch[size_t(vis[j].pos)] = tree.pool[size_t(vis[j].idx)].complete + tree.pool[size_t(vis[j].idx)].virt;
            uint64_t best = 0;
            for (int i = 0; i < n; i++) {
                const uint64_t v = tree.visits_scratch[size_t(i)];
                if (selected < 0 || v < best) {
                    selected = c0 + i;
                    best = v;
                    ties = 1;
                } else if (v == best) {
                    ties++;
                    if (rng.gen_range(ties) == 0) selected = c0 + i;
                }
            }
        } else {
            const FpuMode fpu = cur == Tree<Game>::kRoot ? s.fpu_root : s.fpu_child;
            // the visited children, in child order (the order uct_context sums the policy mass in, tree.rs:49-66)
            const float* policy = tree.net_policy.data() + c0;
            if (tree.uct_scratch.size() < size_t(n) + 8) tree.uct_scratch.resize(size_t(n) + 8), tree.vis_policy.resize(size_t(n) + 8), tree.vis_out.resize(size_t(n) + 8), tree.vis_idx.resize(size_t(n) + 8), tree.vis_pos.resize(size_t(n) + 8);
            for (int j = 0; j < k; j++) {
                tree.vis_idx[size_t(j)] = vis[j].idx, tree.vis_pos[size_t(j)] = vis[j].pos;
                tree.vis_policy[size_t(j)] = policy[vis[j].pos];
            }
            float mass = 0.0f;
            for (int j = 0; j < k; j++) {
                const Visited& ch = tree.pool[size_t(tree.vis_idx[size_t(j)])];
                if (ch.complete + ch.virt > 0) mass += tree.vis_policy[size_t(j)];
            }
            const UctContext ctx{uint64_t(pn.complete) + pn.virt, Tree<Game>::values(pn), mass};
            if (ctx.total_visits == 0) throw std::runtime_error("uct is NaN");  // node.rs:171-173
            const detail::UctParent up = tree.uct_parent(ctx, fpu, s, player);
            float* u = tree.uct_scratch.data();
            detail::uct_unvisited(policy, n, up, s, u);
#if defined(__x86_64__)
            static const bool have_avx2 = __builtin_cpu_supports("avx2");
            if (have_avx2 && k > 2) {
                for (int j = k; j < ((k + 7) & ~7); j++) tree.vis_idx[size_t(j)] = 0, tree.vis_policy[size_t(j)] = 0.0f;
                detail::uct_visited_avx2(tree.pool.data(), tree.vis_idx.data(), tree.vis_policy.data(), k, up, s, player, tree.vis_out.data());
                for (int j = 0; j < k; j++) u[tree.vis_pos[size_t(j)]] = tree.vis_out[size_t(j)];
            } else
#endif
                for (int j = 0; j < k; j++)
                    u[tree.vis_pos[size_t(j)]] = detail::uct_one(tree.pool[size_t(tree.vis_idx[size_t(j)])], tree.vis_policy[size_t(j)], up, s, player);
            // choose_max_by_key with random tie break, kz-util/src/sequence.rs:11-41: a later element replaces the running
            // best when it is greater, and with probability 1/ties when it is equal.
            selected = c0 + detail::argmax_random_ties(u, n, rng);
        }
        if (selected < 0) throw std::logic_error("Board is not done, this node should have a child");
        board.play(tree.last_move[size_t(selected)]);
        cur = tree.visit_child(cur, selected - c0);
        return StepResult::kDescend;
    }
}

// step.rs:61-135 in one go.  Returns true and fills `req` when an un-evaluated node was reached; false when a terminal
// node was reached (its outcome has been propagated).
template <typename Game, typename LeafHook = NoLeafHook>
bool zero_step_gather(Tree<Game>& tree, const SearchSettings& s, Rng& rng, Request<Game>& req, std::vector<uint32_t>& scratch,
                      LeafHook on_leaf = LeafHook()) {
    Descent<Game> d;
    d.begin(tree);
    while (true) {
        const StepResult r = descent_step(tree, s, rng, d, req, scratch, on_leaf);
        if (r != StepResult::kDescend) return r == StepResult::kRequest;
    }
}

// step.rs:140-167.  `policy` has one entry per child, in available_moves order.
template <typename Game>
void zero_step_apply(Tree<Game>& tree, int node, int next_player, const ValuesPov& values, const float* policy, size_t n_policy) {
    Visited& n = tree.pool[size_t(node)];
    if (n.has_net_values) throw std::logic_error("Node was already evaluated by the network");
    n.has_net_values = 1;
    if (n.child_start < 0) throw std::logic_error("Applied node should have initialized children");
    if (size_t(n.child_count) != n_policy) throw std::logic_error("Wrong children length");
    std::memcpy(tree.net_policy.data() + n.child_start, policy, n_policy * sizeof(float));
    tree.propagate(node, un_pov(values, next_player));
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
