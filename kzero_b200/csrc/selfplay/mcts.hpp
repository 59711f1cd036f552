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
#include <algorithm>
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
// tree is split in three:
//   child slots     per node ONE slice of `child_data`: the children's net_policy (n floats) with their moves (n uint16) right behind
//                   it, so that the move of the child a selection step picks lies on the line after the policy slice it has just
//                   streamed through (two separate arrays cost one more cold line per level)
//   nodes           one 32-byte entry per node that has been visited at least once: the links (parent, child slots,
//                   visited block) -- no statistics
//   visited blocks  per node, the statistics of its VISITED children: one 32-byte row (visit counters, value sums, node
//                   index) per visited child in first-visit order, preceded by two u16 arrays (the child position of
//                   every row; the rows sorted by position), in one arena that grows by doubling.  A node's own
//                   statistics therefore live in its parent's block (the root's in `root_stat`); rows never move within
//                   a block, and when a block is moved to grow, its children's `row` links are rewritten.
// A selection step reads the children's policy slice and the parent's block -- a few CONTIGUOUS cache lines -- instead
// of seven statistics slices over all children or one scattered line per visited child; with dozens of trees per
// generator thread and thousands per host every tree visit starts cold, so lines touched are what the search costs.
// Unvisited children all share q = fpu, so their uct is one multiply-add chain over the policy slice; visited children
// go through the same formula as the reference's Node::uct, 8 rows at a time.  The IEEE operations and their order are
// the reference's, so trees stay bit-identical to the oracle's.
struct UctContext {  // node.rs:55-64
    uint64_t total_visits;
    ValuesAbs values;
    float visited_policy_mass;
};

struct alignas(32) ChildStat {
    uint32_t complete = 0, virt = 0;                          // complete_visits, virtual_visits
    float value = 0, win_a = 0, draw = 0, win_b = 0, ml = 0;  // sum_values (abs)
    int32_t node = 0;                                         // index of this child in `nodes`
};
struct Node {
    int32_t parent = -1;       // index of the parent, -1 for the root
    int32_t child_start = -1;  // child slots; children == None  <=>  child_start < 0
    int32_t block = -1;        // visited block: arena offset in 32-byte units, -1 while no child has been visited
    int32_t row = -1;          // arena index of this node's own statistics row (in the parent's block), -1 for the root
    uint16_t child_count = 0;
    uint16_t vis_count = 0, vis_cap = 0;
    uint8_t has_net_values = 0;
};

namespace detail {
struct UctParent {  // the per-parent part of Node::uct, computed once per selection step
    float fpu, sqrt_visits, moves_left_m1;
};
// node.rs:163-206 + Uct::total :87-98 for one child
inline float uct_one(const ChildStat& v, float policy, const UctParent& up, const SearchSettings& s, int player) {
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
// The same formula for 8 visited children at a time: their rows (complete, virt, value, win_a, draw, win_b, ml, node)
// are contiguous in the parent's block and are transposed into columns.  Same IEEE operations in the same order as
// uct_one.  Rows and `policy` must be readable up to the next multiple of 8 (what lies there is computed and ignored).
__attribute__((target("avx2"))) inline void uct_visited_avx2(const ChildStat* rows, const float* policy, int k, const UctParent& up,
                                                             const SearchSettings& s, int player, float* out) {
    const __m256 vl = _mm256_set1_ps(s.virtual_loss), fpu = _mm256_set1_ps(up.fpu), sq = _mm256_set1_ps(up.sqrt_visits);
    const __m256 mlm1 = _mm256_set1_ps(up.moves_left_m1), ds = _mm256_set1_ps(s.q_mode.draw_score), zero = _mm256_setzero_ps();
    const UctWeights& w = s.weights;
    const __m256 ew = _mm256_set1_ps(w.exploration_weight), mw = _mm256_set1_ps(w.moves_left_weight);
    const __m256 clip = _mm256_set1_ps(w.moves_left_clip), nclip = _mm256_set1_ps(-w.moves_left_clip), sharp = _mm256_set1_ps(w.moves_left_sharpness);
    const __m256 one = _mm256_set1_ps(1.0f), none = _mm256_set1_ps(-1.0f);
    for (int g = 0; g < k; g += 8) {
        __m256 r[8];
        for (int j = 0; j < 8; j++) r[j] = _mm256_load_ps(reinterpret_cast<const float*>(rows + g + j));
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
// Game::play_interior(move), where a game offers it: play a move into a position that is KNOWN not to be terminal (the search tree
// already holds its children), skipping whatever the game does to find out -- for chess the search for a legal reply
template <typename Game>
auto play_move(Game& b, uint32_t mv, bool known_not_terminal) -> decltype(b.play_interior(mv), void()) {
    if (known_not_terminal) b.play_interior(mv);
    else b.play(mv);
}
template <typename Game, typename... Ignored>
void play_move(Game& b, uint32_t mv, bool, Ignored...) {
    b.play(mv);
}
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
    // child slots: per node `n` policy floats followed by (n + 1) / 2 float-sized words holding the n moves as uint16
    std::vector<float> child_data;
    size_t child_slots = 1;  // the root and every created child
    // visited nodes; [0] is the root
    std::vector<Node> nodes;
    ChildStat root_stat;
    // visited blocks, in 32-byte units
    std::vector<ChildStat> arena;
    size_t arena_used = 0;
    std::vector<float> uct_scratch, vis_policy, vis_out;
    std::vector<uint32_t> visits_scratch;
    static constexpr int kRoot = 0;
    static constexpr size_t kArenaTail = 8;  // rows that may be read (never used) past the last block

    explicit Tree(const Game& root) : root_board(root) {  // tree.rs:31-39
        if (root.done()) throw std::runtime_error("Cannot build tree for done board");
        nodes.emplace_back();
        arena.resize(kArenaTail);
    }
    // the same tree object for the next search: keeps the memory of every array (a new Tree per move would allocate, fault in and
    // free ~0.5 MB per move and per game)
    void reset(const Game& root) {
        if (root.done()) throw std::runtime_error("Cannot build tree for done board");
        root_board = root;
        child_data.clear(), nodes.clear();
        child_slots = 1;
        nodes.emplace_back();
        root_stat = ChildStat();
        arena.assign(kArenaTail, ChildStat());
        arena_used = 0;
    }
    size_t size() const { return child_slots; }  // nodes in the reference's sense: the root and every created child
    void reserve(size_t slots, size_t visited) {
        child_data.reserve(slots + slots / 2 + visited);
        nodes.reserve(visited + 1), arena.reserve(5 * visited + 64);
    }
    // all children of a node at once (step.rs:89-97), with a uniform prior
    int push_children(const std::vector<uint32_t>& moves, float p) {
        const size_t start = child_data.size(), n = moves.size();
        child_data.resize(start + n + (n + 1) / 2, p);
        unsigned char* m = reinterpret_cast<unsigned char*>(child_data.data() + start + n);
        for (size_t i = 0; i < n; i++) {
            if (moves[i] > 0xFFFFu) throw std::logic_error("move ids are stored in 16 bits");
            const uint16_t v = uint16_t(moves[i]);
            std::memcpy(m + 2 * i, &v, 2);
        }
        child_slots += n;
        return int(start);
    }
    // the children of the node whose slice starts at `child_start` and has `n` children
    float* policy_of(int child_start) { return child_data.data() + child_start; }
    const float* policy_of(int child_start) const { return child_data.data() + child_start; }
    uint32_t move_of(int child_start, int n, int i) const {
        uint16_t v;
        std::memcpy(&v, reinterpret_cast<const unsigned char*>(child_data.data() + child_start + n) + 2 * size_t(i), 2);
        return v;
    }

    // block layout: u16 row_pos[cap] | u16 order[cap] | ChildStat rows[cap]
    static size_t head_units(size_t cap) { return (cap + 7) / 8; }
    uint16_t* block_row_pos(const Node& n) { return reinterpret_cast<uint16_t*>(arena.data() + n.block); }
    const uint16_t* block_row_pos(const Node& n) const { return reinterpret_cast<const uint16_t*>(arena.data() + n.block); }
    uint16_t* block_order(const Node& n) { return block_row_pos(n) + n.vis_cap; }
    const uint16_t* block_order(const Node& n) const { return block_row_pos(n) + n.vis_cap; }
    ChildStat* block_rows(const Node& n) { return arena.data() + n.block + head_units(n.vis_cap); }
    const ChildStat* block_rows(const Node& n) const { return arena.data() + n.block + head_units(n.vis_cap); }

    // index, in `parent`'s block, of the row of the child at position `pos`; the row (and the child's node) is created
    // on the first visit
    int visit_child(int parent, int pos) {
        {
            const Node& p = nodes[size_t(parent)];
            if (p.block >= 0) {
                const uint16_t* rp = block_row_pos(p);
                for (int j = 0; j < p.vis_count; j++)
                    if (rp[j] == pos) return j;
            }
        }
        const int v = int(nodes.size());
        nodes.emplace_back();
        nodes.back().parent = parent;
        Node& p = nodes[size_t(parent)];
        const int k = p.vis_count;
        if (k == p.vis_cap) {  // move the block to the end of the arena with twice the room
            const size_t cap = p.vis_cap ? 2 * size_t(p.vis_cap) : 4, off = arena_used;
            arena_used += head_units(cap) + cap;
            if (arena.size() < arena_used + kArenaTail) arena.resize(std::max(arena_used + kArenaTail, 2 * arena.size()));
            uint16_t* head = reinterpret_cast<uint16_t*>(arena.data() + off);
            ChildStat* rows = arena.data() + off + head_units(cap);
            if (k) {
                std::memcpy(head, block_row_pos(p), size_t(k) * sizeof(uint16_t));
                std::memcpy(head + cap, block_order(p), size_t(k) * sizeof(uint16_t));
                std::memcpy(rows, block_rows(p), size_t(k) * sizeof(ChildStat));
                for (int j = 0; j < k; j++) nodes[size_t(rows[j].node)].row = int32_t(rows + j - arena.data());
            }
            p.block = int32_t(off);
            p.vis_cap = uint16_t(cap);
        }
        uint16_t* rp = block_row_pos(p);
        uint16_t* order = block_order(p);
        ChildStat* rows = block_rows(p);
        rp[k] = uint16_t(pos);
        int t = k;
        for (; t > 0 && rp[order[t - 1]] > pos; t--) order[t] = order[t - 1];
        order[t] = uint16_t(k);
        rows[k] = ChildStat();
        rows[k].node = v;
        nodes[size_t(v)].row = int32_t(rows + k - arena.data());
        p.vis_count++;
        return k;
    }
    // the statistics of a visited node: a row of its parent's block
    ChildStat& stat_of(int node) {
        const Node& n = nodes[size_t(node)];
        return n.row < 0 ? root_stat : arena[size_t(n.row)];
    }
    // complete visits of every child of a node, in child order
    void child_visits(int node, std::vector<uint32_t>& out) const {
        const Node& p = nodes[size_t(node)];
        out.assign(size_t(p.child_count), 0u);
        if (p.block < 0) return;
        const uint16_t* rp = block_row_pos(p);
        const ChildStat* rows = block_rows(p);
        for (int j = 0; j < p.vis_count; j++) out[size_t(rp[j])] = rows[j].complete;
    }
    const Node& root() const { return nodes[kRoot]; }
    uint64_t root_visits() const { return root_stat.complete; }
    static ValuesAbs sum_values(const ChildStat& v) { return {v.value, v.win_a, v.draw, v.win_b, v.ml}; }
    static ValuesAbs values(const ChildStat& v) { return sum_values(v).div(float(v.complete)); }  // node.rs:126-128
    ValuesAbs root_values() const { return values(root_stat); }

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
            ChildStat& n = stat_of(cur);
            if (n.virt == 0) throw std::logic_error("propagate: node has no virtual visit");
            n.complete += 1;
            n.virt -= 1;
            n.value += v.value, n.win_a += v.win_a, n.draw += v.draw, n.win_b += v.win_b, n.ml += v.moves_left;
            cur = nodes[size_t(cur)].parent;
            if (cur < 0) break;
            v = v.parent();
        }
    }

    // tree.rs:132-141: visit distribution over the root's children
    void policy(std::vector<float>& out) const {
        const Node& r = nodes[kRoot];
        const float denom = std::fmax(float(root_stat.complete) - 1.0f, 0.0f);
        out.assign(size_t(r.child_count), 0.0f / denom);
        if (r.block < 0) return;
        const uint16_t* rp = block_row_pos(r);
        const ChildStat* rows = block_rows(r);
        for (int j = 0; j < r.vis_count; j++) out[size_t(rp[j])] = float(rows[j].complete) / denom;
    }
};

template <typename Game>
struct Request {
    int node = -1;  // index (in Tree::nodes) of the node to evaluate
    int child_start = 0, child_count = 0;  // its freshly created child slots
    Game board;
    bool is_root() const { return node == Tree<Game>::kRoot; }
};

// One in-flight zero_step_gather (step.rs:61-135), advanced one tree level per descent_step call.
template <typename Game>
struct Descent {
    int cur = Tree<Game>::kRoot;
    int cur_row = -1;  // arena index of cur's statistics row, -1 for the root
    Game board;
    void begin(Tree<Game>& tree) {
        cur = Tree<Game>::kRoot;
        cur_row = -1;
        board = tree.root_board;
        tree.root_stat.virt += 1;
    }
};
enum class StepResult { kDescend, kRequest, kTerminal };

inline void prefetch_span(const void* p, size_t bytes) {
    const char* c = static_cast<const char*>(p);
    for (size_t o = 0; o < bytes + 63; o += 64) __builtin_prefetch(c + o);
}

struct NoLeafHook {
    template <typename Game>
    void operator()(const Game&) const {}
};
// kRequest: an un-evaluated node was reached and `req` is filled; kTerminal: a terminal node was reached and its outcome
// has been propagated; kDescend: moved one level down, call again.  The virtual visit of the node a step arrives at is
// added when the node is selected (Descent::begin for the root).
// `on_leaf(board)` runs when an un-evaluated node is reached, before its children are created: the caller's chance to
// start fetching whatever it will look up for this board (the evaluation cache) while the expansion still has work to do.
template <typename Game, typename LeafHook = NoLeafHook>
StepResult descent_step(Tree<Game>& tree, const SearchSettings& s, Rng& rng, Descent<Game>& d, Request<Game>& req, std::vector<uint32_t>& scratch,
                        LeafHook on_leaf = LeafHook()) {
    const int cur = d.cur;
    Game& board = d.board;
    if (board.done()) {
        tree.propagate(cur, ValuesAbs::from_outcome(board.outcome(), 0.0f));
        return StepResult::kTerminal;
    }
    if (tree.nodes[size_t(cur)].child_start < 0) {
        // initialise the children with a uniform policy, step.rs:84-103
        on_leaf(board);
        board.moves(scratch);
        const float p = 1.0f / float(scratch.size());
        const int start = tree.push_children(scratch, p);
        Node& n = tree.nodes[size_t(cur)];
        n.child_start = start;
        n.child_count = uint16_t(scratch.size());
        n.has_net_values = 0;
        req.node = cur;
        req.child_start = start;
        req.child_count = n.child_count;
        req.board = board;
        return StepResult::kRequest;
    }
    const Node& pn = tree.nodes[size_t(cur)];
    const ChildStat own = d.cur_row < 0 ? tree.root_stat : tree.arena[size_t(d.cur_row)];
    const int c0 = pn.child_start, n = pn.child_count, k = pn.vis_count;
    const uint16_t* vis_pos = k ? tree.block_row_pos(pn) : nullptr;  // child position of every row
    const uint16_t* order = k ? tree.block_order(pn) : nullptr;      // rows sorted by child position
    const ChildStat* rows = k ? tree.block_rows(pn) : nullptr;
    const int player = board.next_player();
    int arg = -1;
    if (own.complete == 0) {
        // a random least-visited child, step.rs:112-114 (choose_max_by_key over Reverse(total_visits))
        tree.visits_scratch.assign(size_t(n), 0u);
        for (int j = 0; j < k; j++) tree.visits_scratch[size_t(vis_pos[j])] = rows[j].complete + rows[j].virt;
        uint64_t best = 0;
        uint32_t ties = 0;
        for (int i = 0; i < n; i++) {
            const uint64_t v = tree.visits_scratch[size_t(i)];
            if (arg < 0 || v < best) {
                arg = i;
                best = v;
                ties = 1;
            } else if (v == best) {
                ties++;
                if (rng.gen_range(ties) == 0) arg = i;
            }
        }
    } else {
        const FpuMode fpu = cur == Tree<Game>::kRoot ? s.fpu_root : s.fpu_child;
        const float* policy = tree.policy_of(c0);
        if (tree.uct_scratch.size() < size_t(n) + 8) tree.uct_scratch.resize(size_t(n) + 8), tree.vis_policy.resize(size_t(n) + 8), tree.vis_out.resize(size_t(n) + 8);
        // the policy mass of the visited children is summed in child order, like uct_context (tree.rs:49-66)
        for (int j = 0; j < k; j++) tree.vis_policy[size_t(j)] = policy[vis_pos[j]];
        float mass = 0.0f;
        for (int t = 0; t < k; t++) {
            const int j = order[t];
            if (rows[j].complete + rows[j].virt > 0) mass += tree.vis_policy[size_t(j)];
        }
        const UctContext ctx{uint64_t(own.complete) + own.virt, Tree<Game>::values(own), mass};
        if (ctx.total_visits == 0) throw std::runtime_error("uct is NaN");  // node.rs:171-173
        const detail::UctParent up = tree.uct_parent(ctx, fpu, s, player);
        float* u = tree.uct_scratch.data();
        detail::uct_unvisited(policy, n, up, s, u);
#if defined(__x86_64__)
        static const bool have_avx2 = __builtin_cpu_supports("avx2");
        if (have_avx2 && k > 2) {
            detail::uct_visited_avx2(rows, tree.vis_policy.data(), k, up, s, player, tree.vis_out.data());
            for (int j = 0; j < k; j++) u[vis_pos[j]] = tree.vis_out[size_t(j)];
        } else
#endif
            for (int j = 0; j < k; j++) u[vis_pos[j]] = detail::uct_one(rows[j], tree.vis_policy[size_t(j)], up, s, player);
        // choose_max_by_key with random tie break, kz-util/src/sequence.rs:11-41: a later element replaces the running
        // best when it is greater, and with probability 1/ties when it is equal.
        arg = detail::argmax_random_ties(u, n, rng);
    }
    if (arg < 0) throw std::logic_error("Board is not done, this node should have a child");
    const int j = tree.visit_child(cur, arg);  // may move cur's block: take the row from the tree again
    const Node& now = tree.nodes[size_t(cur)];
    ChildStat* row = tree.block_rows(now) + j;
    row->virt += 1;
    d.cur = row->node;
    d.cur_row = int(row - tree.arena.data());
    // a child that has children of its own was found not to be terminal when it was first reached: the game may skip that test
    detail::play_move(board, tree.move_of(c0, n, arg), tree.nodes[size_t(d.cur)].child_start >= 0);
    return StepResult::kDescend;
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
    Node& n = tree.nodes[size_t(node)];
    if (n.has_net_values) throw std::logic_error("Node was already evaluated by the network");
    n.has_net_values = 1;
    if (n.child_start < 0) throw std::logic_error("Applied node should have initialized children");
    if (size_t(n.child_count) != n_policy) throw std::logic_error("Wrong children length");
    std::memcpy(tree.policy_of(n.child_start), policy, n_policy * sizeof(float));
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
