// Unit checks of the vectorised pieces of kzero_b200/csrc/selfplay/mcts.hpp against their scalar definitions.
// Compiled and run by tests/test_host_units.py.  Prints "ok" or the first mismatch.
#include <cstdio>
#include <cstring>

#include "../../kzero_b200/csrc/selfplay/games.hpp"
#include "../../kzero_b200/csrc/selfplay/mcts.hpp"

using namespace kzb::selfplay;

// choose_max_by_key with random tie break, written the obvious way (kz-util/src/sequence.rs:11-41)
static int naive_argmax(const float* u, int n, Rng& rng) {
    int selected = -1;
    float best = 0;
    uint32_t ties = 0;
    for (int i = 0; i < n; i++) {
        if (selected < 0 || u[i] > best) {
            selected = i;
            best = u[i];
            ties = 1;
        } else if (u[i] == best) {
            ties++;
            if (rng.gen_range(ties) == 0) selected = i;
        }
    }
    return selected;
}

static int check_argmax() {
    Rng gen(99);
    for (int trial = 0; trial < 200000; trial++) {
        const int n = 1 + int(gen.gen_range(70));
        const uint32_t distinct = 1 + gen.gen_range(trial % 3 == 0 ? 3 : 1000);  // few distinct values: many ties
        float u[80];
        for (int i = 0; i < n; i++) u[i] = float(gen.gen_range(distinct)) * 0.25f - 3.0f;
        Rng a(trial + 1), b(trial + 1);
        const int want = naive_argmax(u, n, a), got = detail::argmax_random_ties(u, n, b);
        if (want != got || a.s != b.s) {
            std::printf("argmax mismatch at trial %d (n=%d): want %d got %d, rng %s\n", trial, n, want, got, a.s == b.s ? "same" : "differs");
            return 1;
        }
    }
    float with_nan[12] = {1, 2, 3, 4, 5, 6, 7, 8, 9, NAN, 0, 0};
    Rng r(1);
    try {
        detail::argmax_random_ties(with_nan, 12, r);
        std::printf("argmax did not reject NaN\n");
        return 1;
    } catch (const std::runtime_error&) {
    }
    return 0;
}

static int check_uct() {
#if defined(__x86_64__)
    if (!__builtin_cpu_supports("avx2")) return 0;
    Rng gen(7);
    std::vector<ChildStat> rows(48);
    for (int trial = 0; trial < 20000; trial++) {
        SearchSettings s;
        s.q_mode.wdl = gen.gen_range(2) != 0;
        s.q_mode.draw_score = gen.gen_range(2) ? 0.0f : 0.3f;
        s.virtual_loss = gen.gen_range(3) == 0 ? 2.5f : 1.0f;
        if (gen.gen_range(4) == 0) s.weights.moves_left_weight = 0.0f;
        for (ChildStat& v : rows) {
            v.complete = gen.gen_range(4) == 0 ? 0 : gen.gen_range(500);
            v.virt = gen.gen_range(3) == 0 ? 0 : gen.gen_range(20);
            const float c = float(v.complete);
            v.win_a = float(gen.uniform()) * c, v.win_b = float(gen.uniform()) * (c - v.win_a), v.draw = c - v.win_a - v.win_b;
            v.value = v.win_a - v.win_b;
            v.ml = float(gen.uniform()) * 60.0f * c;
            v.node = int32_t(gen.gen_range(1000));
        }
        const detail::UctParent up{float(gen.uniform()) * 2.0f - 1.0f, std::sqrt(float(1 + gen.gen_range(800))), float(gen.uniform()) * 50.0f};
        const int player = int(gen.gen_range(2)), k = 1 + int(gen.gen_range(40));
        float policy[48] = {}, out[48], unvisited[48];
        for (int j = 0; j < k; j++) policy[j] = float(gen.uniform());
        detail::uct_visited_avx2(rows.data(), policy, k, up, s, player, out);
        detail::uct_unvisited(policy, k, up, s, unvisited);
        const ChildStat never;
        for (int j = 0; j < k; j++) {
            const float want = detail::uct_one(rows[size_t(j)], policy[j], up, s, player);
            if (std::memcmp(&want, &out[j], 4) != 0 && !(want == 0.0f && out[j] == 0.0f)) {
                std::printf("uct mismatch at trial %d lane %d: scalar %.9g vector %.9g\n", trial, j, want, out[j]);
                return 1;
            }
            const float zero_want = detail::uct_one(never, policy[j], up, s, player);
            if (!(zero_want == unvisited[j])) {
                std::printf("unvisited uct mismatch at trial %d lane %d: scalar %.9g fast %.9g\n", trial, j, zero_want, unvisited[j]);
                return 1;
            }
        }
    }
#endif
    return 0;
}

// visited blocks: one row per visited position, rows never move inside a block, `order` sorts them by position, and
// every child's `row` link follows its row when the block is moved to grow
static int check_visited_blocks() {
    SynthChess board = SynthChess::start(3);
    Tree<SynthChess> tree(board);
    Rng gen(5);
    std::vector<uint32_t> moves(200);
    tree.nodes[0].child_start = tree.push_children(moves, 0.005f);
    tree.nodes[0].child_count = 200;
    std::vector<int> node_of(200, -1);
    for (int step = 0; step < 5000; step++) {
        const int pos = int(gen.gen_range(200));
        const int j = tree.visit_child(0, pos);
        const Node& r = tree.nodes[0];
        const ChildStat* rows = tree.block_rows(r);
        const uint16_t* rp = tree.block_row_pos(r);
        const uint16_t* order = tree.block_order(r);
        if (rp[j] != pos || (node_of[size_t(pos)] >= 0 && node_of[size_t(pos)] != rows[j].node)) {
            std::printf("visit_child returned a wrong row for position %d\n", pos);
            return 1;
        }
        node_of[size_t(pos)] = rows[j].node;
        tree.stat_of(rows[j].node).complete += 1;  // through the node's own link
        for (int t = 0; t < r.vis_count; t++) {
            const int row = order[t];
            const Node& child = tree.nodes[size_t(rows[row].node)];
            if ((t > 0 && rp[order[t - 1]] >= rp[row]) || child.parent != 0 || child.row != int32_t(rows + row - tree.arena.data())) {
                std::printf("visited block broken at step %d entry %d\n", step, t);
                return 1;
            }
        }
    }
    std::vector<uint32_t> visits;
    tree.child_visits(0, visits);
    uint32_t total = 0;
    for (uint32_t v : visits) total += v;
    if (total != 5000) {
        std::printf("child_visits sums to %u, expected 5000\n", total);
        return 1;
    }
    return 0;
}

int main() {
    if (check_argmax() || check_uct() || check_visited_blocks()) return 1;
    std::printf("ok\n");
    return 0;
}
