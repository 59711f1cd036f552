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
    std::vector<Visited> pool(64);
    for (int trial = 0; trial < 20000; trial++) {
        SearchSettings s;
        s.q_mode.wdl = gen.gen_range(2) != 0;
        s.q_mode.draw_score = gen.gen_range(2) ? 0.0f : 0.3f;
        s.virtual_loss = gen.gen_range(3) == 0 ? 2.5f : 1.0f;
        if (gen.gen_range(4) == 0) s.weights.moves_left_weight = 0.0f;
        for (size_t i = 1; i < pool.size(); i++) {
            Visited& v = pool[i];
            v.complete = gen.gen_range(4) == 0 ? 0 : gen.gen_range(500);
            v.virt = gen.gen_range(3) == 0 ? 0 : gen.gen_range(20);
            const float c = float(v.complete);
            v.win_a = float(gen.uniform()) * c, v.win_b = float(gen.uniform()) * (c - v.win_a), v.draw = c - v.win_a - v.win_b;
            v.value = v.win_a - v.win_b;
            v.ml = float(gen.uniform()) * 60.0f * c;
        }
        const detail::UctParent up{float(gen.uniform()) * 2.0f - 1.0f, std::sqrt(float(1 + gen.gen_range(800))), float(gen.uniform()) * 50.0f};
        const int player = int(gen.gen_range(2)), k = 1 + int(gen.gen_range(40));
        int32_t idx[48] = {};
        float policy[48] = {}, out[48], unvisited[48];
        for (int j = 0; j < k; j++) idx[j] = int32_t(gen.gen_range(64)), policy[j] = float(gen.uniform());  // index 0 = the sentinel
        detail::uct_visited_avx2(pool.data(), idx, policy, k, up, s, player, out);
        detail::uct_unvisited(policy, k, up, s, unvisited);
        for (int j = 0; j < k; j++) {
            const float want = detail::uct_one(pool[size_t(idx[j])], policy[j], up, s, player);
            if (std::memcmp(&want, &out[j], 4) != 0 && !(want == 0.0f && out[j] == 0.0f)) {
                std::printf("uct mismatch at trial %d lane %d: scalar %.9g vector %.9g\n", trial, j, want, out[j]);
                return 1;
            }
            const float zero_want = detail::uct_one(pool[0], policy[j], up, s, player);
            if (!(zero_want == unvisited[j])) {
                std::printf("unvisited uct mismatch at trial %d lane %d: scalar %.9g fast %.9g\n", trial, j, zero_want, unvisited[j]);
                return 1;
            }
        }
    }
#endif
    return 0;
}

// visited lists stay sorted by position and map every position to one pool entry
static int check_visited_lists() {
    SynthChess board = SynthChess::start(3);
    Tree<SynthChess> tree(board);
    Rng gen(5);
    std::vector<uint32_t> moves(200);
    Visited& root = tree.pool[Tree<SynthChess>::kRoot];
    root.child_start = tree.push_children(moves, 0.005f);
    root.child_count = 200;
    std::vector<int> seen(200, -1);
    for (int step = 0; step < 5000; step++) {
        const int pos = int(gen.gen_range(200));
        const int v = tree.visit_child(Tree<SynthChess>::kRoot, pos);
        if (seen[size_t(pos)] >= 0 && seen[size_t(pos)] != v) {
            std::printf("visit_child returned a second entry for position %d\n", pos);
            return 1;
        }
        seen[size_t(pos)] = v;
        const Visited& r = tree.pool[Tree<SynthChess>::kRoot];
        const VisRef* list = tree.vis_arena.data() + r.vis_off;
        for (int j = 0; j < r.vis_count; j++)
            if ((j > 0 && list[j - 1].pos >= list[j].pos) || seen[size_t(list[j].pos)] != list[j].idx || tree.pool[size_t(list[j].idx)].parent != Tree<SynthChess>::kRoot) {
                std::printf("visited list broken at step %d entry %d\n", step, j);
                return 1;
            }
    }
    return 0;
}

int main() {
    if (check_argmax() || check_uct() || check_visited_lists()) return 1;
    std::printf("ok\n");
    return 0;
}
