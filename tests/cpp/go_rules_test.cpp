// Rule and encoding checks of the 9x9 go restatement in kzero_b200/csrc/selfplay/games.hpp (Go9).
// Compiled and run by tests/test_host_units.py.  Prints "ok" or the first failed check.
#include <algorithm>
#include <cstdio>

#include "../../kzero_b200/csrc/selfplay/games.hpp"
#include "../../kzero_b200/csrc/selfplay/mcts.hpp"

using namespace kzb::selfplay;

#define CHECK(cond)                                              \
    do {                                                         \
        if (!(cond)) {                                           \
            std::printf("line %d: %s\n", __LINE__, #cond);       \
            return 1;                                            \
        }                                                        \
    } while (0)

static uint32_t at(int x, int y) { return uint32_t(1 + y * 9 + x); }
static bool has(const Go9& g, uint32_t mv) {
    std::vector<uint32_t> m;
    g.moves(m);
    return std::find(m.begin(), m.end(), mv) != m.end();
}
static int bit(const uint8_t* bits, int plane, int p) {
    const int b = plane * 81 + p;
    return (bits[b >> 3] >> (b & 7)) & 1;
}

static int check_rules() {
    Go9 g = Go9::start(1);
    g.multi_suicide = 0;
    std::vector<uint32_t> m;
    g.moves(m);
    CHECK(m.size() == 82 && m[0] == 0 && m[1] == 1 && m[81] == 81);  // pass first, then every point
    CHECK(g.next_player() == 0 && !g.done());

    // capture: black surrounds the white stone at (1,1)
    g.play(at(1, 0));  // B
    g.play(at(1, 1));  // W
    g.play(at(0, 1));  // B
    g.play(at(8, 8));  // W elsewhere
    g.play(at(2, 1));  // B
    g.play(at(8, 7));  // W elsewhere
    CHECK(g.stones[1 * 9 + 1] == 2);
    g.play(at(1, 2));  // B takes the last liberty
    CHECK(g.stones[1 * 9 + 1] == 0);  // captured
    // suicide: white may not play into (1,1) now (four black neighbours, nothing captured)
    CHECK(g.next_player() == 1 && !has(g, at(1, 1)));
    uint8_t bits[41];
    float sc[6];
    g.encode(bits, sc);
    CHECK(bit(bits, 3, 1 * 9 + 1) == 1 && bit(bits, 3, 5 * 9 + 5) == 0);   // the suicide point is the "illegal" plane
    CHECK(bit(bits, 0, 8 * 9 + 8) == 1 && bit(bits, 1, 0 * 9 + 1) == 1);   // plane 0 = mover's stones (white), plane 1 = black
    for (int p = 0; p < 81; p++) CHECK(bit(bits, 2, p) == 1);              // in-board plane
    CHECK(sc[0] == 0.0f && sc[1] == 1.0f && sc[2] == 0.0f && sc[3] == 0.0f && sc[5] == 0.0f);
    CHECK(sc[4] == -(float(g.komi_2) * 0.5f) / 15.0f);                     // komi from white's side

    // ko:   . B W .      black plays (1,1)... build the classic shape around (1,1)/(2,1)
    Go9 k = Go9::start(2);
    const int seq[][2] = {{1, 0}, {2, 0}, {0, 1}, {3, 1}, {1, 2}, {2, 2}, {2, 1}, {1, 1}};  // B W B W B W B, then W captures (2,1)
    for (auto& s : seq) k.play(at(s[0], s[1]));
    CHECK(k.stones[1 * 9 + 2] == 0 && k.stones[1 * 9 + 1] == 2);  // white took the black stone at (2,1)
    CHECK(k.next_player() == 0 && !has(k, at(2, 1)));  // black may not retake at once: that would recreate the position before white's capture
    k.encode(bits, sc);
    CHECK(bit(bits, 3, 1 * 9 + 2) == 1);                                   // and the ko point is in the "illegal" plane
    k.play(at(7, 7));                                                      // black elsewhere
    k.play(at(7, 6));                                                      // white elsewhere
    CHECK(has(k, at(2, 1)));                                               // now the ko can be retaken: the stones elsewhere make the position new
    k.play(at(2, 1));
    CHECK(k.stones[1 * 9 + 1] == 0 && !has(k, at(1, 1)));                  // and it is a ko again, the other way round
    // positional superko beyond the simple ko: after both sides pass, the retake is still forbidden (same stones as two plies ago)
    Go9 sk = Go9::start(2);
    for (auto& s : seq) sk.play(at(s[0], s[1]));
    sk.play(0);  // black passes
    CHECK(!sk.done() && sk.next_player() == 1);
    sk.play(at(7, 7));  // white elsewhere
    CHECK(has(sk, at(2, 1)));  // black may retake now: white's extra stone makes the result a new position
    Go9 sk2 = Go9::start(2);
    for (auto& s : seq) sk2.play(at(s[0], s[1]));
    sk2.play(0);  // black passes
    sk2.play(0);  // white passes: the game is over, and a finished board marks nothing as illegal (is_available_move(..).unwrap_or(true))
    sk2.encode(bits, sc);
    CHECK(sk2.done() && sc[3] == 1.0f);
    for (int p = 0; p < 81; p++) CHECK(bit(bits, 3, p) == 0);

    // multi-stone suicide: white has two stones at (0,0),(1,0) whose last liberty is ... build: black surrounds a two-point eye
    // space; white fills one point, then the second would remove both white stones
    for (int rules = 0; rules < 2; rules++) {
        Go9 ms = Go9::start(6);
        ms.multi_suicide = uint8_t(rules);
        // black: (2,0) (0,1) (1,1); white plays (0,0) in between and then wants (1,0)
        ms.play(at(2, 0));  // B
        ms.play(at(0, 0));  // W
        ms.play(at(0, 1));  // B
        ms.play(at(8, 8));  // W elsewhere
        ms.play(at(1, 1));  // B
        CHECK(ms.next_player() == 1);
        CHECK(has(ms, at(1, 0)) == (rules == 1));  // (1,0) joins (0,0) into a group without liberties and captures nothing
        ms.encode(bits, sc);
        CHECK(sc[5] == float(rules) && bit(bits, 3, 0 * 9 + 1) == (rules == 1 ? 0 : 1));
        if (rules == 1) {
            ms.play(at(1, 0));
            CHECK(ms.stones[0] == 0 && ms.stones[1] == 0 && ms.next_player() == 0);  // both white stones are gone
        }
        // a single stone may never kill itself, whatever the rules: black's eye at ... white into a one-point eye
        Go9 ss = Go9::start(6);
        ss.multi_suicide = uint8_t(rules);
        ss.play(at(1, 0));  // B
        ss.play(at(8, 8));  // W
        ss.play(at(0, 1));  // B
        CHECK(ss.next_player() == 1 && !has(ss, at(0, 0)));
    }

    // passes and scoring
    Go9 e = Go9::start(3);
    e.komi_2 = 15;
    e.play(0);
    e.encode(bits, sc);
    CHECK(!e.done() && e.passes == 1 && sc[2] == 1.0f);
    e.play(0);
    CHECK(e.done() && e.score_2() == -15 && e.outcome() == -1);  // empty board: komi decides for white
    Go9 t = Go9::start(4);
    t.komi_2 = 15;
    t.play(at(4, 4));
    t.play(0);
    t.play(0);
    CHECK(t.done() && t.score_2() == 2 * 81 - 15 && t.outcome() == 1);  // one black stone owns the whole board
    Go9 h = Go9::start(5);
    h.komi_2 = 0;
    for (int y = 0; y < 9; y++) {  // black wall on column 3, white wall on column 5: black 4 columns, white 4 columns, column 4 neutral
        h.play(at(3, y));
        h.play(at(5, y));
    }
    h.play(0);
    h.play(0);
    CHECK(h.done() && h.score_2() == 0 && h.outcome() == 0);
    return 0;
}

// random playouts: after every move no group is left without liberties, and legal moves are what moves() says
static int check_playouts() {
    int rule_sets[2] = {0, 0};
    for (uint64_t seed = 1; seed <= 40; seed++) {
        Go9 g = Go9::start(seed);
        Rng rng(seed);
        std::vector<uint32_t> m;
        std::vector<std::vector<uint8_t>> positions{std::vector<uint8_t>(81, 0)};
        rule_sets[g.multi_suicide]++;
        for (int ply = 0; ply < 300 && !g.done(); ply++) {
            g.moves(m);
            CHECK(!m.empty() && m[0] == 0 && std::is_sorted(m.begin(), m.end()) && m.back() <= 81);
            uint32_t mv = m[rng.gen_range(uint32_t(m.size()))];
            if (mv == 0 && m.size() > 20) mv = m[1 + rng.gen_range(uint32_t(m.size() - 1))];  // keep the game going while the board is open
            const uint64_t before = g.hash();
            g.play(mv);
            CHECK(g.hash() != before);
            if (mv != 0) {  // superko: the stones never repeat
                positions.push_back(std::vector<uint8_t>(g.stones, g.stones + 81));
                for (size_t i = 0; i + 1 < positions.size(); i++) CHECK(positions[i] != positions.back());
            }
            Go9::Groups gr;
            g.groups(gr);
            for (int p = 0; p < 81; p++)
                if (g.stones[p]) CHECK(gr.libs[gr.gid[p]] > 0);
        }
    }
    CHECK(rule_sets[0] >= 10 && rule_sets[1] >= 10);  // both rule sets are drawn (go_start_pos: equal probability)
    return 0;
}

int main() {
    if (check_rules() || check_playouts()) return 1;
    std::printf("ok\n");
    return 0;
}
