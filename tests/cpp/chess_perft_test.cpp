// Perft of the chess move generator in kzero_b200/csrc/selfplay/chess_game.hpp, through the policy-index interface the
// search uses (moves() -> indices, play(index)): node counts of the standard test positions, and on the way that the
// indices of a position's legal moves are pairwise distinct.  Compiled and run by tests/test_host_units.py.
#include <algorithm>
#include <cstdio>

#include "../../kzero_b200/csrc/selfplay/chess_game.hpp"

using namespace kzb::selfplay;

static bool g_index_clash = false;

static uint64_t perft(const Chess& b, int depth) {
    std::vector<uint32_t> m;
    b.moves(m);
    std::vector<uint32_t> sorted = m;
    std::sort(sorted.begin(), sorted.end());
    if (std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end() || (!sorted.empty() && sorted.back() >= 1880)) g_index_clash = true;
    if (depth == 1) return m.size();
    uint64_t n = 0;
    for (uint32_t mv : m) {
        Chess c = b;
        c.play(mv);
        // perft counts positions, not games: draws by repetition / 50 moves / material do not stop it, mate and stalemate do
        if (c.terminal == 2 && c.has_legal_move()) c.terminal = 0;
        if (!c.done()) n += perft(c, depth - 1);
    }
    return n;
}

int main() {
    struct Case {
        const char* fen;
        int depth;
        uint64_t nodes;
    };
    const Case cases[] = {
        {"rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", 1, 20},
        {"rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", 2, 400},
        {"rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", 3, 8902},
        {"rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", 4, 197281},
        {"r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1", 1, 48},
        {"r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1", 2, 2039},
        {"r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1", 3, 97862},
        {"8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", 1, 14},
        {"8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", 2, 191},
        {"8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", 3, 2812},
        {"8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", 4, 43238},
        {"r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", 1, 6},
        {"r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", 2, 264},
        {"r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", 3, 9467},
        {"rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", 1, 44},
        {"rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", 2, 1486},
        {"rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", 3, 62379},
    };
    for (const Case& c : cases) {
        const uint64_t got = perft(Chess::from_fen(c.fen), c.depth);
        if (got != c.nodes) {
            std::printf("perft(%d) of %s: %llu, expected %llu\n", c.depth, c.fen, (unsigned long long)got, (unsigned long long)c.nodes);
            return 1;
        }
    }
    if (g_index_clash) {
        std::printf("two legal moves of one position share a policy index\n");
        return 1;
    }
    std::printf("ok\n");
    return 0;
}
