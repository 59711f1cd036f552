// Perft of the chess move generator in kzero_b200/csrc/selfplay/chess_game.hpp, through the policy-index interface the
// search uses (moves() -> indices, play(index)): node counts of the standard test positions, and on the way that the
// indices of a position's legal moves are pairwise distinct.  Compiled and run by tests/test_host_units.py.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include <chrono>

#include "../../kzero_b200/csrc/selfplay/chess_game.hpp"
#include "../../kzero_b200/csrc/selfplay/mcts.hpp"
#include "chess_mailbox.hpp"

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

// random playouts: the cached fields stay consistent with what they cache, and the bookkeeping of play() with a brute-force history
static int check_playouts() {
    int mates = 0, draws_rep = 0, draws_50 = 0, draws_material = 0, stalemates = 0;
    for (uint64_t seed = 1; seed <= 300; seed++) {
        Chess b = Chess::start(seed);
        Rng rng(seed);
        std::vector<uint32_t> m;
        std::vector<uint64_t> keys;  // keys of all earlier positions since the last irreversible move
        for (int ply = 0; ply < 400 && !b.done(); ply++) {
            b.moves(m);
            if (m.empty()) return std::printf("no moves in a running game\n"), 1;
            const uint32_t mv = m[rng.gen_range(uint32_t(m.size()))];
            const Chess before = b;
            b.play(mv);
            const auto& t = kzb::selfplay::chess_detail::flat_moves();
            const int from = Chess::pov_square(t.from[mv], before.side), to = Chess::pov_square(t.to[mv], before.side);
            const bool pawn = std::abs(int(before.sq[from])) == 1, capture = before.sq[to] != 0 || (pawn && to == before.ep);
            if (pawn || capture || b.castle != before.castle) keys.clear();
            else keys.push_back(before.key);
            int reps = 0;
            for (uint64_t k : keys) reps += k == b.key;
            if (b.key != b.position_key() || b.reps != reps || b.halfmove != ((pawn || capture) ? 0 : before.halfmove + 1)) return std::printf("bookkeeping broken at seed %llu ply %d\n", (unsigned long long)seed, ply), 1;
            if (b.sq[b.king[0]] != 6 || b.sq[b.king[1]] != -6 || b.low_material != b.insufficient_material()) return std::printf("cached fields broken\n"), 1;
            if (b.attacked(b.king[b.side ^ 1], b.side)) return std::printf("the side that just moved left its king in check\n"), 1;
            const bool any = b.has_legal_move();
            const int want = !any ? (b.in_check() ? 1 : 2) : (b.halfmove >= 100 || b.reps >= 2 || b.low_material) ? 2 : 0;
            if (b.terminal != want) return std::printf("terminal flag %d, expected %d\n", int(b.terminal), want), 1;
            if (b.terminal == 1) mates++;
            else if (b.terminal == 2) (!any ? stalemates : b.low_material ? draws_material : b.reps >= 2 ? draws_rep : draws_50)++;
        }
    }
    // random play reaches every kind of ending except (rarely) none of some kind; require the common ones
    if (mates == 0 || draws_material + draws_rep + draws_50 + stalemates == 0) return std::printf("playouts ended in no mate or no draw at all\n"), 1;
    return 0;
}

// the bitboard generator against round 1's mailbox generator, move for move over random games: the same legal moves (the mailbox
// list sorted into the canonical order), the same keys, clocks, rights,
// repetition counts, terminal flags and encodings
static int check_against_mailbox() {
    long positions = 0;
    for (uint64_t seed = 1000; seed < 1400; seed++) {
        Chess b = Chess::start(seed);
        ChessMailbox ref = ChessMailbox::start(seed);
        Rng rng(seed * 7 + 1);
        std::vector<uint32_t> m;
        for (int ply = 0; ply < 300; ply++, positions++) {
            uint8_t bits[104], ref_bits[104];
            float sc[8], ref_sc[8];
            b.encode(bits, sc), ref.encode(ref_bits, ref_sc);
            if (b.key != ref.key || b.reps != ref.reps || b.halfmove != ref.halfmove || b.castle != ref.castle || b.ep != ref.ep || b.terminal != ref.terminal ||
                b.side != ref.side || b.hash() != ref.hash() || std::memcmp(b.sq, ref.sq, 64) != 0 || std::memcmp(bits, ref_bits, 104) != 0 ||
                std::memcmp(sc, ref_sc, sizeof(sc)) != 0)
                return std::printf("bitboard and mailbox positions differ at seed %llu ply %d\n", (unsigned long long)seed, ply), 1;
            if (b.done()) break;
            // the canonical order (chess_game.hpp): pawn moves set by set -- pushes, double pushes, captures towards the a-file, towards the
            // h-file, each by destination with promotions Q R B N, then en passant by origin -- then N B R Q K by origin and destination
            struct K {
                int type, cat, a, b, slot;
                uint32_t index;
                bool operator<(const K& o) const {
                    if (type != o.type) return type < o.type;
                    if (cat != o.cat) return cat < o.cat;
                    if (a != o.a) return a < o.a;
                    if (b != o.b) return b < o.b;
                    return slot < o.slot;
                }
            };
            std::vector<K> want;
            ref.legal_moves([&](const ChessMailbox::Mv& mv) {
                const int type = std::abs(int(ref.sq[mv.from])), slot = mailbox_detail::FlatMoves::slot_of(mv.promo);
                K k{type, 0, mv.from, mv.to, slot, ref.index_of(mv)};
                if (type == 1) {
                    const int df = mv.to % 8 - mv.from % 8, dist = std::abs(int(mv.to) - int(mv.from));
                    k.cat = df == 0 ? (dist == 8 ? 0 : 1) : (!ref.sq[mv.to] ? 4 : (df < 0 ? 2 : 3));
                    k.a = k.cat == 4 ? mv.from : mv.to;
                    k.b = 0;
                }
                want.push_back(k);
                return true;
            });
            std::sort(want.begin(), want.end());
            if (ply % 2) kzb::selfplay::chess_detail::reply_cache().n = -1;  // every other position: generated afresh, not read from the cache
            b.moves(m);
            bool same = m.size() == want.size();
            for (size_t i = 0; same && i < m.size(); i++) same = m[i] == want[i].index;
            if (!same) return std::printf("legal moves differ at seed %llu ply %d (%zu vs %zu)\n", (unsigned long long)seed, ply, m.size(), want.size()), 1;
            const uint32_t mv = m[rng.gen_range(uint32_t(m.size()))];
            b.play(mv), ref.play(mv);
        }
    }
    // play_interior keeps everything but the terminal flag
    {
        Chess a = Chess::start(0), c = Chess::start(0);
        Rng rng(5);
        std::vector<uint32_t> m;
        for (int ply = 0; ply < 60 && !a.done(); ply++) {
            a.moves(m);
            const uint32_t mv = m[rng.gen_range(uint32_t(m.size()))];
            a.play(mv), c.play_interior(mv);
            if (a.key != c.key || a.reps != c.reps || a.halfmove != c.halfmove || a.hash() != c.hash() || std::memcmp(a.sq, c.sq, 64) != 0 || a.colour[0] != c.colour[0] ||
                a.colour[1] != c.colour[1] || std::memcmp(a.kind, c.kind, sizeof(a.kind)) != 0)
                return std::printf("play_interior diverges from play at ply %d\n", ply), 1;
        }
    }
    // what the search does per leaf -- play a move into a fresh position, list its replies -- timed for both generators
    auto time_it = [&](auto start_board, const char* name) {
        using B = decltype(start_board);
        const auto t0 = std::chrono::steady_clock::now();
        long n = 0;
        uint64_t sink = 0;
        for (uint64_t seed = 1; seed <= 200; seed++) {
            B b = start_board;
            Rng rng(seed);
            std::vector<uint32_t> m;
            for (int ply = 0; ply < 200 && !b.done(); ply++, n++) {
                b.moves(m);
                b.play(m[rng.gen_range(uint32_t(m.size()))]);
                sink += m.size();
            }
        }
        const double ns = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count();
        std::fprintf(stderr, "[chess] %-9s %.0f ns per play + moves (%ld positions, %llu)\n", name, ns / double(n), n, (unsigned long long)sink);
    };
    time_it(Chess::start(0), "bitboard");
    time_it(ChessMailbox::start(0), "mailbox");
    std::fprintf(stderr, "[chess] %ld positions compared with the mailbox generator\n", positions);
    return 0;
}

int main() {
    if (check_playouts()) return 1;
    if (check_against_mailbox()) return 1;
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
    // encoding (ChessStdMapper, chess.rs:136-170): planes from the mover's side, ranks flipped for black
    {
        Chess b = Chess::from_fen("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1");
        uint8_t bits[104];
        float sc[8];
        uint64_t planes[13];
        b.encode(bits, sc);
        std::memcpy(planes, bits, 104);
        const float want_w[8] = {1, 0, 1, 1, 1, 1, 0, 0};
        if (std::memcmp(sc, want_w, sizeof(sc)) != 0 || planes[0] != 0xFF00ull || planes[6] != 0x00FF000000000000ull || planes[5] != 0x10ull ||
            planes[11] != 0x1000000000000000ull || planes[3] != 0x81ull || planes[12] != 0) {
            std::printf("encoding of the initial position is wrong\n");
            return 1;
        }
        // 1. e4 c5 2. e5 d5: white may capture en passant on d6 (43); the en-passant PLANE marks the pawn that just advanced, d5 (35),
        // which is what `inner.en_passant()` of the chess 3.2.0 crate returns (mapping/chess.rs:168); white's view, no flip
        const auto& t = kzb::selfplay::chess_detail::flat_moves();
        auto play = [&](int from, int to) {  // absolute squares; the index is looked up from the mover's side
            const int f = Chess::pov_square(from, b.side), o = Chess::pov_square(to, b.side);
            b.play(uint32_t(t.index[f][o][0]));
        };
        play(12, 28), play(50, 34), play(28, 36), play(51, 35);
        b.encode(bits, sc);
        std::memcpy(planes, bits, 104);
        if (b.ep != 43 || planes[12] != (1ull << 35) || sc[0] != 1.0f || sc[7] != 0.0f) {
            std::printf("en passant after 1. e4 c5 2. e5 d5 is wrong (ep %d)\n", int(b.ep));
            return 1;
        }
        play(6, 21);  // 3. Nf3: black to move, everything flipped; the en-passant right is gone, one quiet ply on the clock
        b.encode(bits, sc);
        std::memcpy(planes, bits, 104);
        const uint64_t black_pawns_pov = (0xFF00ull & ~((1ull << 10) | (1ull << 11))) | (1ull << 26) | (1ull << 27);  // c5, d5 seen from black
        if (sc[0] != 0.0f || sc[1] != 1.0f || sc[7] != 1.0f || planes[12] != 0 || planes[0] != black_pawns_pov) {
            std::printf("black's view after 3. Nf3 is wrong (pawns %llx)\n", (unsigned long long)planes[0]);
            return 1;
        }
        // fool's mate: 1. f3 e5 2. g4 Qh4#
        Chess m = Chess::from_fen("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1");
        b = m;
        play(13, 21), play(52, 36), play(14, 30), play(59, 31);
        if (!b.done() || b.outcome() != -1) {
            std::printf("fool's mate is not a win for black\n");
            return 1;
        }
        // threefold repetition by shuffling knights
        b = m;
        for (int rep = 0; rep < 2; rep++) play(6, 21), play(62, 45), play(21, 6), play(45, 62);
        if (!b.done() || b.outcome() != 0 || b.reps != 2) {
            std::printf("threefold repetition is not a draw (reps %d)\n", int(b.reps));
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
