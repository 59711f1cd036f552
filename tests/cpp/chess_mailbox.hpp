// TEST INFRASTRUCTURE (not part of the product): round 1's mailbox chess -- square-by-square move generation, pins found by walking
// the rays, make-and-test legality where it matters.  It was the self-play driver's chess until the bitboard generator replaced it
// (kzero_b200/csrc/selfplay/chess_game.hpp); it stays here as the second implementation the bitboard one is compared with on random
// playouts (tests/cpp/chess_perft_test.cpp: legal move sets, keys, repetition counts, terminal flags, encodings).
// Moves are emitted square by square from a1; tests sort them into the product's canonical order.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../kzero_b200/csrc/selfplay/games.hpp"

namespace kzb {
namespace selfplay {

namespace mailbox_detail {
enum : int8_t { kPawn = 1, kKnight, kBishop, kRook, kQueen, kKing };

struct FlatMoves {
    // POV move of every policy index: from, to, promotion piece (0 = none)
    uint8_t from[1880], to[1880], promo[1880];
    int16_t index[64][64][5];  // [from][to][promo slot: 0 none, 1 Q, 2 R, 3 B, 4 N] -> policy index or -1
    FlatMoves() {
        std::memset(index, -1, sizeof(index));
        int n = 0;
        auto add = [&](int f, int t, int promo_piece, int slot) {
            from[n] = uint8_t(f), to[n] = uint8_t(t), promo[n] = uint8_t(promo_piece);
            index[f][t][slot] = int16_t(n++);
        };
        for (int f = 0; f < 64; f++)  // queen-like moves
            for (int t = 0; t < 64; t++) {
                const int df = f % 8 - t % 8, dr = f / 8 - t / 8;
                if (((df == 0) != (dr == 0)) || (df != 0 && std::abs(df) == std::abs(dr))) add(f, t, 0, 0);
            }
        for (int f = 0; f < 64; f++)  // knight moves
            for (int t = 0; t < 64; t++) {
                const int df = std::abs(f % 8 - t % 8), dr = std::abs(f / 8 - t / 8);
                if ((df == 1 && dr == 2) || (df == 2 && dr == 1)) add(f, t, 0, 0);
            }
        const int8_t pieces[4] = {kQueen, kRook, kBishop, kKnight};
        for (int p = 0; p < 4; p++)  // promotions, rank 7 -> rank 8
            for (int ff = 0; ff < 8; ff++)
                for (int tf = 0; tf < 8; tf++)
                    if (std::abs(ff - tf) <= 1) add(6 * 8 + ff, 7 * 8 + tf, pieces[p], p + 1);
    }
    static int slot_of(int promo_piece) { return promo_piece == 0 ? 0 : promo_piece == kQueen ? 1 : promo_piece == kRook ? 2 : promo_piece == kBishop ? 3 : 4; }
};
inline const FlatMoves& flat_moves() {
    static const FlatMoves t;
    return t;
}
struct ZobristTable {
    uint64_t v[13][64];  // [piece code + 6][square]
    ZobristTable() {
        for (int p = 0; p < 13; p++)
            for (int s = 0; s < 64; s++) v[p][s] = splitmix64(uint64_t(p + 10) * 64 + uint64_t(s) + 0xC0FFEEull);
    }
};
inline uint64_t zobrist(int piece_code, int square) {
    static const ZobristTable t;
    return t.v[piece_code + 6][square];
}
}  // namespace mailbox_detail

struct ChessMailbox {
    // square = rank * 8 + file, a1 = 0.  Piece code: +type white, -type black (type 1..6 = P N B R Q K), 0 empty.
    int8_t sq[64] = {};
    uint8_t side = 0;        // 0 white to move, 1 black
    uint8_t castle = 0;      // bit 0 white king side, 1 white queen side, 2 black king side, 3 black queen side
    int8_t ep = -1;          // en-passant capture target square, -1 none
    uint8_t halfmove = 0;    // plies without a pawn move or capture
    uint8_t terminal = 0;    // 0 running, 1 side to move is mated, 2 draw
    uint16_t ply = 0;
    uint8_t reps = 0;        // earlier occurrences of this position since the last irreversible move
    uint8_t king[2] = {4, 60};
    uint8_t low_material = 0;  // insufficient mating material (recomputed when material changes)
    uint8_t hist_n = 0;
    uint64_t key = 0;        // position_key() of the current position
    uint64_t piece_key = 0;  // the placement part of it, kept incrementally
    uint64_t hist[100];      // position keys since the last irreversible move (not including the current one)

    ChessMailbox() = default;
    ChessMailbox(const ChessMailbox& o) { *this = o; }
    ChessMailbox& operator=(const ChessMailbox& o) {  // copies only the used part of the history
        std::memcpy(static_cast<void*>(this), &o, offsetof(ChessMailbox, hist) + size_t(o.hist_n) * sizeof(uint64_t));
        return *this;
    }

    struct Mv {
        uint8_t from, to;
        int8_t promo;  // 0 or piece type
    };

    static GameShape shape() { return {13, 8, 8, 1880}; }
    static const char* name() { return "chess"; }
    static ChessMailbox start(uint64_t /*seed*/) { return from_fen("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"); }

    static ChessMailbox from_fen(const std::string& fen) {
        using namespace mailbox_detail;
        ChessMailbox b;
        size_t i = 0;
        int r = 7, f = 0;
        for (; i < fen.size() && fen[i] != ' '; i++) {
            const char c = fen[i];
            if (c == '/') {
                r--, f = 0;
            } else if (c >= '1' && c <= '8') {
                f += c - '0';
            } else {
                const char* names = "pnbrqk";
                const char lower = char(c | 0x20);
                int type = 0;
                for (int k = 0; k < 6; k++)
                    if (names[k] == lower) type = k + 1;
                b.sq[r * 8 + f] = int8_t(c == lower ? -type : type);
                f++;
            }
        }
        auto next = [&]() {
            while (i < fen.size() && fen[i] == ' ') i++;
            const size_t s = i;
            while (i < fen.size() && fen[i] != ' ') i++;
            return fen.substr(s, i - s);
        };
        b.side = next() == "b" ? 1 : 0;
        const std::string rights = next();
        for (char c : rights) {
            if (c == 'K') b.castle |= 1;
            if (c == 'Q') b.castle |= 2;
            if (c == 'k') b.castle |= 4;
            if (c == 'q') b.castle |= 8;
        }
        const std::string eps = next();
        if (eps.size() == 2 && eps[0] >= 'a' && eps[0] <= 'h') b.ep = int8_t((eps[1] - '1') * 8 + (eps[0] - 'a'));
        const std::string hm = next();
        if (!hm.empty()) b.halfmove = uint8_t(std::atoi(hm.c_str()));
        for (int s = 0; s < 64; s++) {
            if (b.sq[s] == kKing) b.king[0] = uint8_t(s);
            if (b.sq[s] == -kKing) b.king[1] = uint8_t(s);
        }
        b.low_material = b.insufficient_material();
        b.piece_key = b.placement_key();
        b.key = b.piece_key ^ b.state_key();
        b.update_terminal();
        return b;
    }

    int next_player() const { return side; }
    bool done() const { return terminal != 0; }
    int outcome() const { return terminal == 1 ? (side == 0 ? -1 : 1) : 0; }  // the mated side is the one to move

    uint64_t placement_key() const {
        uint64_t h = 0;
        for (int s = 0; s < 64; s++)
            if (sq[s]) h ^= mailbox_detail::zobrist(sq[s], s);
        return h;
    }
    uint64_t state_key() const {  // side, castling rights, en-passant square
        uint64_t h = side ? 0x9E3779B97F4A7C15ull : 0;
        h ^= splitmix64(0xCA57ull + castle);
        if (ep >= 0) h ^= splitmix64(0xE9ull + uint64_t(ep));
        return h;
    }
    uint64_t position_key() const { return placement_key() ^ state_key(); }  // what repetition compares (from scratch)
    uint64_t hash() const { return splitmix64(key ^ (uint64_t(halfmove) << 8) ^ (uint64_t(reps) << 20)); }  // + what the net sees

    static int colour_of(int8_t p) { return p > 0 ? 0 : 1; }
    int king_square(int colour) const { return king[colour]; }
    // is square s attacked by colour `by`; the square `transparent` counts as empty (a king that steps away does not
    // shelter the squares behind it)
    bool attacked(int s, int by, int transparent = -1) const {
        using namespace mailbox_detail;
        const int r = s / 8, f = s % 8, sign = by == 0 ? 1 : -1;
        const int pr = r - sign;  // a pawn of colour `by` attacks from the rank behind (from its own side)
        if (pr >= 0 && pr < 8) {
            if (f > 0 && sq[pr * 8 + f - 1] == sign * kPawn) return true;
            if (f < 7 && sq[pr * 8 + f + 1] == sign * kPawn) return true;
        }
        static const int kn[8][2] = {{2, 1}, {1, 2}, {-1, 2}, {-2, 1}, {-2, -1}, {-1, -2}, {1, -2}, {2, -1}};
        for (auto& d : kn) {
            const int rr = r + d[0], ff = f + d[1];
            if (rr >= 0 && rr < 8 && ff >= 0 && ff < 8 && sq[rr * 8 + ff] == sign * kKnight) return true;
        }
        static const int dirs[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
        for (int d = 0; d < 8; d++) {
            int rr = r + dirs[d][0], ff = f + dirs[d][1];
            for (int dist = 1; rr >= 0 && rr < 8 && ff >= 0 && ff < 8; rr += dirs[d][0], ff += dirs[d][1], dist++) {
                const int8_t p = sq[rr * 8 + ff];
                if (!p || rr * 8 + ff == transparent) continue;
                if (p * sign > 0) {
                    const int type = p * sign;
                    if (type == kQueen || (d < 4 && type == kRook) || (d >= 4 && type == kBishop) || (dist == 1 && type == kKing)) return true;
                }
                break;
            }
        }
        return false;
    }
    bool in_check() const { return attacked(king_square(side), side ^ 1); }

    template <typename F>
    void pseudo_moves(F&& emit) const {  // emit(Mv) returns false to stop
        using namespace mailbox_detail;
        const int sign = side == 0 ? 1 : -1;
        for (int s = 0; s < 64; s++) {
            const int8_t p = sq[s];
            if (p * sign <= 0) continue;
            const int type = p * sign, r = s / 8, f = s % 8;
            if (type == kPawn) {
                const int fwd = sign, start_rank = side == 0 ? 1 : 6, last = side == 0 ? 7 : 0;
                const int r1 = r + fwd;
                if (r1 < 0 || r1 > 7) continue;
                auto pawn_to = [&](int t) {
                    if (t / 8 == last) {
                        for (int8_t pp : {kQueen, kRook, kBishop, kKnight})
                            if (!emit(Mv{uint8_t(s), uint8_t(t), pp})) return false;
                        return true;
                    }
                    return emit(Mv{uint8_t(s), uint8_t(t), 0});
                };
                if (!sq[r1 * 8 + f]) {
                    if (!pawn_to(r1 * 8 + f)) return;
                    if (r == start_rank && !sq[(r + 2 * fwd) * 8 + f] && !emit(Mv{uint8_t(s), uint8_t((r + 2 * fwd) * 8 + f), 0})) return;
                }
                for (int df : {-1, 1}) {
                    const int ff = f + df;
                    if (ff < 0 || ff > 7) continue;
                    const int t = r1 * 8 + ff;
                    if ((sq[t] * sign < 0 || t == ep) && !pawn_to(t)) return;
                }
            } else if (type == kKnight || type == kKing) {
                static const int kn[8][2] = {{2, 1}, {1, 2}, {-1, 2}, {-2, 1}, {-2, -1}, {-1, -2}, {1, -2}, {2, -1}};
                static const int kg[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
                for (int d = 0; d < 8; d++) {
                    const int rr = r + (type == kKnight ? kn[d][0] : kg[d][0]), ff = f + (type == kKnight ? kn[d][1] : kg[d][1]);
                    if (rr < 0 || rr > 7 || ff < 0 || ff > 7 || sq[rr * 8 + ff] * sign > 0) continue;
                    if (!emit(Mv{uint8_t(s), uint8_t(rr * 8 + ff), 0})) return;
                }
                if (type == kKing) {  // castling: rights, empty squares, king not in / through / into check
                    const int home = side == 0 ? 4 : 60;
                    if (s == home && !attacked(home, side ^ 1)) {
                        if ((castle & (side == 0 ? 1 : 4)) && !sq[home + 1] && !sq[home + 2] && sq[home + 3] == sign * kRook &&
                            !attacked(home + 1, side ^ 1) && !attacked(home + 2, side ^ 1) && !emit(Mv{uint8_t(s), uint8_t(home + 2), 0}))
                            return;
                        if ((castle & (side == 0 ? 2 : 8)) && !sq[home - 1] && !sq[home - 2] && !sq[home - 3] && sq[home - 4] == sign * kRook &&
                            !attacked(home - 1, side ^ 1) && !attacked(home - 2, side ^ 1) && !emit(Mv{uint8_t(s), uint8_t(home - 2), 0}))
                            return;
                    }
                }
            } else {
                static const int dirs[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
                const int d0 = type == kBishop ? 4 : 0, d1 = type == kRook ? 4 : 8;
                for (int d = d0; d < d1; d++)
                    for (int rr = r + dirs[d][0], ff = f + dirs[d][1]; rr >= 0 && rr < 8 && ff >= 0 && ff < 8; rr += dirs[d][0], ff += dirs[d][1]) {
                        const int8_t q = sq[rr * 8 + ff];
                        if (q * sign > 0) break;
                        if (!emit(Mv{uint8_t(s), uint8_t(rr * 8 + ff), 0})) return;
                        if (q) break;
                    }
            }
        }
    }
    // the placement part of a move (no counters, no history): enough to test legality
    void apply_placement(const Mv& m) {
        using namespace mailbox_detail;
        const int sign = side == 0 ? 1 : -1;
        const int8_t p = sq[m.from];
        const int type = p * sign;
        if (type == kPawn && m.to == ep && !sq[m.to]) sq[(m.from / 8) * 8 + m.to % 8] = 0;  // en passant removes the passed pawn
        sq[m.to] = m.promo ? int8_t(sign * m.promo) : p;
        sq[m.from] = 0;
        if (type == kKing) {
            king[side] = m.to;
            if (std::abs(int(m.to) - int(m.from)) == 2) {  // castling moves the rook as well
                if (m.to > m.from) sq[m.from + 1] = sq[m.from + 3], sq[m.from + 3] = 0;
                else sq[m.from - 1] = sq[m.from - 4], sq[m.from - 4] = 0;
            }
        }
    }
    bool legal(const Mv& m) const {  // the full test: make the move on a scratch board, look at the king
        ChessMailbox c;
        std::memcpy(c.sq, sq, sizeof(sq));
        c.side = side, c.ep = ep, c.king[0] = king[0], c.king[1] = king[1];
        c.apply_placement(m);
        return !c.attacked(c.king_square(side), side ^ 1);
    }
    // pin_dir[s] = index (0..7) of the ray from the own king on which the own piece on s is pinned, -1 otherwise
    void find_pins(int8_t pin_dir[64]) const {
        using namespace mailbox_detail;
        static const int dirs[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
        std::memset(pin_dir, -1, 64);
        const int sign = side == 0 ? 1 : -1, k = king[side], kr = k / 8, kf = k % 8;
        for (int d = 0; d < 8; d++) {
            int candidate = -1;
            for (int rr = kr + dirs[d][0], ff = kf + dirs[d][1]; rr >= 0 && rr < 8 && ff >= 0 && ff < 8; rr += dirs[d][0], ff += dirs[d][1]) {
                const int8_t p = sq[rr * 8 + ff];
                if (!p) continue;
                if (p * sign > 0) {
                    if (candidate >= 0) break;  // two own pieces in a row: nothing is pinned on this ray
                    candidate = rr * 8 + ff;
                } else {
                    const int type = -p * sign;
                    if (candidate >= 0 && (type == kQueen || (d < 4 && type == kRook) || (d >= 4 && type == kBishop))) pin_dir[candidate] = int8_t(d);
                    break;
                }
            }
        }
    }
    static bool on_ray(int k, int s, int d) {  // is s on ray d (one of the 8 directions) from k
        const int dr = s / 8 - k / 8, df = s % 8 - k % 8;
        switch (d) {
            case 0: return df == 0 && dr > 0;
            case 1: return df == 0 && dr < 0;
            case 2: return dr == 0 && df > 0;
            case 3: return dr == 0 && df < 0;
            case 4: return dr == df && dr > 0;
            case 5: return dr == -df && dr > 0;
            case 6: return dr == -df && dr < 0;
            default: return dr == df && dr < 0;
        }
    }
    template <typename F>
    void legal_moves(F&& emit) const {
        // not in check: a piece other than the king may move unless it is pinned, and a pinned piece may move along its
        // pin ray; a king may step onto a square the other side does not attack once the king itself is off the board;
        // en-passant captures and every other move while in check take the full test
        const bool check = in_check();
        const int k = king[side];
        int8_t pin_dir[64];
        if (!check) find_pins(pin_dir);
        pseudo_moves([&](const Mv& m) {
            bool ok;
            if (m.from == k) ok = !attacked(m.to, side ^ 1, k);  // the squares a castling king crosses were tested by the generator
            else if (check || (m.to == ep && std::abs(int(sq[m.from])) == mailbox_detail::kPawn)) ok = legal(m);
            else ok = pin_dir[m.from] < 0 || on_ray(k, m.to, pin_dir[m.from]);
            return !ok || emit(m);
        });
    }
    bool has_legal_move() const {
        bool any = false;
        legal_moves([&](const Mv&) {
            any = true;
            return false;
        });
        return any;
    }
    // board-game's Rules::is_draw ends a game on material only when nothing but the two kings is left; K + minor v K plays on --
    // the reference's own tests play knight moves on "8/8/6k1/8/3N4/6K1/8/8 w" (rust/kz-core/tests/mapper/chess/pairs.rs:98-136)
    bool insufficient_material() const {
        int pieces = 0;
        for (int s = 0; s < 64; s++) pieces += sq[s] != 0;
        return pieces <= 2;
    }
    void update_terminal() {
        if (!has_legal_move()) terminal = in_check() ? 1 : 2;
        else if (halfmove >= 100 || reps >= 2 || low_material) terminal = 2;
        else terminal = 0;
    }

    // moves are policy indices from the mover's point of view (ranks flipped for black, move_pov chess.rs:483-497)
    static int pov_square(int s, int side_) { return side_ == 0 ? s : (7 - s / 8) * 8 + s % 8; }
    uint32_t index_of(const Mv& m) const {
        const auto& t = mailbox_detail::flat_moves();
        return uint32_t(t.index[pov_square(m.from, side)][pov_square(m.to, side)][mailbox_detail::FlatMoves::slot_of(m.promo)]);
    }
    void moves(std::vector<uint32_t>& out) const {
        out.clear();
        legal_moves([&](const Mv& m) {
            out.push_back(index_of(m));
            return true;
        });
    }
    uint32_t move_to_index(uint32_t mv) const { return mv; }
    void play(uint32_t index) {
        apply_move(index);
        update_terminal();
    }
    // the same move into a position the caller knows not to be terminal (mcts.hpp: the tree already holds its children): skips the
    // search for a legal reply; everything later tests need (repetitions, material, clocks) is kept
    void play_interior(uint32_t index) {
        apply_move(index);
        terminal = 0;
    }
    void apply_move(uint32_t index) {
        using namespace mailbox_detail;
        const auto& t = flat_moves();
        const Mv m{uint8_t(pov_square(t.from[index], side)), uint8_t(pov_square(t.to[index], side)), int8_t(t.promo[index])};
        const int sign = side == 0 ? 1 : -1;
        const int type = sq[m.from] * sign;
        const bool capture = sq[m.to] != 0 || (type == kPawn && m.to == ep);
        const uint64_t key_before = key;
        {  // placement key: the mover leaves `from`, whatever stood on `to` (or the pawn passed en passant) goes, the mover or
           // its promotion arrives, a castling rook changes squares
            const int8_t p = sq[m.from];
            piece_key ^= zobrist(p, m.from) ^ zobrist(m.promo ? int8_t(sign * m.promo) : p, m.to);
            if (sq[m.to]) piece_key ^= zobrist(sq[m.to], m.to);
            else if (type == kPawn && m.to == ep) piece_key ^= zobrist(int8_t(-sign * kPawn), (m.from / 8) * 8 + m.to % 8);
            if (type == kKing && std::abs(int(m.to) - int(m.from)) == 2) {
                const int rook_from = m.to > m.from ? m.from + 3 : m.from - 4, rook_to = m.to > m.from ? m.from + 1 : m.from - 1;
                piece_key ^= zobrist(int8_t(sign * kRook), rook_from) ^ zobrist(int8_t(sign * kRook), rook_to);
            }
        }
        apply_placement(m);
        if (capture || m.promo) low_material = insufficient_material();
        // castling rights: a king or rook that moves, or a rook that is captured, loses them
        auto touch = [&](int s) {
            if (s == 4) castle &= uint8_t(~3);
            if (s == 60) castle &= uint8_t(~12);
            if (s == 7) castle &= uint8_t(~1);
            if (s == 0) castle &= uint8_t(~2);
            if (s == 63) castle &= uint8_t(~4);
            if (s == 56) castle &= uint8_t(~8);
        };
        const uint8_t castle_before = castle;
        touch(m.from), touch(m.to);
        // en passant target: only when an enemy pawn stands next to the pawn that just advanced two ranks
        ep = -1;
        if (type == kPawn && std::abs(int(m.to) - int(m.from)) == 16) {
            const int f = m.to % 8;
            if ((f > 0 && sq[m.to - 1] == -sign * kPawn) || (f < 7 && sq[m.to + 1] == -sign * kPawn)) ep = int8_t((int(m.from) + int(m.to)) / 2);
        }
        const bool irreversible = type == kPawn || capture || castle != castle_before;
        if (type == kPawn || capture) halfmove = 0;
        else halfmove++;
        if (irreversible) hist_n = 0;
        else if (hist_n < 100) hist[hist_n++] = key_before;
        side ^= 1;
        ply++;
        reps = 0;
        key = piece_key ^ state_key();
        for (int i = int(hist_n) - 2; i >= 0; i -= 2)  // same side to move: every second entry back
            if (hist[i] == key) reps++;
    }

    void encode(uint8_t* bits, float* scalars) const {  // ChessStdMapper::encode_input, chess.rs:138-170
        using namespace mailbox_detail;
        std::memset(bits, 0, 104);
        const int sign = side == 0 ? 1 : -1;
        uint64_t planes[12] = {};  // mover's P N B R Q K, then the other side's
        for (int s = 0; s < 64; s++) {
            const int p = sq[s] * sign;
            if (p) planes[(p > 0 ? 0 : 6) + std::abs(p) - 1] |= 1ull << pov_square(s, side);
        }
        std::memcpy(bits, planes, sizeof(planes));  // BitBuffer::push_block: little-endian u64 per plane
        // `inner.en_passant()` of the `chess` 3.2.0 crate is the square of the PAWN that just advanced two ranks (make_move calls
        // set_ep(dest); the capture's destination is ep_sq.uforward(side_to_move)), not the capture target this struct keeps
        // for move generation: one rank towards the mover's own side of the target
        const uint64_t epb = ep >= 0 ? 1ull << pov_square(ep + (side == 0 ? -8 : 8), side) : 0;
        std::memcpy(bits + 12 * 8, &epb, 8);
        scalars[0] = side == 0 ? 1.0f : 0.0f;
        scalars[1] = side == 1 ? 1.0f : 0.0f;
        const int own_k = side == 0 ? 1 : 4, own_q = side == 0 ? 2 : 8, opp_k = side == 0 ? 4 : 1, opp_q = side == 0 ? 8 : 2;
        scalars[2] = (castle & own_k) ? 1.0f : 0.0f;
        scalars[3] = (castle & own_q) ? 1.0f : 0.0f;
        scalars[4] = (castle & opp_k) ? 1.0f : 0.0f;
        scalars[5] = (castle & opp_q) ? 1.0f : 0.0f;
        scalars[6] = float(reps);
        scalars[7] = float(halfmove);
    }
};

}  // namespace selfplay
}  // namespace kzb
