// Chess for the self-play driver: legal move generation on bitboards, the reference's ChessStdMapper encoding
// (rust/kz-core/src/mapping/chess.rs:126-170: 13 bool planes + 8 scalars from the mover's point of view) and its flat
// 1880-move policy indexing (generate_all_flat_moves_pov, chess.rs:439-481; move_pov flips ranks for black).
//
// The rules live in the un-vendored `chess` / `board-game` crates and are restated here from the rules of chess:
// castling, en passant, promotions, check / mate / stalemate, the 50-move rule (100 plies without a pawn move or capture),
// threefold repetition, bare kings (K + minor v K plays on, like board-game Rules::default: tests/mapper/chess/pairs.rs:98-136).
// Pinned by perft counts (tests/cpp/chess_perft_test.cpp: initial position, Kiwipete and three more standard positions), by the
// reference's own known answers (tests/test_chess_pairs.py <- tests/mapper/chess/pairs.rs), by a second, mailbox implementation
// (tests/cpp/chess_mailbox.hpp) on random playouts and by the oracle's third one through whole search trees (tests/test_mcts.py).
//
// Why bitboards: with four host cores per B200 the self-play loop was bound by this file (round 2, profiles/r02_selfplay.md: 2.5 us of
// generator CPU per node with the mailbox generator against 0.96 us for the synthetic game).  A position is two colour sets and six
// piece-type sets; one generation computes the enemy's attack map (king lifted off the board), the checkers and the pinned pieces
// once, and every piece's targets are then a few mask operations; sliders use hyperbola quintessence (files, diagonals) and a
// first-rank table (no BMI2 dependency, 2.5 KB of tables).  play() generates the replies once -- that is the mate / stalemate test --
// and keeps them in a per-thread one-entry cache that the search's moves() call right afterwards reads back.
//
// Canonical move order (shared with the oracle): pawn moves set by set (pushes, double pushes, captures towards the a-file, captures
// towards the h-file, each by destination; promotions Q R B N; then en passant), then knights, bishops, rooks, queens, king, per piece
// type by origin and destination.  The en-passant square exists only while an enemy pawn stands next to the pushed pawn; the en-passant PLANE
// marks the pushed pawn (`Board::en_passant()` of the `chess` 3.2.0 crate), not the capture's destination; `repetitions` counts
// earlier occurrences of the position since the last irreversible move.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "games.hpp"

namespace kzb {
namespace selfplay {

namespace chess_detail {
enum : int8_t { kPawn = 1, kKnight, kBishop, kRook, kQueen, kKing };

struct FlatMoves {
    // POV move of every policy index: from, to, promotion piece (0 = none)
    uint8_t from[1880], to[1880], promo[1880];
    int16_t index[64][64][5];  // [from][to][promo slot: 0 none, 1 Q, 2 R, 3 B, 4 N] -> policy index or -1
    int16_t plain[64 * 64];    // the promo-slot-0 plane of it on its own: 8 KB that stay in the L1 cache of a generator thread
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
        for (int f = 0; f < 64; f++)
            for (int t = 0; t < 64; t++) plain[f * 64 + t] = index[f][t][0];
    }
    static int slot_of(int promo_piece) { return promo_piece == 0 ? 0 : promo_piece == kQueen ? 1 : promo_piece == kRook ? 2 : promo_piece == kBishop ? 3 : 4; }
};
inline const FlatMoves& flat_moves() {
    static const FlatMoves t;
    return t;
}

inline uint64_t bit(int s) { return 1ull << s; }
inline int lsb(uint64_t b) { return __builtin_ctzll(b); }

// attack sets on an empty board, the masks hyperbola quintessence works on, the squares between two aligned squares and the
// line through them, zobrist keys
struct Tables {
    uint64_t knight[64], king[64], pawn_att[2][64];  // pawn_att[colour][s]: squares a pawn of `colour` on s attacks
    uint64_t file_mask[64], diag_mask[64], anti_mask[64];  // the line through s, without s
    uint8_t rank_att[8][64];                               // [file][inner six occupancy bits of the rank] -> attacked files
    uint64_t between[64][64], line[64][64];                // 0 unless the squares share a rank, file or diagonal
    uint64_t zob[13][64];                                  // [piece code + 6][square]
    uint64_t zob_castle[16], zob_ep[64];
    uint8_t castle_keep[64];                               // the rights that survive a move from / to this square
    Tables() {
        auto on = [](int r, int f) { return r >= 0 && r < 8 && f >= 0 && f < 8; };
        static const int kn[8][2] = {{2, 1}, {1, 2}, {-1, 2}, {-2, 1}, {-2, -1}, {-1, -2}, {1, -2}, {2, -1}};
        static const int kg[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
        for (int s = 0; s < 64; s++) {
            const int r = s / 8, f = s % 8;
            knight[s] = king[s] = pawn_att[0][s] = pawn_att[1][s] = file_mask[s] = diag_mask[s] = anti_mask[s] = 0;
            for (auto& d : kn)
                if (on(r + d[0], f + d[1])) knight[s] |= bit((r + d[0]) * 8 + f + d[1]);
            for (auto& d : kg)
                if (on(r + d[0], f + d[1])) king[s] |= bit((r + d[0]) * 8 + f + d[1]);
            for (int df : {-1, 1}) {
                if (on(r + 1, f + df)) pawn_att[0][s] |= bit((r + 1) * 8 + f + df);
                if (on(r - 1, f + df)) pawn_att[1][s] |= bit((r - 1) * 8 + f + df);
            }
            for (int t = 0; t < 64; t++) {
                if (t == s) continue;
                const int dr = t / 8 - r, df = t % 8 - f;
                if (df == 0) file_mask[s] |= bit(t);
                if (dr == df) diag_mask[s] |= bit(t);
                if (dr == -df) anti_mask[s] |= bit(t);
            }
        }
        for (int f = 0; f < 8; f++)
            for (int o = 0; o < 64; o++) {
                const int occ = o << 1;  // files b..g
                uint8_t a = 0;
                for (int x = f + 1; x < 8; x++) {
                    a |= uint8_t(1 << x);
                    if (occ & (1 << x)) break;
                }
                for (int x = f - 1; x >= 0; x--) {
                    a |= uint8_t(1 << x);
                    if (occ & (1 << x)) break;
                }
                rank_att[f][o] = a;
            }
        std::memset(between, 0, sizeof(between));
        std::memset(line, 0, sizeof(line));
        for (int s = 0; s < 64; s++)
            for (auto& d : kg) {
                uint64_t whole = bit(s);  // the whole line through s in direction +-d
                for (int sg : {1, -1})
                    for (int r = s / 8 + sg * d[0], f = s % 8 + sg * d[1]; on(r, f); r += sg * d[0], f += sg * d[1]) whole |= bit(r * 8 + f);
                uint64_t walked = 0;
                for (int r = s / 8 + d[0], f = s % 8 + d[1]; on(r, f); r += d[0], f += d[1]) {
                    between[s][r * 8 + f] = walked;
                    line[s][r * 8 + f] = whole;
                    walked |= bit(r * 8 + f);
                }
            }
        for (int p = 0; p < 13; p++)
            for (int s = 0; s < 64; s++) zob[p][s] = splitmix64(uint64_t(p + 10) * 64 + uint64_t(s) + 0xC0FFEEull);
        std::memset(castle_keep, 15, sizeof(castle_keep));
        castle_keep[4] = 15 & ~3, castle_keep[60] = 15 & ~12, castle_keep[7] = 15 & ~1, castle_keep[0] = 15 & ~2, castle_keep[63] = 15 & ~4, castle_keep[56] = 15 & ~8;
        for (int c = 0; c < 16; c++) zob_castle[c] = splitmix64(0xCA57ull + uint64_t(c));
        for (int s = 0; s < 64; s++) zob_ep[s] = splitmix64(0xE9ull + uint64_t(s));
    }
};
inline const Tables& tables() {
    static const Tables t;
    return t;
}
inline uint64_t zobrist(int piece_code, int square) { return tables().zob[piece_code + 6][square]; }

// hyperbola quintessence: the squares a slider on s reaches along `mask` (a file or a diagonal: one bit per rank, so a byte swap
// reverses the line) with blockers `occ`
inline uint64_t hq_line(uint64_t occ, uint64_t mask, int s) {
    const uint64_t o = occ & mask, b = bit(s);  // mask does not contain s: o - b borrows through the empty squares above s
    const uint64_t fwd = o - b, rev = __builtin_bswap64(__builtin_bswap64(o) - __builtin_bswap64(b));
    return (fwd ^ rev) & mask;
}
inline uint64_t rank_line(const Tables& t, uint64_t occ, int s) {
    const int sh = s & 56;
    return uint64_t(t.rank_att[s & 7][(occ >> (sh + 1)) & 63]) << sh;
}
inline uint64_t bishop_att(const Tables& t, uint64_t occ, int s) { return hq_line(occ, t.diag_mask[s], s) | hq_line(occ, t.anti_mask[s], s); }
inline uint64_t rook_att(const Tables& t, uint64_t occ, int s) { return hq_line(occ, t.file_mask[s], s) | rank_line(t, occ, s); }

// the replies of the position play() generated last on this thread (its mate / stalemate test): the search asks for exactly
// that list next (descent_step: done(), then moves())
struct ReplyCache {
    uint64_t key = 0, white = 0, black = 0;  // the position key and, against a key collision, both colours' occupancy
    int n = -1;
    uint16_t mv[256];
};
inline ReplyCache& reply_cache() {
    static thread_local ReplyCache c;
    return c;
}
}  // namespace chess_detail

struct Chess {
    // square = rank * 8 + file, a1 = 0.  Piece code: +type white, -type black (type 1..6 = P N B R Q K), 0 empty.
    int8_t sq[64] = {};
    uint64_t colour[2] = {};  // occupancy per colour
    uint64_t kind[7] = {};    // occupancy per piece type (1..6), both colours
    uint8_t side = 0;        // 0 white to move, 1 black
    uint8_t castle = 0;      // bit 0 white king side, 1 white queen side, 2 black king side, 3 black queen side
    int8_t ep = -1;          // en-passant capture target square, -1 none
    uint8_t halfmove = 0;    // plies without a pawn move or capture
    uint8_t terminal = 0;    // 0 running, 1 side to move is mated, 2 draw
    uint8_t reps = 0;        // earlier occurrences of this position since the last irreversible move
    uint8_t king[2] = {4, 60};
    uint8_t low_material = 0;  // insufficient mating material
    uint8_t hist_n = 0;
    uint16_t ply = 0;
    uint64_t key = 0;        // position_key() of the current position
    uint64_t piece_key = 0;  // the placement part of it, kept incrementally
    uint64_t hist[100];      // position keys since the last irreversible move (not including the current one)

    Chess() = default;
    Chess(const Chess& o) { *this = o; }
    Chess& operator=(const Chess& o) {  // copies only the used part of the history
        std::memcpy(static_cast<void*>(this), &o, offsetof(Chess, hist) + size_t(o.hist_n) * sizeof(uint64_t));
        return *this;
    }

    struct Mv {
        uint8_t from, to;
        int8_t promo;  // 0 or piece type
    };

    static GameShape shape() { return {13, 8, 8, 1880}; }
    static const char* name() { return "chess"; }
    static Chess start(uint64_t /*seed*/) { return from_fen("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"); }

    void put(int s, int8_t p) {
        sq[s] = p;
        colour[p > 0 ? 0 : 1] |= chess_detail::bit(s);
        kind[p > 0 ? p : -p] |= chess_detail::bit(s);
    }
    void lift(int s) {
        const int8_t p = sq[s];
        sq[s] = 0;
        colour[p > 0 ? 0 : 1] &= ~chess_detail::bit(s);
        kind[p > 0 ? p : -p] &= ~chess_detail::bit(s);
    }

    static Chess from_fen(const std::string& fen) {
        using namespace chess_detail;
        Chess b;
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
                b.put(r * 8 + f, int8_t(c == lower ? -type : type));
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

    uint64_t occupied() const { return colour[0] | colour[1]; }
    uint64_t placement_key() const {
        uint64_t h = 0;
        for (int s = 0; s < 64; s++)
            if (sq[s]) h ^= chess_detail::zobrist(sq[s], s);
        return h;
    }
    uint64_t state_key() const {  // side, castling rights, en-passant square
        const auto& t = chess_detail::tables();
        uint64_t h = side ? 0x9E3779B97F4A7C15ull : 0;
        h ^= t.zob_castle[castle];
        if (ep >= 0) h ^= t.zob_ep[ep];
        return h;
    }
    uint64_t position_key() const { return placement_key() ^ state_key(); }  // what repetition compares (from scratch)
    uint64_t hash() const { return splitmix64(key ^ (uint64_t(halfmove) << 8) ^ (uint64_t(reps) << 20)); }  // + what the net sees

    int king_square(int c) const { return king[c]; }
    // the pieces of colour `by` that attack square s when the board's occupancy is `occ`
    uint64_t attackers(int s, int by, uint64_t occ) const {
        using namespace chess_detail;
        const Tables& t = tables();
        const uint64_t them = colour[by];
        return them & ((t.pawn_att[by ^ 1][s] & kind[kPawn]) | (t.knight[s] & kind[kKnight]) | (t.king[s] & kind[kKing]) |
                       (bishop_att(t, occ, s) & (kind[kBishop] | kind[kQueen])) | (rook_att(t, occ, s) & (kind[kRook] | kind[kQueen])));
    }
    // is square s attacked by colour `by`; the square `transparent` counts as empty (a king that steps away does not
    // shelter the squares behind it)
    bool attacked(int s, int by, int transparent = -1) const {
        const uint64_t occ = transparent >= 0 ? occupied() & ~chess_detail::bit(transparent) : occupied();
        return attackers(s, by, occ) != 0;
    }
    bool in_check() const { return attacked(king[side], side ^ 1); }

    // every square colour `by` attacks, with `occ` as the blockers
    uint64_t attack_map(int by, uint64_t occ) const {
        using namespace chess_detail;
        const Tables& t = tables();
        const uint64_t them = colour[by], pawns = them & kind[kPawn];
        constexpr uint64_t not_a = 0xFEFEFEFEFEFEFEFEull, not_h = 0x7F7F7F7F7F7F7F7Full;
        uint64_t a = by == 0 ? (((pawns & not_a) << 7) | ((pawns & not_h) << 9)) : (((pawns & not_a) >> 9) | ((pawns & not_h) >> 7));
        for (uint64_t b = them & kind[kKnight]; b; b &= b - 1) a |= t.knight[lsb(b)];
        for (uint64_t b = them & (kind[kBishop] | kind[kQueen]); b; b &= b - 1) a |= bishop_att(t, occ, lsb(b));
        for (uint64_t b = them & (kind[kRook] | kind[kQueen]); b; b &= b - 1) a |= rook_att(t, occ, lsb(b));
        return a | t.king[king[by]];
    }

    // The legal moves in canonical order; emit(Mv) returns false to stop.  Order: pawn moves set by set -- single pushes, double pushes,
    // captures towards the a-file, captures towards the h-file, each set by destination square ascending and each promotion as Q R B N,
    // then the en-passant captures by origin -- then knights, bishops, rooks, queens and the king (castling included): per piece type
    // by origin ascending, per piece by destination ascending.  One loop per piece type and set-wise pawn moves keep the generator
    // free of per-piece dispatch (the branch mispredictions of a square-by-square walk were two thirds of its time).
    template <typename F>
    void legal_moves(F&& emit) const {
        using namespace chess_detail;
        const Tables& t = tables();
        const int us = side, them = side ^ 1, k = king[us];
        const uint64_t own = colour[us], enemy = colour[them], occ = own | enemy;
        const uint64_t danger = attack_map(them, occ & ~bit(k));  // with the king lifted: it cannot hide behind itself
        const uint64_t checkers = attackers(k, them, occ);
        // non-king moves must end on `target`: anywhere (no check), on the checker or between it and the king (one check), nowhere (two)
        uint64_t target = ~own;
        if (checkers) target = (checkers & (checkers - 1)) ? 0 : (checkers | t.between[k][lsb(checkers)]);
        // own pieces that stand alone between the king and an enemy slider that would otherwise attack it
        uint64_t pinned = 0;
        const uint64_t snipers = enemy & ((rook_att(t, 0, k) & (kind[kRook] | kind[kQueen])) | (bishop_att(t, 0, k) & (kind[kBishop] | kind[kQueen])));
        for (uint64_t b = snipers; b; b &= b - 1) {
            const uint64_t mid = t.between[k][lsb(b)] & occ;
            if (mid && !(mid & (mid - 1))) pinned |= mid & own;
        }
        if (target) {
            // ---- pawns, set-wise.  A pinned pawn moves only along its pin: it may push when it shares the king's file, capture when
            // the capture runs along the diagonal it shares with the king
            const uint64_t pawns = own & kind[kPawn], free = ~pinned;
            const uint64_t promo_rank = us == 0 ? 0xFF00000000000000ull : 0xFFull;
            constexpr uint64_t not_a = 0xFEFEFEFEFEFEFEFEull, not_h = 0x7F7F7F7F7F7F7F7Full;
            const uint64_t kfile = t.file_mask[k] | bit(k), kdiag = t.diag_mask[k], kanti = t.anti_mask[k];
            const uint64_t pushers = pawns & (free | kfile);
            uint64_t single, dbl, left, right;  // destination sets
            int d_push, d_left, d_right;        // destination - origin
            if (us == 0) {
                single = (pushers << 8) & ~occ;
                dbl = ((single & 0xFF0000ull) << 8) & ~occ;
                left = ((pawns & (free | kanti) & not_a) << 7) & enemy;   // towards the a-file: up the anti-diagonal
                right = ((pawns & (free | kdiag) & not_h) << 9) & enemy;  // towards the h-file: up the diagonal
                d_push = 8, d_left = 7, d_right = 9;
            } else {
                single = (pushers >> 8) & ~occ;
                dbl = ((single & 0xFF0000000000ull) >> 8) & ~occ;
                left = ((pawns & (free | kdiag) & not_a) >> 9) & enemy;   // towards the a-file: down the diagonal
                right = ((pawns & (free | kanti) & not_h) >> 7) & enemy;  // towards the h-file: down the anti-diagonal
                d_push = -8, d_left = -9, d_right = -7;
            }
            auto pawn_set = [&](uint64_t to, int delta) {
                for (to &= target; to; to &= to - 1) {
                    const int dst = lsb(to);
                    if (bit(dst) & promo_rank) {
                        for (int8_t pp : {kQueen, kRook, kBishop, kKnight})
                            if (!emit(Mv{uint8_t(dst - delta), uint8_t(dst), pp})) return false;
                    } else if (!emit(Mv{uint8_t(dst - delta), uint8_t(dst), 0})) {
                        return false;
                    }
                }
                return true;
            };
            if (!pawn_set(single, d_push) || !pawn_set(dbl, 2 * d_push) || !pawn_set(left, d_left) || !pawn_set(right, d_right)) return;
            if (ep >= 0) {
                // en passant: two pawns leave a rank at once -- make the capture on the occupancy and look at the king
                const int cap = ep - d_push;
                for (uint64_t b = t.pawn_att[them][ep] & pawns; b; b &= b - 1) {
                    const int s = lsb(b);
                    const uint64_t occ2 = (occ ^ bit(s) ^ bit(cap)) | bit(ep);
                    const uint64_t left_over = enemy & ~bit(cap);
                    const uint64_t att = left_over & ((t.pawn_att[us][k] & kind[kPawn]) | (t.knight[k] & kind[kKnight]) |
                                                      (bishop_att(t, occ2, k) & (kind[kBishop] | kind[kQueen])) |
                                                      (rook_att(t, occ2, k) & (kind[kRook] | kind[kQueen])));
                    if (!att && !emit(Mv{uint8_t(s), uint8_t(ep), 0})) return;
                }
            }
            // ---- knights (a pinned knight cannot move), then the sliders (a pinned slider stays on the line through the king)
            for (uint64_t b = own & kind[kKnight] & free; b; b &= b - 1) {
                const int s = lsb(b);
                for (uint64_t to = t.knight[s] & target; to; to &= to - 1)
                    if (!emit(Mv{uint8_t(s), uint8_t(lsb(to)), 0})) return;
            }
            for (uint64_t b = own & kind[kBishop]; b; b &= b - 1) {
                const int s = lsb(b);
                uint64_t to = bishop_att(t, occ, s) & target;
                if (pinned & bit(s)) to &= t.line[k][s];
                for (; to; to &= to - 1)
                    if (!emit(Mv{uint8_t(s), uint8_t(lsb(to)), 0})) return;
            }
            for (uint64_t b = own & kind[kRook]; b; b &= b - 1) {
                const int s = lsb(b);
                uint64_t to = rook_att(t, occ, s) & target;
                if (pinned & bit(s)) to &= t.line[k][s];
                for (; to; to &= to - 1)
                    if (!emit(Mv{uint8_t(s), uint8_t(lsb(to)), 0})) return;
            }
            for (uint64_t b = own & kind[kQueen]; b; b &= b - 1) {
                const int s = lsb(b);
                uint64_t to = (bishop_att(t, occ, s) | rook_att(t, occ, s)) & target;
                if (pinned & bit(s)) to &= t.line[k][s];
                for (; to; to &= to - 1)
                    if (!emit(Mv{uint8_t(s), uint8_t(lsb(to)), 0})) return;
            }
        }
        // ---- king: any square the other side does not attack; castling: rights, empty squares, not in / through / into check
        uint64_t to = t.king[k] & ~own & ~danger;
        const int home = us == 0 ? 4 : 60;
        if (k == home && !checkers && (castle & (us == 0 ? 3 : 12))) {
            const int8_t rook = int8_t(us == 0 ? kRook : -kRook);
            if ((castle & (us == 0 ? 1 : 4)) && !(occ & (bit(home + 1) | bit(home + 2))) && sq[home + 3] == rook &&
                !(danger & (bit(home + 1) | bit(home + 2))))
                to |= bit(home + 2);
            if ((castle & (us == 0 ? 2 : 8)) && !(occ & (bit(home - 1) | bit(home - 2) | bit(home - 3))) && sq[home - 4] == rook &&
                !(danger & (bit(home - 1) | bit(home - 2))))
                to |= bit(home - 2);
        }
        for (; to; to &= to - 1)
            if (!emit(Mv{uint8_t(k), uint8_t(lsb(to)), 0})) return;
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
    bool insufficient_material() const {  // each colour is down to one piece, its king (no popcount: the build targets plain x86-64)
        return !(colour[0] & (colour[0] - 1)) && !(colour[1] & (colour[1] - 1));
    }

    // the legal replies as policy indices, into the per-thread cache; returns how many
    int generate_replies() const {
        chess_detail::ReplyCache& c = chess_detail::reply_cache();
        int n = 0;
        const auto& t = chess_detail::flat_moves();
        const int flip = side == 0 ? 0 : 56;  // pov_square(s) = s ^ 56 for black
        legal_moves([&](const Mv& m) {
            c.mv[n++] = uint16_t(m.promo ? t.index[m.from ^ flip][m.to ^ flip][chess_detail::FlatMoves::slot_of(m.promo)]
                                         : t.plain[(m.from ^ flip) * 64 + (m.to ^ flip)]);
            return true;
        });
        c.key = key, c.white = colour[0], c.black = colour[1];
        c.n = n;
        return n;
    }
    void update_terminal() {
        if (generate_replies() == 0) terminal = in_check() ? 1 : 2;
        else if (halfmove >= 100 || reps >= 2 || low_material) terminal = 2;
        else terminal = 0;
    }

    // moves are policy indices from the mover's point of view (ranks flipped for black, move_pov chess.rs:483-497)
    static int pov_square(int s, int side_) { return side_ == 0 ? s : s ^ 56; }
    uint32_t index_of(const Mv& m) const {
        const auto& t = chess_detail::flat_moves();
        return uint32_t(t.index[pov_square(m.from, side)][pov_square(m.to, side)][chess_detail::FlatMoves::slot_of(m.promo)]);
    }
    void moves(std::vector<uint32_t>& out) const {
        const chess_detail::ReplyCache& c = chess_detail::reply_cache();
        // the legal moves depend on placement, side, rights and ep square: the key
        if (c.n < 0 || c.key != key || c.white != colour[0] || c.black != colour[1]) generate_replies();
        out.assign(c.mv, c.mv + c.n);
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
        using namespace chess_detail;
        const auto& ft = flat_moves();
        const Tables& t = tables();
        const int flip = side == 0 ? 0 : 56;
        const int from = ft.from[index] ^ flip, to = ft.to[index] ^ flip, promo = ft.promo[index];
        const int sign = side == 0 ? 1 : -1;
        const int8_t p = sq[from];
        const int type = p * sign;
        const uint64_t key_before = key;
        bool capture = false;
        if (sq[to]) {
            capture = true;
            piece_key ^= t.zob[sq[to] + 6][to];
            lift(to);
        } else if (type == kPawn && to == ep) {  // en passant removes the passed pawn
            capture = true;
            const int cap = (from & 56) | (to & 7);
            piece_key ^= t.zob[sq[cap] + 6][cap];
            lift(cap);
        }
        const int8_t arrives = promo ? int8_t(sign * promo) : p;
        piece_key ^= t.zob[p + 6][from] ^ t.zob[arrives + 6][to];
        lift(from);
        put(to, arrives);
        if (type == kKing) {
            king[side] = uint8_t(to);
            if (to - from == 2 || from - to == 2) {  // castling moves the rook as well
                const int rook_from = to > from ? from + 3 : from - 4, rook_to = to > from ? from + 1 : from - 1;
                const int8_t rook = sq[rook_from];
                piece_key ^= t.zob[rook + 6][rook_from] ^ t.zob[rook + 6][rook_to];
                lift(rook_from);
                put(rook_to, rook);
            }
        }
        low_material = insufficient_material();
        // castling rights: a king or rook that moves, or a rook that is captured, loses them
        const uint8_t castle_before = castle;
        castle &= t.castle_keep[from] & t.castle_keep[to];
        // en passant target: only when an enemy pawn stands next to the pawn that just advanced two ranks
        ep = -1;
        if (type == kPawn && (to - from == 16 || from - to == 16)) {
            const uint64_t beside = (((bit(to) & 0xFEFEFEFEFEFEFEFEull) >> 1) | ((bit(to) & 0x7F7F7F7F7F7F7F7Full) << 1));
            if (beside & colour[side ^ 1] & kind[kPawn]) ep = int8_t((from + to) / 2);
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
        using namespace chess_detail;
        // BitBuffer::push_block: one little-endian u64 per plane; the mover's P N B R Q K, then the other side's.  Black sees the
        // board with the ranks flipped: a byte swap of the bitboard
        uint64_t planes[13];
        for (int c = 0; c < 2; c++)
            for (int ty = 0; ty < 6; ty++) {
                const uint64_t b = kind[ty + 1] & colour[c == 0 ? side : side ^ 1];
                planes[c * 6 + ty] = side == 0 ? b : __builtin_bswap64(b);
            }
        // `inner.en_passant()` of the `chess` 3.2.0 crate is the square of the PAWN that just advanced two ranks (make_move calls
        // set_ep(dest); the capture's destination is ep_sq.uforward(side_to_move)), not the capture target this struct keeps
        // for move generation: one rank towards the mover's own side of the target
        planes[12] = ep >= 0 ? 1ull << pov_square(ep + (side == 0 ? -8 : 8), side) : 0;
        std::memcpy(bits, planes, sizeof(planes));
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
