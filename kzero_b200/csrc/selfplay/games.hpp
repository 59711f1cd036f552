// Games for the self-play driver.  Board state -> (bits, scalars) record and move -> policy index follow the
// reference's mappers; the game rules themselves live in the un-vendored `board-game 0.8.2` crate
// (rust/Cargo.lock:254-258) and are restated from its published behaviour, NOT pinned by reference fixtures.
//
//   SynthChess   a chess-SHAPED synthetic game ("use a synthetic 'fake legal list' game first"):
//                positions are 64-bit hashes, each has 20..45 legal moves with distinct indices into the 1880-entry
//                chess policy (rust/kz-core/src/mapping/chess.rs:185), a 13x8x8 bool + 8 scalar input record
//                (chess.rs:126-134) and a bounded length.  It exercises the same tensor shapes, batch raggedness
//                and cache behaviour as chess self-play without a chess move generator.
//   Ataxx        7x7 ataxx with the reference's AtaxxStdMapper encoding (rust/kz-core/src/mapping/ataxx.rs:60-132):
//                bools = (next player's tiles, other tiles, gaps), scalar = moves_since_last_copy / 100,
//                policy index: copy -> to, jump -> (1 + FROM_DX_DY index) * A + to, pass -> 17 * A.
//   Go9          9x9 go with the reference's GoStdMapper encoding (4 bool + 6 scalar planes) and restated rules, see below.
//   Chess        legal chess, in chess_game.hpp.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

namespace kzb {
namespace selfplay {

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct GameShape {
    int bool_channels, scalar_count, board, policy_len;
    int bits_bytes() const { return (bool_channels * board * board + 7) / 8; }
};

// ------------------------------------------------------------------------------------------------ MaxMoves
// board-game's MaxMovesBoard, which the reference wraps every self-play board in BEFORE it builds a search tree
// (rust/kz-selfplay/src/server/generator_alphazero.rs:86-87): the game is over -- a draw -- once `max_moves` moves have been played,
// and the search sees that as well: a position at the limit is a terminal draw inside the tree, not a leaf the network evaluates.
// The move count is part of the board's identity (the wrapper derives Hash / Eq over it), so it is part of the cache key here.
template <typename G>
struct MaxMoves {
    G inner;
    uint32_t moves_played = 0, max_moves = 0xFFFFFFFFu;

    static GameShape shape() { return G::shape(); }
    static const char* name() { return G::name(); }
    static MaxMoves start(uint64_t seed) {
        MaxMoves b;
        b.inner = G::start(seed);
        return b;
    }
    bool at_limit() const { return moves_played >= max_moves; }
    bool done() const { return inner.done() || at_limit(); }
    int outcome() const { return inner.done() ? inner.outcome() : 0; }
    int next_player() const { return inner.next_player(); }
    void moves(std::vector<uint32_t>& out) const { inner.moves(out); }
    uint32_t move_to_index(uint32_t mv) const { return inner.move_to_index(mv); }
    void play(uint32_t mv) {
        inner.play(mv);
        moves_played++;
    }
    void play_interior(uint32_t mv) {  // into a position known not to be terminal (neither the game's end nor the length cap)
        play_inner_interior(inner, mv, 0);
        moves_played++;
    }
    template <typename T>
    static auto play_inner_interior(T& b, uint32_t mv, int) -> decltype(b.play_interior(mv), void()) { b.play_interior(mv); }
    template <typename T>
    static void play_inner_interior(T& b, uint32_t mv, long) { b.play(mv); }
    uint64_t hash() const { return splitmix64(inner.hash() ^ (uint64_t(moves_played) * 0x9E3779B97F4A7C15ull)); }
    void encode(uint8_t* bits, float* scalars) const { inner.encode(bits, scalars); }
};

// ------------------------------------------------------------------------------------------------ SynthChess
struct SynthChess {
    uint64_t h = 0;
    uint32_t ply = 0, max_len = 80;

    static GameShape shape() { return {13, 8, 8, 1880}; }
    static const char* name() { return "chess"; }  // the shapes of python/lib/games.py:232-244, so the reference's reader accepts the records
    static SynthChess start(uint64_t seed) {
        SynthChess g;
        g.h = splitmix64(seed ^ 0xC4E55ull);
        g.max_len = 60 + uint32_t(splitmix64(seed) % 61);  // 60..120 plies
        return g;
    }
    uint64_t hash() const { return h ^ (uint64_t(ply) << 56); }
    int next_player() const { return int(ply & 1); }
    bool done() const { return ply >= max_len || (ply > 10 && (h & 127) == 0); }
    int outcome() const { return int(h % 3) - 1; }
    uint32_t move_count() const { return 20 + uint32_t((h >> 3) % 26); }
    // move j <-> policy index (b + j * a) mod 1880 with a in {3, 13, 23, 33} (units mod 1880): distinct for j < 1880
    uint32_t index_of(uint32_t j) const {
        const uint32_t a = 3 + 10 * uint32_t((h >> 8) & 3), b = uint32_t((h >> 16) % 1880);
        return (b + j * a) % 1880;
    }
    void moves(std::vector<uint32_t>& out) const {  // a move IS its policy index
        const uint32_t n = move_count(), a = 3 + 10 * uint32_t((h >> 8) & 3);
        out.resize(n);
        uint32_t idx = uint32_t((h >> 16) % 1880);  // index_of(j), stepped instead of recomputed
        for (uint32_t j = 0; j < n; j++) {
            out[j] = idx;
            idx += a;
            idx -= idx >= 1880 ? 1880 : 0;
        }
    }
    uint32_t move_to_index(uint32_t mv) const { return mv; }
    void play(uint32_t mv) {
        h = splitmix64(h ^ (uint64_t(mv) * 0x9E3779B97F4A7C15ull));
        ply++;
    }
    // 12 disjoint piece boards + an (almost always empty) en-passant board; scalars as chess.rs:140-160
    void encode(uint8_t* bits, float* scalars) const {
        std::memset(bits, 0, 104);
        uint64_t occupied = 0, s = h;
        for (int piece = 0; piece < 12; piece++) {
            const int count = (piece % 6 == 0) ? 6 : (piece % 6 == 5 ? 1 : 2);
            s = splitmix64(s);  // one hash per piece type; its 6-bit fields are the squares
            uint64_t bb = 0;
            for (int k = 0; k < count; k++) bb |= 1ull << ((s >> (6 * k)) & 63);
            bb &= ~occupied;
            occupied |= bb;
            std::memcpy(bits + piece * 8, &bb, 8);  // BitBuffer::push_block: little-endian u64, bit_buffer.rs:37-55
        }
        const int stm = next_player();
        scalars[0] = stm == 0 ? 1.0f : 0.0f;
        scalars[1] = stm == 1 ? 1.0f : 0.0f;
        for (int k = 0; k < 4; k++) scalars[2 + k] = float((h >> (24 + k)) & 1);
        scalars[6] = float((h >> 30) % 3);
        scalars[7] = float(ply % 100);
    }
};

// ------------------------------------------------------------------------------------------------ Ataxx 7x7
struct Ataxx {
    static constexpr int S = 7, A = 49;
    uint64_t tiles[2] = {0, 0};  // bit index = y * S + x (Coord8::dense_index)
    uint64_t gaps = 0;
    uint32_t ply = 0, since_copy = 0;
    bool finished = false;
    int result = 0;

    static GameShape shape() { return {3, 1, 7, 17 * 49 + 1}; }
    static const char* name() { return "ataxx-7"; }
    static constexpr uint64_t full() { return (1ull << A) - 1; }
    static Ataxx start(uint64_t /*seed*/) {
        Ataxx g;
        g.tiles[0] = (1ull << 0) | (1ull << (A - 1));              // A: corners a1 / g7
        g.tiles[1] = (1ull << (S - 1)) | (1ull << (S * (S - 1)));  // B: corners g1 / a7
        return g;
    }
    uint64_t hash() const { return splitmix64(tiles[0] * 3 + splitmix64(tiles[1] * 5 + gaps + (uint64_t(ply & 1) << 62) + (uint64_t(since_copy) << 50))); }
    int next_player() const { return int(ply & 1); }
    bool done() const { return finished; }
    int outcome() const { return result; }

    // neighbourhood masks per tile (Chebyshev distance <= 1 / <= 2), built once
    struct Masks {
        uint64_t r1[A], r2[A];
        Masks() {
            for (int i = 0; i < A; i++) {
                r1[i] = r2[i] = 0;
                const int x = i % S, y = i / S;
                for (int dy = -2; dy <= 2; dy++)
                    for (int dx = -2; dx <= 2; dx++) {
                        const int xx = x + dx, yy = y + dy;
                        if (xx < 0 || xx >= S || yy < 0 || yy >= S) continue;
                        r2[i] |= 1ull << (yy * S + xx);
                        if (dx >= -1 && dx <= 1 && dy >= -1 && dy <= 1) r1[i] |= 1ull << (yy * S + xx);
                    }
            }
        }
    };
    static const Masks& masks() {
        static const Masks m;
        return m;
    }
    static uint64_t ring(uint64_t m, int r) {  // all tiles at Chebyshev distance <= r of any tile in m
        const Masks& k = masks();
        uint64_t out = 0;
        while (m) {
            const int i = __builtin_ctzll(m);
            m &= m - 1;
            out |= r == 1 ? k.r1[i] : k.r2[i];
        }
        return out;
    }
    static int from_dxdy_index(int dx, int dy) {  // FROM_DX_DY, ataxx.rs:134-151
        static const int8_t T[16][2] = {{-2, -2}, {-1, -2}, {0, -2}, {1, -2}, {2, -2}, {-2, -1}, {2, -1}, {-2, 0},
                                        {2, 0},   {-2, 1},  {2, 1},  {-2, 2}, {-1, 2}, {0, 2},   {1, 2},  {2, 2}};
        for (int i = 0; i < 16; i++)
            if (T[i][0] == dx && T[i][1] == dy) return i;
        return -1;
    }
    // a move is its policy index (ataxx.rs:60-81): copy -> to; jump -> (1 + from_index) * A + to; pass -> 17 * A
    void moves(std::vector<uint32_t>& out) const {
        out.clear();
        const uint64_t mine = tiles[ply & 1], free_tiles = full() & ~(tiles[0] | tiles[1] | gaps);
        const uint64_t copy_targets = ring(mine, 1) & free_tiles;
        for (int to = 0; to < A; to++)
            if ((copy_targets >> to) & 1) out.push_back(uint32_t(to));
        for (int from = 0; from < A; from++) {
            if (!((mine >> from) & 1)) continue;
            const uint64_t jt = ring(1ull << from, 2) & ~ring(1ull << from, 1) & free_tiles;
            for (int to = 0; to < A; to++) {
                if (!((jt >> to) & 1)) continue;
                const int dx = from % S - to % S, dy = from / S - to / S;
                out.push_back(uint32_t((1 + from_dxdy_index(dx, dy)) * A + to));
            }
        }
        if (out.empty()) out.push_back(17 * A);  // forced pass
    }
    uint32_t move_to_index(uint32_t mv) const { return mv; }
    void play(uint32_t mv) {
        const int me = int(ply & 1), other = me ^ 1;
        if (mv == 17 * A) {
            since_copy++;
        } else {
            const int to = int(mv % A);
            if (mv >= uint32_t(A)) {  // jump: remove the origin
                static const int8_t T[16][2] = {{-2, -2}, {-1, -2}, {0, -2}, {1, -2}, {2, -2}, {-2, -1}, {2, -1}, {-2, 0},
                                                {2, 0},   {-2, 1},  {2, 1},  {-2, 2}, {-1, 2}, {0, 2},   {1, 2},  {2, 2}};
                const int fi = int(mv / A) - 1;
                const int fx = to % S + T[fi][0], fy = to / S + T[fi][1];
                tiles[me] &= ~(1ull << (fy * S + fx));
                since_copy++;
            } else {
                since_copy = 0;
            }
            tiles[me] |= 1ull << to;
            const uint64_t converted = ring(1ull << to, 1) & tiles[other];
            tiles[me] |= converted;
            tiles[other] &= ~converted;
        }
        ply++;
        // game end: a side is wiped out, the board is full, nobody can move, or 100 moves without a copy
        const uint64_t free_tiles = full() & ~(tiles[0] | tiles[1] | gaps);
        const bool any_move = (ring(tiles[0], 2) & free_tiles) || (ring(tiles[1], 2) & free_tiles);
        if (!tiles[0] || !tiles[1] || !free_tiles || !any_move || since_copy >= 100) {
            finished = true;
            const int a = __builtin_popcountll(tiles[0]), b = __builtin_popcountll(tiles[1]);
            result = since_copy >= 100 && tiles[0] && tiles[1] && free_tiles && any_move ? 0 : (a > b) - (a < b);
        }
    }
    void encode(uint8_t* bits, float* scalars) const {  // ataxx.rs:106-115
        std::memset(bits, 0, size_t((3 * A + 7) / 8));
        const uint64_t planes[3] = {tiles[ply & 1], tiles[(ply & 1) ^ 1], gaps};
        for (int p = 0; p < 3; p++)
            for (int i = 0; i < A; i++)
                if ((planes[p] >> i) & 1) {
                    const int bit = p * A + i;
                    bits[bit >> 3] |= uint8_t(1u << (bit & 7));  // LSB-first, bit_buffer.rs:27-35
                }
        scalars[0] = float(since_copy) / 100.0f;
    }
};

// ------------------------------------------------------------------------------------------------ Go 9x9
// Go with the reference's GoStdMapper encoding without the territory planes (rust/kz-core/src/mapping/go.rs:46-112 with
// territory = false, the 4 + 6 plane form python/lib/games.py:180-194 describes): bools = (next player's stones, other
// stones, in-board, empty-but-illegal), scalars = (black to move, white to move, one pass, two passes, komi from the
// mover's side / 15, multi-stone suicide allowed), policy index 0 = pass, 1 + y * S + x = place (go.rs:26-31).
// Rules, restated (the `board-game` crate is not vendored): area scoring with komi, two consecutive passes end the game, positional
// superko (a placement may not recreate the stones of any earlier position of the game; contains the simple ko rule), a single stone
// may never kill itself; per game one of the two rule sets go_start_pos draws (start_pos.rs:81-84): Rules::cgos() -- suicide is
// illegal -- or Rules::tromp_taylor() -- a move may remove its own group of two or more stones (`allow_multi_stone_suicide`, the
// sixth scalar).  Start positions follow go_start_pos (start_pos.rs:69-88): komi 7.5 most of the time.
struct Go9 {
    static constexpr int S = 9, A = 81;
    static constexpr int kHist = 512;  // earlier positions remembered for the superko rule (games are capped far below this)
    uint8_t stones[A] = {};  // 0 empty, 1 player A (black), 2 player B (white)
    uint8_t passes = 0;         // 0 normal, 1 one pass, 2 done
    uint8_t multi_suicide = 0;  // Rules::tromp_taylor(): a move may remove its own group of two or more stones; Rules::cgos(): it may not
    mutable uint8_t legal_valid = 0;
    int16_t komi_2 = 15;     // komi in half points, in white's favour
    uint16_t hist_n = 1;
    uint32_t ply = 0;
    mutable uint64_t legal[2] = {};  // the legal placements of the player to move, bit p (computed by play(), or on demand)
    uint64_t key = 0;                // zobrist key of the stones
    uint64_t hist[kHist];            // keys of every position of the game so far, the current one included

    Go9() { hist[0] = 0; }  // the empty board is the first position of the game
    Go9(const Go9& o) { *this = o; }
    Go9& operator=(const Go9& o) {  // copies only the used part of the history
        std::memcpy(static_cast<void*>(this), &o, offsetof(Go9, hist) + size_t(o.hist_n) * sizeof(uint64_t));
        return *this;
    }

    static GameShape shape() { return {4, 6, S, A + 1}; }
    static const char* name() { return "go-9"; }
    // komi and rule set drawn per game like go_start_pos (rust/kz-selfplay/src/server/start_pos.rs:69-88): komi weights 4 : 4 : 2,
    // Rules::cgos() or Rules::tromp_taylor() with equal probability
    static Go9 start(uint64_t seed) {
        Go9 g;
        const uint64_t h = splitmix64(seed ^ 0x60B0A4Dull);
        const uint32_t pick = uint32_t(h % 10);
        if (pick < 4) g.komi_2 = 15;
        else if (pick < 8) g.komi_2 = int16_t(10 + (h >> 8) % 10);
        else g.komi_2 = int16_t(int((h >> 8) % 60) - 30);
        g.multi_suicide = uint8_t((h >> 40) & 1);
        return g;
    }
    static uint64_t zobrist(int colour, int p) { return splitmix64(0x60ull * 1000 + uint64_t(colour) * 128 + uint64_t(p)); }
    int next_player() const { return int(ply & 1); }
    bool done() const { return passes >= 2; }
    // what the network's answer depends on: stones, side, passes, komi, rule set -- and the set of legal moves, which under superko
    // depends on the game's history
    uint64_t hash() const {
        ensure_legal();
        uint64_t h = key ^ (0x9E3779B97F4A7C15ull * (uint64_t(ply & 1) + 3)) ^ (passes * 0xD6E8FEB86659FD93ull) ^ (uint64_t(uint16_t(komi_2)) * 0xA0761D6478BD642Full) ^
                     (multi_suicide ? 0x5851F42D4C957F2Dull : 0);
        return splitmix64(splitmix64(h ^ legal[0]) ^ legal[1]);
    }

    // groups and their liberties for the whole board: gid[p] (0 = empty), libs[g] = liberty count of group g, zob[g] = key of its stones
    struct Groups {
        uint8_t gid[A];
        uint8_t libs[A + 1];
        uint64_t zob[A + 1];
        int count = 0;
    };
    template <typename F>
    static void for_neighbours(int p, F&& f) {
        const int x = p % S, y = p / S;
        if (x > 0) f(p - 1);
        if (x < S - 1) f(p + 1);
        if (y > 0) f(p - S);
        if (y < S - 1) f(p + S);
    }
    void groups(Groups& g) const {
        std::memset(g.gid, 0, sizeof(g.gid));
        g.count = 0;
        int stack[A];
        uint8_t seen_lib[A];  // last group that counted this empty point as a liberty
        std::memset(seen_lib, 0, sizeof(seen_lib));
        for (int p0 = 0; p0 < A; p0++) {
            if (!stones[p0] || g.gid[p0]) continue;
            const int id = ++g.count;
            int top = 0, libs = 0;
            uint64_t z = 0;
            stack[top++] = p0;
            g.gid[p0] = uint8_t(id);
            while (top) {
                const int p = stack[--top];
                z ^= zobrist(stones[p0], p);
                for_neighbours(p, [&](int q) {
                    if (!stones[q]) {
                        if (seen_lib[q] != id) {
                            seen_lib[q] = uint8_t(id);
                            libs++;
                        }
                    } else if (stones[q] == stones[p0] && !g.gid[q]) {
                        g.gid[q] = uint8_t(id);
                        stack[top++] = q;
                    }
                });
            }
            g.libs[id] = uint8_t(libs);
            g.zob[id] = z;
        }
    }
    bool seen_before(uint64_t k) const {
        bool hit = false;
        for (int i = 0; i < int(hist_n); i++) hit |= hist[i] == k;
        return hit;
    }
    // Legal placements of the player to move (board-game's GoBoard::is_available_move): the point is empty; the stone ends up with a
    // liberty -- its own, a friendly group's, or by capturing -- or, under Tromp-Taylor rules, takes at least one friendly stone with
    // it (multi-stone suicide; a single stone may never kill itself); and the stones after the move are not those of any earlier
    // position of the game (positional superko, which contains the simple ko rule)
    void compute_legal() const {
        Groups g;
        groups(g);
        const uint8_t me = uint8_t(1 + (ply & 1)), other = uint8_t(3 - me);
        legal[0] = legal[1] = 0;
        for (int p = 0; p < A; p++) {
            if (stones[p]) continue;
            bool has_empty = false, captures = false, own_safe = false, own_any = false;
            uint64_t cap_key = 0, own_key = 0;
            uint8_t seen[4];
            int n_seen = 0;
            for_neighbours(p, [&](int q) {
                if (!stones[q]) {
                    has_empty = true;
                    return;
                }
                const uint8_t id = g.gid[q];
                bool first = true;
                for (int i = 0; i < n_seen; i++) first &= seen[i] != id;
                if (first) seen[n_seen++] = id;
                if (stones[q] == other) {
                    if (g.libs[id] == 1) {
                        captures = true;
                        if (first) cap_key ^= g.zob[id];
                    }
                } else {
                    own_any = true;
                    if (g.libs[id] >= 2) own_safe = true;
                    if (first) own_key ^= g.zob[id];
                }
            });
            uint64_t after;
            if (has_empty || captures || own_safe) after = key ^ zobrist(me, p) ^ cap_key;
            else if (own_any && multi_suicide) after = key ^ own_key;  // the stone and the groups it joined leave the board
            else continue;
            if (!seen_before(after)) legal[p >> 6] |= 1ull << (p & 63);
        }
        legal_valid = 1;
    }
    void ensure_legal() const {
        if (!legal_valid) compute_legal();
    }
    bool is_legal(int p) const { return (legal[p >> 6] >> (p & 63)) & 1; }
    void moves(std::vector<uint32_t>& out) const {  // a move is its policy index; pass first (available_moves order)
        ensure_legal();
        out.clear();
        out.push_back(0);
        for (int p = 0; p < A; p++)
            if (is_legal(p)) out.push_back(uint32_t(1 + p));
    }
    uint32_t move_to_index(uint32_t mv) const { return mv; }
    void play(uint32_t mv) {
        play_interior(mv);
        if (!done()) compute_legal();  // the search asks for the replies of a new leaf right away; hash() and encode() use them as well
    }
    // the move alone: the replies are computed when somebody asks for them (mcts.hpp: interior nodes of a descent never do)
    void play_interior(uint32_t mv) {
        legal_valid = 0;
        if (mv == 0) {
            passes++;
            ply++;
            return;
        }
        const int p = int(mv) - 1;
        const uint8_t me = uint8_t(1 + (ply & 1)), other = uint8_t(3 - me);
        stones[p] = me;
        key ^= zobrist(me, p);
        Groups g;
        groups(g);
        bool dead[A + 1] = {};
        bool any_dead = false;
        for_neighbours(p, [&](int q) {
            if (stones[q] == other && g.libs[g.gid[q]] == 0) dead[g.gid[q]] = any_dead = true;
        });
        if (!any_dead && g.libs[g.gid[p]] == 0) dead[g.gid[p]] = any_dead = true;  // multi-stone suicide (only ever legal under Tromp-Taylor)
        if (any_dead)
            for (int q = 0; q < A; q++)
                if (stones[q] && dead[g.gid[q]]) {
                    key ^= zobrist(stones[q], q);
                    stones[q] = 0;
                }
        if (hist_n < kHist) hist[hist_n++] = key;
        passes = 0;
        ply++;
    }
    // area score in half points from black's side, komi included
    int score_2() const {
        int black = 0, white = 0;
        bool seen[A] = {};
        int stack[A];
        for (int p0 = 0; p0 < A; p0++) {
            if (stones[p0] == 1) black++;
            else if (stones[p0] == 2) white++;
            else if (!seen[p0]) {  // an empty region belongs to a colour when it touches only that colour
                int top = 0, size = 0, touches = 0;
                stack[top++] = p0;
                seen[p0] = true;
                while (top) {
                    const int p = stack[--top];
                    size++;
                    for_neighbours(p, [&](int q) {
                        if (stones[q]) touches |= stones[q];
                        else if (!seen[q]) {
                            seen[q] = true;
                            stack[top++] = q;
                        }
                    });
                }
                if (touches == 1) black += size;
                else if (touches == 2) white += size;
            }
        }
        return 2 * (black - white) - komi_2;
    }
    int outcome() const {
        const int s2 = score_2();
        return (s2 > 0) - (s2 < 0);
    }
    // owner[p]: 1 / 2 = the colour of the stone on p, or of the only colour an empty region touches; 0 = nobody's (`chains().territory()`
    // + `Territory::player()` of the board-game crate, as GoStdMapper reads them, go.rs:90-98)
    void territory(uint8_t owner[A]) const {
        bool seen[A] = {};
        int stack[A], region[A];
        for (int p0 = 0; p0 < A; p0++) {
            if (stones[p0]) owner[p0] = stones[p0];
            else if (!seen[p0]) {
                int top = 0, size = 0, touches = 0;
                stack[top++] = p0;
                seen[p0] = true;
                while (top) {
                    const int p = stack[--top];
                    region[size++] = p;
                    for_neighbours(p, [&](int q) {
                        if (stones[q]) touches |= stones[q];
                        else if (!seen[q]) {
                            seen[q] = true;
                            stack[top++] = q;
                        }
                    });
                }
                for (int i = 0; i < size; i++) owner[region[i]] = uint8_t(touches == 1 || touches == 2 ? touches : 0);
            }
        }
    }
    void encode(uint8_t* bits, float* scalars) const { encode_planes(bits, scalars, false); }
    void encode_planes(uint8_t* bits, float* scalars, bool with_territory) const {  // GoStdMapper::encode_input, rust/kz-core/src/mapping/go.rs:62-112
        std::memset(bits, 0, size_t(((with_territory ? 7 : 4) * A + 7) / 8));
        // a finished board has no unavailable moves: is_available_move fails on it and the mapper reads that as "available"
        // (`.unwrap_or(true)`, go.rs:84)
        const bool running = !done();
        if (running) ensure_legal();
        const uint8_t me = uint8_t(1 + (ply & 1)), other = uint8_t(3 - me);
        auto set = [&](int plane, int p) {
            const int bit = plane * A + p;
            bits[bit >> 3] |= uint8_t(1u << (bit & 7));  // LSB-first, bit_buffer.rs:27-35
        };
        for (int p = 0; p < A; p++) {
            if (stones[p] == me) set(0, p);
            if (stones[p] == other) set(1, p);
            set(2, p);
            if (running && !stones[p] && !is_legal(p)) set(3, p);
        }
        if (with_territory) {  // three more planes: owned by the mover, by nobody, by the other side (go.rs:90-98)
            uint8_t owner[A];
            territory(owner);
            for (int p = 0; p < A; p++) set(owner[p] == me ? 4 : owner[p] == 0 ? 5 : 6, p);
        }
        const float komi = float(komi_2) * 0.5f;
        scalars[0] = next_player() == 0 ? 1.0f : 0.0f;
        scalars[1] = next_player() == 1 ? 1.0f : 0.0f;
        scalars[2] = passes == 1 ? 1.0f : 0.0f;
        scalars[3] = passes >= 2 ? 1.0f : 0.0f;
        scalars[4] = (next_player() == 0 ? komi : -komi) / 15.0f;
        scalars[5] = multi_suicide ? 1.0f : 0.0f;
    }
};

// The same game with GoStdMapper::new(size, true) -- the mapper the reference's self-play SERVER constructs (rust/kz-selfplay/src/server/
// server.rs:193): three territory planes after the four basic ones, 7 + 6 input channels.  (python/lib/games.py:185-186 still declares the
// 4-plane form, so the reference's Python loader reads Go9's records, not these: SURVEY.md 8 notes the inconsistency at HEAD.)
struct Go9Territory : Go9 {
    static GameShape shape() { return {7, 6, S, A + 1}; }
    static Go9Territory start(uint64_t seed) {
        Go9Territory g;
        static_cast<Go9&>(g) = Go9::start(seed);
        return g;
    }
    void encode(uint8_t* bits, float* scalars) const { encode_planes(bits, scalars, true); }
};

}  // namespace selfplay
}  // namespace kzb
