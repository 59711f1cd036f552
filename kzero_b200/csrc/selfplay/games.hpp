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
#pragma once
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

}  // namespace selfplay
}  // namespace kzb
