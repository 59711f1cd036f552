"""CPU restatement of the reference's AlphaZero tree search -- ORACLE / TEST INFRASTRUCTURE ONLY.

Checks the C++ search behind `kzb_mcts_trace` / `kzb_selfplay_run` (kzero_b200/csrc/selfplay/mcts.hpp).  Restates, in
plain Python with float32 arithmetic in the same operation order:
  zero_step_gather / zero_step_apply / tree_propagate_values    rust/kz-core/src/zero/step.rs:61-188
  Node::uct, Uct::total, UctWeights::default                     rust/kz-core/src/zero/node.rs:66-98,163-206
  Tree::uct_context, Tree::policy, Tree::values                  rust/kz-core/src/zero/tree.rs:49-66,95-141
  ZeroValuesAbs::{pov, from_outcome, parent}                     rust/kz-core/src/zero/values.rs:21-68
  choose_max_by_key                                              rust/kz-util/src/sequence.rs:11-41
  build_tree's gather-batch / apply loop (no cache, no noise)    rust/kz-selfplay/src/server/generator_alphazero.rs:151-215

Parity pinning: the reference's own tests for this code only assert that a search terminates (rust/kz-core/tests/tree.rs:16-68)
-- there are no golden trees, so "parity unpinned" applies to the SEARCH oracle: it is an independent second
implementation of the cited lines, not a copy of reference outputs.  The random generator (xorshift64*), the synthetic
chess-shaped game and the stand-in networks are this repo's own and are implemented twice (here and in C++).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

F = np.float32
M64 = (1 << 64) - 1


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


class Rng:
    """xorshift64*, the twin of kzb::selfplay::Rng."""

    def __init__(self, seed: int):
        self.s = (seed * 0x9E3779B97F4A7C15 + 0xD1B54A32D192ED03) & M64
        if self.s == 0:
            self.s = 0x2545F4914F6CDD1D

    def next_u64(self) -> int:
        s = self.s
        s ^= s >> 12
        s = (s ^ (s << 25)) & M64
        s ^= s >> 27
        self.s = s
        return (s * 0x2545F4914F6CDD1D) & M64

    def gen_range(self, n: int) -> int:
        return (self.next_u64() >> 32) % n


class SynthChess:
    """Twin of kzb::selfplay::SynthChess (kzero_b200/csrc/selfplay/games.hpp)."""

    def __init__(self, h: int, ply: int, max_len: int):
        self.h, self.ply, self.max_len = h, ply, max_len

    @staticmethod
    def start(seed: int) -> "SynthChess":
        return SynthChess(splitmix64(seed ^ 0xC4E55), 0, 60 + splitmix64(seed) % 61)

    def clone(self):
        return SynthChess(self.h, self.ply, self.max_len)

    def hash(self) -> int:
        return self.h ^ ((self.ply << 56) & M64)

    def next_player(self) -> int:
        return self.ply & 1

    def done(self) -> bool:
        return self.ply >= self.max_len or (self.ply > 10 and (self.h & 127) == 0)

    def outcome(self) -> int:
        return self.h % 3 - 1

    def moves(self) -> List[int]:
        n = 20 + (self.h >> 3) % 26
        a = 3 + 10 * ((self.h >> 8) & 3)
        b = (self.h >> 16) % 1880
        return [(b + j * a) % 1880 for j in range(n)]

    def play(self, mv: int) -> None:
        self.h = splitmix64(self.h ^ ((mv * 0x9E3779B97F4A7C15) & M64))
        self.ply += 1


class Go9:
    """Twin of kzb::selfplay::Go9 (kzero_b200/csrc/selfplay/games.hpp), written independently from the rules as stated there:
    area scoring with komi, two passes end the game, positional superko (the stones after a placement may not be those of any
    earlier position of the game), a single stone may not kill itself, and per game either no suicide at all (cgos) or suicide of
    two or more stones allowed (Tromp-Taylor); move 0 = pass, 1 + y * 9 + x = place.  Legality here is "make the move on a copy
    and look": no group tables, no incremental keys."""
    S, A = 9, 81

    def __init__(self):
        self.stones = [0] * self.A  # 0 empty, 1 black (player A), 2 white
        self.ply, self.komi_2, self.passes, self.multi_suicide = 0, 15, 0, 0
        self.seen = {tuple(self.stones)}  # the stones of every position so far, the current one included

    @staticmethod
    def start(seed: int) -> "Go9":
        g = Go9()
        h = splitmix64(seed ^ 0x60B0A4D)
        pick = h % 10
        if pick < 4:
            g.komi_2 = 15
        elif pick < 8:
            g.komi_2 = 10 + (h >> 8) % 10
        else:
            g.komi_2 = (h >> 8) % 60 - 30
        g.multi_suicide = (h >> 40) & 1
        return g

    def clone(self):
        g = Go9.__new__(Go9)
        g.stones, g.ply, g.komi_2, g.passes, g.multi_suicide = list(self.stones), self.ply, self.komi_2, self.passes, self.multi_suicide
        g.seen = set(self.seen)
        return g

    def hash(self) -> int:
        key = 0
        for p, v in enumerate(self.stones):
            if v:
                key ^= splitmix64(0x60 * 1000 + v * 128 + p)
        legal = sum(1 << p for p in range(self.A) if self._legal(p))
        h = key ^ ((0x9E3779B97F4A7C15 * ((self.ply & 1) + 3)) & M64) ^ ((self.passes * 0xD6E8FEB86659FD93) & M64) \
            ^ (((self.komi_2 & 0xFFFF) * 0xA0761D6478BD642F) & M64) ^ (0x5851F42D4C957F2D if self.multi_suicide else 0)
        return splitmix64(splitmix64(h ^ (legal & M64)) ^ (legal >> 64))

    def next_player(self) -> int:
        return self.ply & 1

    def done(self) -> bool:
        return self.passes >= 2

    def _neighbours(self, p: int):
        x, y = p % self.S, p // self.S
        if x > 0:
            yield p - 1
        if x < self.S - 1:
            yield p + 1
        if y > 0:
            yield p - self.S
        if y < self.S - 1:
            yield p + self.S

    def _group(self, stones, p: int):
        """-> (stones of the group containing p, its liberties)"""
        colour, group, libs, todo = stones[p], {p}, set(), [p]
        while todo:
            q = todo.pop()
            for r in self._neighbours(q):
                if stones[r] == 0:
                    libs.add(r)
                elif stones[r] == colour and r not in group:
                    group.add(r)
                    todo.append(r)
        return group, libs

    def _placed(self, p: int):
        """The stones after the player to move places on the empty point p, or None when the rules forbid the placement itself
        (superko is the caller's business)."""
        me = 1 + (self.ply & 1)
        stones = list(self.stones)
        stones[p] = me
        for q in self._neighbours(p):
            if stones[q] not in (0, me):
                group, libs = self._group(stones, q)
                if not libs:
                    for r in group:
                        stones[r] = 0
        group, libs = self._group(stones, p)
        if not libs:  # nothing was captured (a capture would have freed a point next to p): suicide
            if len(group) < 2 or not self.multi_suicide:
                return None
            for r in group:
                stones[r] = 0
        return stones

    def _legal(self, p: int) -> bool:
        if self.stones[p]:
            return False
        after = self._placed(p)
        return after is not None and tuple(after) not in self.seen

    def moves(self) -> List[int]:
        return [0] + [1 + p for p in range(self.A) if self._legal(p)]

    def play(self, mv: int) -> None:
        if mv == 0:
            self.passes += 1
            self.ply += 1
            return
        self.stones = self._placed(mv - 1)
        self.seen.add(tuple(self.stones))
        self.passes = 0
        self.ply += 1

    TERRITORY = False  # Go9Territory: three more planes

    def _owners(self):
        """Per point: the colour of its stone, or of the only colour its empty region touches, else 0."""
        owner = list(self.stones)
        done = set()
        for p0 in range(self.A):
            if self.stones[p0] or p0 in done:
                continue
            region, touches, todo = {p0}, set(), [p0]
            while todo:
                q = todo.pop()
                for r in self._neighbours(q):
                    if self.stones[r]:
                        touches.add(self.stones[r])
                    elif r not in region:
                        region.add(r)
                        todo.append(r)
            done |= region
            for q in region:
                owner[q] = next(iter(touches)) if len(touches) == 1 else 0
        return owner

    def encode(self):
        """-> (bools [4 | 7, 9, 9] u8, scalars f32): GoStdMapper::encode_input, rust/kz-core/src/mapping/go.rs:62-112."""
        me = 1 + (self.ply & 1)
        planes = np.zeros((7 if self.TERRITORY else 4, self.S, self.S), np.uint8)
        owner = self._owners() if self.TERRITORY else None
        for p, v in enumerate(self.stones):
            y, x = divmod(p, self.S)
            planes[0, y, x] = v == me
            planes[1, y, x] = v not in (0, me)
            planes[2, y, x] = 1
            planes[3, y, x] = v == 0 and not self.done() and not self._legal(p)  # a finished board has no unavailable moves (go.rs:84)
            if self.TERRITORY:  # owned by the mover, by nobody, by the other side (go.rs:90-98)
                planes[4 if owner[p] == me else 5 if owner[p] == 0 else 6, y, x] = 1
        komi = F(self.komi_2) * F(0.5)
        black = self.next_player() == 0
        return planes, np.array([black, not black, self.passes == 1, self.passes >= 2, (komi if black else -komi) / F(15.0), self.multi_suicide], F)

    def outcome(self) -> int:
        black = sum(1 for v in self.stones if v == 1)
        white = sum(1 for v in self.stones if v == 2)
        seen = set()
        for p0 in range(self.A):
            if self.stones[p0] or p0 in seen:
                continue
            region, touches, todo = {p0}, set(), [p0]
            while todo:
                q = todo.pop()
                for r in self._neighbours(q):
                    if self.stones[r]:
                        touches.add(self.stones[r])
                    elif r not in region:
                        region.add(r)
                        todo.append(r)
            seen |= region
            if touches == {1}:
                black += len(region)
            elif touches == {2}:
                white += len(region)
        score_2 = 2 * (black - white) - self.komi_2
        return (score_2 > 0) - (score_2 < 0)


class Go9Territory(Go9):
    """Go9 with GoStdMapper::new(9, true), the mapper the reference's self-play server constructs (server.rs:193)."""
    TERRITORY = True

    @staticmethod
    def start(seed: int) -> "Go9Territory":
        g = Go9.start(seed)
        g.__class__ = Go9Territory
        return g

    def clone(self):
        g = Go9.clone(self)
        g.__class__ = Go9Territory
        return g


class Ataxx7:
    """Twin of kzb::selfplay::Ataxx (kzero_b200/csrc/selfplay/games.hpp), written independently with sets of (x, y) instead
    of bitboards: copy to a free tile at distance 1, jump to one at distance exactly 2 (the origin empties), the moved-to
    tile converts adjacent enemy tiles, pass only when nothing else is possible; the game ends when a side is wiped out,
    the board is full, nobody can move, or after 100 moves without a copy (draw).  Move = policy index (ataxx.rs:60-81)."""
    S = 7
    JUMPS = [(-2, -2), (-1, -2), (0, -2), (1, -2), (2, -2), (-2, -1), (2, -1), (-2, 0), (2, 0), (-2, 1), (2, 1), (-2, 2), (-1, 2), (0, 2),
             (1, 2), (2, 2)]  # FROM_DX_DY, ataxx.rs:134-151

    def __init__(self):
        self.tiles = [{(0, 0), (6, 6)}, {(6, 0), (0, 6)}]
        self.ply, self.since_copy, self.finished, self.result = 0, 0, False, 0

    @staticmethod
    def start(seed: int) -> "Ataxx7":
        return Ataxx7()

    def clone(self):
        g = Ataxx7()
        g.tiles = [set(self.tiles[0]), set(self.tiles[1])]
        g.ply, g.since_copy, g.finished, g.result = self.ply, self.since_copy, self.finished, self.result
        return g

    def _bits(self, tiles) -> int:
        return sum(1 << (y * self.S + x) for x, y in tiles)

    def hash(self) -> int:
        inner = (self._bits(self.tiles[1]) * 5 + ((self.ply & 1) << 62) + (self.since_copy << 50)) & M64
        return splitmix64((self._bits(self.tiles[0]) * 3 + splitmix64(inner)) & M64)

    def next_player(self) -> int:
        return self.ply & 1

    def done(self) -> bool:
        return self.finished

    def outcome(self) -> int:
        return self.result

    def _free(self, x: int, y: int) -> bool:
        return 0 <= x < self.S and 0 <= y < self.S and (x, y) not in self.tiles[0] and (x, y) not in self.tiles[1]

    def _reach(self, tiles, distance: int):
        out = set()
        for x, y in tiles:
            for dy in range(-distance, distance + 1):
                for dx in range(-distance, distance + 1):
                    if max(abs(dx), abs(dy)) == distance and self._free(x + dx, y + dy):
                        out.add((x + dx, y + dy))
        return out

    def moves(self) -> List[int]:
        mine = self.tiles[self.ply & 1]
        out = [y * self.S + x for x, y in sorted(self._reach(mine, 1), key=lambda t: t[1] * self.S + t[0])]
        for fx, fy in sorted(mine, key=lambda t: t[1] * self.S + t[0]):
            for tx, ty in sorted(self._reach({(fx, fy)}, 2), key=lambda t: t[1] * self.S + t[0]):
                out.append((1 + self.JUMPS.index((fx - tx, fy - ty))) * self.S * self.S + ty * self.S + tx)
        return out if out else [17 * self.S * self.S]

    def encode(self):
        """-> (bools [3, 7, 7] u8, scalars f32): AtaxxStdMapper::encode_input, rust/kz-core/src/mapping/ataxx.rs:106-115."""
        planes = np.zeros((3, self.S, self.S), np.uint8)
        for k, tiles in enumerate((self.tiles[self.ply & 1], self.tiles[(self.ply & 1) ^ 1])):
            for x, y in tiles:
                planes[k, y, x] = 1
        return planes, np.array([F(self.since_copy) / F(100.0)], F)

    def play(self, mv: int) -> None:
        area = self.S * self.S
        me, other = self.ply & 1, (self.ply & 1) ^ 1
        if mv == 17 * area:
            self.since_copy += 1
        else:
            to = (mv % area % self.S, mv % area // self.S)
            if mv >= area:
                dx, dy = self.JUMPS[mv // area - 1]
                self.tiles[me].discard((to[0] + dx, to[1] + dy))
                self.since_copy += 1
            else:
                self.since_copy = 0
            self.tiles[me].add(to)
            converted = {t for t in self.tiles[other] if max(abs(t[0] - to[0]), abs(t[1] - to[1])) == 1}
            self.tiles[me] |= converted
            self.tiles[other] -= converted
        self.ply += 1
        a, b = len(self.tiles[0]), len(self.tiles[1])
        free_tiles = self.S * self.S - a - b
        any_move = any(self._reach(self.tiles[k], 1) or self._reach(self.tiles[k], 2) for k in (0, 1))
        if a == 0 or b == 0 or free_tiles == 0 or not any_move or self.since_copy >= 100:
            self.finished = True
            stalled_only = self.since_copy >= 100 and a and b and free_tiles and any_move
            self.result = 0 if stalled_only else (a > b) - (a < b)


class Chess:
    """Twin of kzb::selfplay::Chess (kzero_b200/csrc/selfplay/chess_game.hpp), written independently: a mailbox board,
    legality by making the move and looking at the king (no bitboards, no attack maps, no pins, no shortcuts).  What the two
    must share is the SPECIFICATION: moves are policy indices from the mover's side (ranks flipped for black) in the
    reference's flat table (chess.rs:439-481), listed in the canonical order (the search breaks ties by position in this list):
    pawn moves set by set -- pushes, double pushes, captures towards the a-file, captures towards the h-file, each set by
    destination square and each promotion as Q R B N, then en-passant captures by origin -- then knights, bishops, rooks, queens
    and the king (castling is a king move), per piece type by origin, per piece by destination.  The en-passant
    square exists only while an enemy pawn stands next to the pushed pawn; a position repeats when placement, side, castling
    rights and en-passant square agree; draw on the third occurrence, after 100 quiet plies, or with bare kings."""
    KNIGHT = [(2, 1), (1, 2), (-1, 2), (-2, 1), (-2, -1), (-1, -2), (1, -2), (2, -1)]  # (rank, file) steps
    KING = [(1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (1, -1), (-1, 1), (-1, -1)]
    SLIDES = {3: KING[4:], 4: KING[:4], 5: KING}  # bishop, rook, queen
    _flat = None

    def __init__(self):
        back = [4, 2, 3, 5, 6, 3, 2, 4]
        self.sq = back + [1] * 8 + [0] * 32 + [-1] * 8 + [-v for v in back]
        self.side, self.castle, self.ep, self.halfmove, self.ply = 0, 15, -1, 0, 0
        self.history: List[int] = []  # keys of earlier positions since the last irreversible move
        self.reps, self.terminal = 0, 0

    @staticmethod
    def start(seed: int) -> "Chess":
        return Chess()

    def clone(self):
        c = Chess.__new__(Chess)
        c.sq, c.history = list(self.sq), list(self.history)
        c.side, c.castle, c.ep, c.halfmove, c.ply, c.reps, c.terminal = self.side, self.castle, self.ep, self.halfmove, self.ply, self.reps, self.terminal
        return c

    @classmethod
    def flat(cls):
        if cls._flat is None:
            table = []
            for f in range(64):
                for t in range(64):
                    df, dr = f % 8 - t % 8, f // 8 - t // 8
                    if ((df == 0) != (dr == 0)) or (df != 0 and abs(df) == abs(dr)):
                        table.append((f, t, 0))
            for f in range(64):
                for t in range(64):
                    df, dr = abs(f % 8 - t % 8), abs(f // 8 - t // 8)
                    if (df, dr) in ((1, 2), (2, 1)):
                        table.append((f, t, 0))
            for piece in (5, 4, 3, 2):
                for ff in range(8):
                    for tf in range(8):
                        if abs(ff - tf) <= 1:
                            table.append((48 + ff, 56 + tf, piece))
            cls._flat = (table, {m: i for i, m in enumerate(table)})
        return cls._flat

    def _pov(self, s: int) -> int:
        return s if self.side == 0 else (7 - s // 8) * 8 + s % 8

    def position_key(self) -> int:
        h = 0x9E3779B97F4A7C15 if self.side else 0
        for s, p in enumerate(self.sq):
            if p:
                h ^= splitmix64((p + 16) * 64 + s + 0xC0FFEE)
        h ^= splitmix64(0xCA57 + self.castle)
        if self.ep >= 0:
            h ^= splitmix64(0xE9 + self.ep)
        return h

    def hash(self) -> int:
        return splitmix64(self.position_key() ^ (self.halfmove << 8) ^ (self.reps << 20))

    def next_player(self) -> int:
        return self.side

    def done(self) -> bool:
        return self.terminal != 0

    def outcome(self) -> int:
        return (-1 if self.side == 0 else 1) if self.terminal == 1 else 0

    def _attacked(self, s: int, by: int) -> bool:
        sign = 1 if by == 0 else -1
        r, f = divmod(s, 8)
        for df in (-1, 1):
            rr, ff = r - sign, f + df
            if 0 <= rr < 8 and 0 <= ff < 8 and self.sq[rr * 8 + ff] == sign:
                return True
        for steps, piece in ((self.KNIGHT, 2), (self.KING, 6)):
            for dr, df in steps:
                rr, ff = r + dr, f + df
                if 0 <= rr < 8 and 0 <= ff < 8 and self.sq[rr * 8 + ff] == sign * piece:
                    return True
        for i, (dr, df) in enumerate(self.KING):
            rr, ff = r + dr, f + df
            while 0 <= rr < 8 and 0 <= ff < 8:
                p = self.sq[rr * 8 + ff]
                if p:
                    if p * sign in (5, 4 if i < 4 else 3):
                        return True
                    break
                rr, ff = rr + dr, ff + df
        return False

    def _king(self, colour: int) -> int:
        return self.sq.index(6 if colour == 0 else -6)

    def _pseudo(self):
        sign = 1 if self.side == 0 else -1
        enemy = self.side ^ 1
        for s in range(64):
            p = self.sq[s] * sign
            if p <= 0:
                continue
            r, f = divmod(s, 8)
            if p == 1:
                last, first = (7, 1) if self.side == 0 else (0, 6)
                r1 = r + sign
                targets = []
                if not self.sq[r1 * 8 + f]:
                    targets.append(r1 * 8 + f)
                    if r == first and not self.sq[(r + 2 * sign) * 8 + f]:
                        targets.append((r + 2 * sign) * 8 + f)
                for df in (-1, 1):
                    if 0 <= f + df < 8:
                        t = r1 * 8 + f + df
                        if self.sq[t] * sign < 0 or t == self.ep:
                            targets.append(t)
                for t in targets:
                    if t // 8 == last:
                        for promo in (5, 4, 3, 2):
                            yield s, t, promo
                    else:
                        yield s, t, 0
            elif p in (2, 6):
                for dr, df in (self.KNIGHT if p == 2 else self.KING):
                    rr, ff = r + dr, f + df
                    if 0 <= rr < 8 and 0 <= ff < 8 and self.sq[rr * 8 + ff] * sign <= 0:
                        yield s, rr * 8 + ff, 0
                home = 4 if self.side == 0 else 60
                if p == 6 and s == home and not self._attacked(home, enemy):
                    k_right, q_right = (1, 2) if self.side == 0 else (4, 8)
                    if self.castle & k_right and not self.sq[home + 1] and not self.sq[home + 2] and self.sq[home + 3] == 4 * sign \
                            and not self._attacked(home + 1, enemy) and not self._attacked(home + 2, enemy):
                        yield s, home + 2, 0
                    if self.castle & q_right and not any(self.sq[home - 3:home]) and self.sq[home - 4] == 4 * sign \
                            and not self._attacked(home - 1, enemy) and not self._attacked(home - 2, enemy):
                        yield s, home - 2, 0
            else:
                for dr, df in self.SLIDES[p]:
                    rr, ff = r + dr, f + df
                    while 0 <= rr < 8 and 0 <= ff < 8:
                        q = self.sq[rr * 8 + ff] * sign
                        if q > 0:
                            break
                        yield s, rr * 8 + ff, 0
                        if q:
                            break
                        rr, ff = rr + dr, ff + df

    def encode(self):
        """-> (bools [13, 8, 8] u8, scalars f32): ChessStdMapper::encode_input, rust/kz-core/src/mapping/chess.rs:136-170."""
        sign = 1 if self.side == 0 else -1
        planes = np.zeros((13, 8, 8), np.uint8)
        for s, p in enumerate(self.sq):
            if p:
                r, f = divmod(self._pov(s), 8)
                planes[(0 if p * sign > 0 else 6) + abs(p) - 1, r, f] = 1
        if self.ep >= 0:  # the plane marks the pawn that just advanced two ranks (chess 3.2.0: Board::en_passant()), not the target
            r, f = divmod(self._pov(self.ep + (-8 if self.side == 0 else 8)), 8)
            planes[12, r, f] = 1
        own_k, own_q, opp_k, opp_q = (1, 2, 4, 8) if self.side == 0 else (4, 8, 1, 2)
        scalars = [self.side == 0, self.side == 1, bool(self.castle & own_k), bool(self.castle & own_q), bool(self.castle & opp_k),
                   bool(self.castle & opp_q), self.reps, self.halfmove]
        return planes, np.array(scalars, F)

    def _place(self, frm: int, to: int, promo: int) -> None:
        sign = 1 if self.side == 0 else -1
        p = self.sq[frm]
        if abs(p) == 1 and to == self.ep and not self.sq[to]:
            self.sq[(frm // 8) * 8 + to % 8] = 0
        self.sq[to] = sign * promo if promo else p
        self.sq[frm] = 0
        if abs(p) == 6 and abs(to - frm) == 2:
            rook_from, rook_to = (frm + 3, frm + 1) if to > frm else (frm - 4, frm - 1)
            self.sq[rook_to], self.sq[rook_from] = self.sq[rook_from], 0

    def _legal(self):
        for frm, to, promo in self._pseudo():
            c = self.clone()
            c._place(frm, to, promo)
            if not c._attacked(c._king(self.side), self.side ^ 1):
                yield frm, to, promo

    def moves(self) -> List[int]:
        index = self.flat()[1]
        def canonical(m):
            frm, to, promo = m
            piece = abs(self.sq[frm])
            if piece != 1:
                return piece, 0, frm, to, 0
            df = to % 8 - frm % 8
            if df == 0:
                return 1, (0 if abs(to - frm) == 8 else 1), to, 0, -promo
            if not self.sq[to]:
                return 1, 4, frm, 0, 0  # en passant
            return 1, (2 if df < 0 else 3), to, 0, -promo
        ordered = sorted(self._legal(), key=canonical)  # the generator below walks square by square, direction by direction
        return [index[(self._pov(f), self._pov(t), promo)] for f, t, promo in ordered]

    def play(self, mv: int) -> None:
        f, t, promo = self.flat()[0][mv]
        frm, to = self._pov(f), self._pov(t)
        sign = 1 if self.side == 0 else -1
        piece = abs(self.sq[frm])
        capture = self.sq[to] != 0 or (piece == 1 and to == self.ep)
        key_before, castle_before = self.position_key(), self.castle
        self._place(frm, to, promo)
        for s in (frm, to):
            self.castle &= {4: ~3, 60: ~12, 7: ~1, 0: ~2, 63: ~4, 56: ~8}.get(s, 15) & 15
        self.ep = -1
        if piece == 1 and abs(to - frm) == 16:
            file = to % 8
            if (file > 0 and self.sq[to - 1] == -sign) or (file < 7 and self.sq[to + 1] == -sign):
                self.ep = (frm + to) // 2
        if piece == 1 or capture:
            self.halfmove = 0
        else:
            self.halfmove += 1
        if piece == 1 or capture or self.castle != castle_before:
            self.history = []
        else:
            self.history.append(key_before)
        self.side ^= 1
        self.ply += 1
        self.reps = self.history.count(self.position_key())
        others = [abs(p) for p in self.sq if p and abs(p) != 6]
        low_material = not others  # bare kings only (K + minor v K plays on: rust/kz-core/tests/mapper/chess/pairs.rs:98-136)
        if not any(True for _ in self._legal()):
            self.terminal = 1 if self._attacked(self._king(self.side), self.side ^ 1) else 2
        else:
            self.terminal = 2 if (self.halfmove >= 100 or self.reps >= 2 or low_material) else 0


def pseudo_eval(board, kind: int):
    """-> (values_pov [value, win, draw, loss, moves_left] f32, policy f32); twin of pseudo_eval in selfplay.cpp."""
    n = len(board.moves())
    if kind == 0:  # DummyNetwork, rust/kz-core/src/network/dummy.rs:44-60
        third = F(1.0) / F(3.0)
        return np.array([0, third, third, third, 0], F), np.full(n, F(1.0) / F(n), F)
    h = board.hash()
    raw = np.array([(splitmix64((h + i + 1) & M64) >> 40) % 1000 + 1 for i in range(n)], F)
    total = F(0)
    for r in raw:
        total = F(total + r)
    v = F(F(int((splitmix64(h ^ 0xABCD) >> 40) % 2001) - 1000) / F(1000.0))
    win = F(F(F(F(1.0) + v) * F(0.5)) * F(0.8))
    loss = F(F(F(F(1.0) - v) * F(0.5)) * F(0.8))
    return np.array([v, win, F(0.2), loss, F((h >> 50) % 50)], F), (raw / total).astype(F)


@dataclass
class Settings:
    exploration_weight: float = 2.0
    moves_left_weight: float = 0.03
    moves_left_clip: float = 20.0
    moves_left_sharpness: float = 0.5
    q_mode_wdl: bool = True
    draw_score: float = 0.0
    fpu_root: float = 0.1
    fpu_root_relative: bool = False
    fpu_child: float = 0.0
    fpu_child_relative: bool = True
    virtual_loss: float = 1.0
    policy_temperature_root: float = 1.0
    policy_temperature_child: float = 1.0


@dataclass
class Node:
    parent: int = -1
    last_move: int = 0
    children: Optional[range] = None
    complete_visits: int = 0
    virtual_visits: int = 0
    sum_values: np.ndarray = field(default_factory=lambda: np.zeros(5, F))  # abs: value, win_a, draw, win_b, moves_left
    net_values: Optional[np.ndarray] = None
    net_policy: np.float32 = F(np.nan)

    def total_visits(self) -> int:
        return self.complete_visits + self.virtual_visits

    def values(self) -> np.ndarray:
        return (self.sum_values / F(self.complete_visits)).astype(F)


def pov(v: np.ndarray, player: int) -> np.ndarray:  # values.rs:21-40 (its own inverse)
    return v if player == 0 else np.array([-v[0], v[3], v[2], v[1], v[4]], F)


def q_select(s: Settings, v: np.ndarray) -> np.float32:  # step.rs:237-242
    return F(F(v[1] + F(F(s.draw_score) * v[2])) - v[3]) if s.q_mode_wdl else v[0]


class Tree:
    def __init__(self, root_board):
        assert not root_board.done()
        self.root_board = root_board
        self.nodes: List[Node] = [Node()]

    def uct_context(self, idx: int):
        n = self.nodes[idx]
        mass = F(0)
        for c in n.children:
            if self.nodes[c].total_visits() > 0:
                mass = F(mass + self.nodes[c].net_policy)
        return n.total_visits(), n.values(), mass

    def uct_total(self, child: Node, ctx, fpu_relative: bool, fpu_value: float, s: Settings, player: int) -> np.float32:
        parent_total, parent_values, mass = ctx
        if fpu_relative:
            fpu = F(q_select(s, pov(parent_values, player)) - F(F(fpu_value) * np.sqrt(mass)))
        else:
            fpu = F(fpu_value)
        vl = F(s.virtual_loss)
        tvv = F(F(child.complete_visits) + F(vl * F(child.virtual_visits)))
        if tvv == 0:
            q = fpu
        else:
            total_value = q_select(s, pov(child.sum_values, player))
            q = F(F(total_value - F(vl * F(child.virtual_visits))) / tvv)
        u = F(F(child.net_policy * np.sqrt(F(parent_total - 1))) / F(1 + child.total_visits()))
        m = F(0) if child.complete_visits == 0 else F(child.values()[4] - F(parent_values[4] - F(1.0)))
        m_unit = F(0)
        if s.moves_left_weight != 0.0:
            clip = F(s.moves_left_clip)
            m_clipped = min(max(m, F(-clip)), clip)
            m_unit = min(max(F(F(F(s.moves_left_sharpness) * m_clipped) * F(-q)), F(-1.0)), F(1.0))
        return F(F(q + F(F(s.exploration_weight) * u)) + F(F(s.moves_left_weight) * m_unit))

    def propagate(self, idx: int, values: np.ndarray) -> None:  # step.rs:171-188
        cur = idx
        values = values.copy()
        while True:
            n = self.nodes[cur]
            assert n.virtual_visits > 0
            n.complete_visits += 1
            n.virtual_visits -= 1
            n.sum_values = (n.sum_values + values).astype(F)
            if n.parent < 0:
                break
            cur = n.parent
            values[4] = F(values[4] + F(1.0))


def zero_step_gather(tree: Tree, s: Settings, rng: Rng):
    """-> (node index, board) of the reached un-evaluated node, or None after propagating a terminal outcome."""
    cur = 0
    board = tree.root_board.clone()
    while True:
        tree.nodes[cur].virtual_visits += 1
        if board.done():
            o = board.outcome()
            tree.propagate(cur, np.array([o, 1.0 if o > 0 else 0.0, 1.0 if o == 0 else 0.0, 1.0 if o < 0 else 0.0, 0.0], F))
            return None
        n = tree.nodes[cur]
        if n.children is None:
            mvs = board.moves()
            p = F(F(1.0) / F(len(mvs)))
            start = len(tree.nodes)
            for mv in mvs:
                tree.nodes.append(Node(parent=cur, last_move=mv, net_policy=p))
            n.children = range(start, start + len(mvs))
            n.net_values = None
            return cur, board
        player = board.next_player()
        selected, ties, best = -1, 0, None
        if n.complete_visits == 0:
            for c in n.children:
                v = tree.nodes[c].total_visits()
                if selected < 0 or v < best:
                    selected, best, ties = c, v, 1
                elif v == best:
                    ties += 1
                    if rng.gen_range(ties) == 0:
                        selected = c
        else:
            rel, val = (s.fpu_root_relative, s.fpu_root) if cur == 0 else (s.fpu_child_relative, s.fpu_child)
            ctx = tree.uct_context(cur)
            for c in n.children:
                u = tree.uct_total(tree.nodes[c], ctx, rel, val, s, player)
                assert not np.isnan(u)
                if selected < 0 or u > best:
                    selected, best, ties = c, u, 1
                elif u == best:
                    ties += 1
                    if rng.gen_range(ties) == 0:
                        selected = c
        cur = selected
        board.play(tree.nodes[cur].last_move)


def zero_step_apply(tree: Tree, idx: int, next_player: int, values_pov: np.ndarray, policy: np.ndarray) -> None:
    n = tree.nodes[idx]
    assert n.net_values is None
    abs_values = pov(values_pov, next_player)
    n.net_values = abs_values
    tree.propagate(idx, abs_values)
    assert n.children is not None and len(n.children) == len(policy)
    for c, p in zip(n.children, policy):
        tree.nodes[c].net_policy = F(p)


def search(game_seed: int, plies: int, rng_seed: int, visits: int, search_batch: int, eval_kind: int, s: Settings, game: str = "chess"):
    """The twin of trace_search in selfplay.cpp: -> dict(child_visits, child_moves, child_policy, root_values, ...)."""
    board = {"go-9": Go9, "go-9-territory": Go9Territory, "ataxx-7": Ataxx7, "chess": SynthChess, "chess-real": Chess}[game].start(game_seed)
    rng = Rng(rng_seed)
    for _ in range(plies):
        if board.done():
            break
        mvs = board.moves()
        board.play(mvs[rng.gen_range(len(mvs))])
    tree = Tree(board)
    evals = 0
    while tree.nodes[0].complete_visits < visits:
        requests, terminal = [], 0
        while len(requests) < search_batch and terminal < search_batch:
            r = zero_step_gather(tree, s, rng)
            if r is None:
                terminal += 1
            else:
                requests.append(r)
        for idx, b in requests:
            values, policy = pseudo_eval(b, eval_kind)
            t = s.policy_temperature_root if idx == 0 else s.policy_temperature_child
            if t != 1.0:
                policy = np.power(policy, F(1.0 / t)).astype(F)
                total = F(0)
                for p in policy:
                    total = F(total + p)
                policy = (policy / total).astype(F)
            zero_step_apply(tree, idx, b.next_player(), values, policy)
            evals += 1
    root = tree.nodes[0]
    kids = [tree.nodes[c] for c in root.children]
    return dict(child_visits=np.array([k.complete_visits for k in kids], np.uint64),
                child_moves=np.array([k.last_move for k in kids], np.uint32),
                child_policy=np.array([k.net_policy for k in kids], F),
                root_values=pov(root.values(), board.next_player()), root_visits=root.complete_visits,
                tree_nodes=len(tree.nodes), evals=evals)
