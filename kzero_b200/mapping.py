"""Host-side policy index tables for chess, restated from the reference's Rust mappers.

These are needed wherever a chess net is built outside the reference (synthetic nets for the bench
and the GPU tests): the legacy conv policy head ends in `Gather(flat_to_conv)` and the attention head
in `Gather(flat_to_att)` (python/lib/model/post_act.py:69-88,115-141).

Restated from:
  generate_all_flat_moves_pov   rust/kz-core/src/mapping/chess.rs:439-481  (1880 POV moves)
  ClassifiedPovMove::from_move / to_channel   chess.rs:305-357             (73-channel conv index)
  flat_to_att                   rust/kz-misc/src/bin/write_chess_mapping.rs:50-66
Square index = rank*8 + file, A1 = 0.  tests/test_netgen.py checks these tables against the Gather
constants inside the golden ONNX fixtures exported from the reference (tests/golden/).
"""
from __future__ import annotations

from functools import lru_cache
from typing import List, Optional, Tuple

import numpy as np

FLAT_MOVE_COUNT = 1880
CONV_POLICY_CHANNELS = 73

# clockwise starting from NNE / N (chess.rs:425-431), as (rank delta, file delta)
KNIGHT_DELTAS = [(2, 1), (1, 2), (-1, 2), (-2, 1), (-2, -1), (-1, -2), (1, -2), (2, -1)]
QUEEN_DIRECTIONS = [(1, 0), (1, 1), (0, 1), (-1, 1), (-1, 0), (-1, -1), (0, -1), (1, -1)]
# promotion piece codes
QUEEN, ROOK, BISHOP, KNIGHT = "q", "r", "b", "n"
UNDERPROMOTION_PIECES = [ROOK, BISHOP, KNIGHT]

Move = Tuple[int, int, Optional[str]]  # (from square, to square, promotion)


def _sign(v: int) -> int:
    return (v > 0) - (v < 0)


@lru_cache(maxsize=None)
def chess_flat_moves_pov() -> List[Move]:
    result: List[Move] = []
    for frm in range(64):  # queen-like moves
        for to in range(64):
            df = frm % 8 - to % 8
            dr = frm // 8 - to // 8
            if ((df == 0) != (dr == 0)) or (df != 0 and abs(df) == abs(dr)):
                result.append((frm, to, None))
    for frm in range(64):  # knight moves
        for to in range(64):
            df = frm % 8 - to % 8
            dr = frm // 8 - to // 8
            if (abs(df) == 1 and abs(dr) == 2) or (abs(df) == 2 and abs(dr) == 1):
                result.append((frm, to, None))
    for piece in [QUEEN, ROOK, BISHOP, KNIGHT]:  # promotions, rank 7 -> rank 8
        for from_f in range(8):
            for to_f in range(8):
                if abs(from_f - to_f) <= 1:
                    result.append((6 * 8 + from_f, 7 * 8 + to_f, piece))
    assert len(result) == FLAT_MOVE_COUNT
    return result


def chess_conv_channel(mv: Move) -> int:
    frm, to, promo = mv
    rank_delta = to // 8 - frm // 8
    file_delta = to % 8 - frm % 8
    if promo in UNDERPROMOTION_PIECES:
        return 56 + 8 + (_sign(file_delta) + 1) * 3 + UNDERPROMOTION_PIECES.index(promo)
    d = (_sign(rank_delta), _sign(file_delta))
    if d in QUEEN_DIRECTIONS:
        direction = QUEEN_DIRECTIONS.index(d)
        distance = max(abs(rank_delta), abs(file_delta))
        if rank_delta == d[0] * distance and file_delta == d[1] * distance:
            return direction * 7 + (distance - 1)
    return 56 + KNIGHT_DELTAS.index((rank_delta, file_delta))


@lru_cache(maxsize=None)
def chess_flat_to_conv() -> np.ndarray:
    """flat index -> channel*64 + from_square (chess.rs:224-236)."""
    return np.array([chess_conv_channel(mv) * 64 + mv[0] for mv in chess_flat_moves_pov()], dtype=np.int64)


@lru_cache(maxsize=None)
def chess_flat_to_att() -> np.ndarray:
    """flat index -> from*88 + to, promotions at 64 + to_file*3 + p (write_chess_mapping.rs:50-66)."""
    out = []
    for frm, to, promo in chess_flat_moves_pov():
        if promo is None:
            att_to = to
        else:
            att_to = 64 + (to % 8) * 3 + [QUEEN, ROOK, BISHOP, KNIGHT].index(promo)
        out.append(frm * 88 + att_to)
    return np.array(out, dtype=np.int64)
