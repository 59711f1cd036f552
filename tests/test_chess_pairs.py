"""The reference's own known answers for the chess policy mapping, rust/kz-core/tests/mapper/chess/pairs.rs:17-385 (all 11
tests) and tests/mapper/chess/mod.rs:6-17, held against the C++ chess of the self-play driver (chess_game.hpp: Chess::from_fen
-> legal moves -> flat policy index from the mover's point of view) -- rows A2 / A7 / N1 for chess.

pairs.rs states every expectation as (index of ChessLegacyConvPolicyMapper, move): index = channel * 64 + from-square, both
from the mover's point of view.  The flat index the C++ side produces is taken there through the reference's OWN table
chess_flat_to_conv (the Gather constant inside tests/golden/net_chess_conv_2x32.onnx, exported by the reference's code =
python/lib/mapping/chess_flat_to_conv.txt).  Both directions are checked like test_pairs does (pairs.rs:354-385): the move
is available and maps to the index; the index maps back to that move.  test_valid_policy_mapping (tests/mapper/mod.rs:29-72):
no two legal moves share an index.  Squares: a1 = 0 ... h8 = 63.
"""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle.onnx_min import load_model

ROOT = Path(__file__).resolve().parent
GOLDEN = ROOT / "golden"


def sq(name):
    return (int(name[1]) - 1) * 8 + "abcdefgh".index(name[0].lower())


def mv(frm, to, promo="-"):
    return (sq(frm), sq(to), promo)


D4 = sq("d4")
PAIRS = {
    "queen_distance_white": ("8/8/8/6k1/8/6K1/8/Q7 w - - 0 1",
                             [(i * 64, mv("a1", f"a{i + 2}")) for i in range(7)]),
    "queen_distance_black": ("q7/8/8/6k1/8/6K1/8/8 b - - 0 1",
                             [(i * 64, mv("a8", f"a{7 - i}")) for i in range(7)]),
    "queen_direction_white": ("8/8/6k1/8/3Q4/6K1/8/8 w - - 0 1",
                              [(d * 7 * 64 + D4, mv("d4", t)) for d, t in enumerate(["d5", "e5", "e4", "e3", "d3", "c3", "c4", "c5"])]),
    "queen_direction_black": ("8/8/6k1/3q4/8/6K1/8/8 b - - 0 1",
                              [(d * 7 * 64 + D4, mv("d5", t)) for d, t in enumerate(["d4", "e4", "e5", "e6", "d6", "c6", "c5", "c4"])]),
    "knight_direction_white": ("8/8/6k1/8/3N4/6K1/8/8 w - - 0 1",
                               [((56 + k) * 64 + D4, mv("d4", t)) for k, t in enumerate(["e6", "f5", "f3", "e2", "c2", "b3", "b5", "c6"])]),
    "knight_direction_black": ("8/8/6k1/3n4/8/6K1/8/8 b - - 0 1",
                               [((56 + k) * 64 + D4, mv("d5", t)) for k, t in enumerate(["e3", "f4", "f6", "e7", "c7", "b6", "b4", "c3"])]),
    "white_potential_promotions": ("r1r5/1P4R1/5RNP/2k5/5K2/pnr5/1r4p1/5R1R w - - 0 1", [
        ((0 * 7 + 1) * 64 + sq("f6"), mv("f6", "f8")), ((0 * 7 + 0) * 64 + sq("g7"), mv("g7", "g8")),
        (63 * 64 + sq("g6"), mv("g6", "f8")), (56 * 64 + sq("g6"), mv("g6", "h8")),
        ((7 * 7 + 0) * 64 + sq("b7"), mv("b7", "a8", "q")), ((0 * 7 + 0) * 64 + sq("b7"), mv("b7", "b8", "q")),
        ((1 * 7 + 0) * 64 + sq("b7"), mv("b7", "c8", "q")),
        (64 * 64 + sq("b7"), mv("b7", "a8", "r")), (67 * 64 + sq("b7"), mv("b7", "b8", "r")), (70 * 64 + sq("b7"), mv("b7", "c8", "r")),
        (65 * 64 + sq("b7"), mv("b7", "a8", "b")), (68 * 64 + sq("b7"), mv("b7", "b8", "b")), (71 * 64 + sq("b7"), mv("b7", "c8", "b")),
        (66 * 64 + sq("b7"), mv("b7", "a8", "n")), (69 * 64 + sq("b7"), mv("b7", "b8", "n")), (72 * 64 + sq("b7"), mv("b7", "c8", "n"))]),
    # "careful, move indices are from the POV of black!" (pairs.rs:211)
    "black_potential_promotions": ("r1r5/1P4R1/5RNP/2k5/5K2/pnr5/1r4p1/5R1R b - - 0 1", [
        ((0 * 7 + 1) * 64 + sq("c6"), mv("c3", "c1")), ((0 * 7 + 0) * 64 + sq("b7"), mv("b2", "b1")),
        (56 * 64 + sq("b6"), mv("b3", "c1")), (63 * 64 + sq("b6"), mv("b3", "a1")),
        ((7 * 7 + 0) * 64 + sq("g7"), mv("g2", "f1", "q")), ((0 * 7 + 0) * 64 + sq("g7"), mv("g2", "g1", "q")),
        ((1 * 7 + 0) * 64 + sq("g7"), mv("g2", "h1", "q")),
        (67 * 64 + sq("g7"), mv("g2", "g1", "r")), (70 * 64 + sq("g7"), mv("g2", "h1", "r")), (64 * 64 + sq("g7"), mv("g2", "f1", "r")),
        (68 * 64 + sq("g7"), mv("g2", "g1", "b")), (71 * 64 + sq("g7"), mv("g2", "h1", "b")), (65 * 64 + sq("g7"), mv("g2", "f1", "b")),
        (69 * 64 + sq("g7"), mv("g2", "g1", "n")), (72 * 64 + sq("g7"), mv("g2", "h1", "n")), (66 * 64 + sq("g7"), mv("g2", "f1", "n"))]),
    "en_passant_white": ("8/8/5k2/1pP5/8/5K2/8/8 w - b6 0 2", [((7 * 7 + 0) * 64 + sq("c5"), mv("c5", "b6"))]),
    "en_passant_black": ("8/8/5k2/8/1pP5/5K2/8/8 b - c3 0 1", [((1 * 7 + 0) * 64 + sq("b5"), mv("b4", "c3"))]),
    "castles_white": ("r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1",
                      [((2 * 7 + 1) * 64 + sq("e1"), mv("e1", "g1")), ((6 * 7 + 1) * 64 + sq("e1"), mv("e1", "c1"))]),
    "castles_black": ("r3k2r/8/8/8/8/8/8/R3K2R b KQkq - 0 1",
                      [((2 * 7 + 1) * 64 + sq("e1"), mv("e8", "g8")), ((6 * 7 + 1) * 64 + sq("e1"), mv("e8", "c8"))]),
    "basic_board_mapping": ("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", []),
}


@pytest.fixture(scope="module")
def dump(tmp_path_factory):
    exe = tmp_path_factory.mktemp("chess") / "chess_fen_dump"
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", str(exe), str(ROOT / "cpp" / "chess_fen_dump.cpp")], check=True)
    fens = [v[0] for v in PAIRS.values()]
    out = subprocess.run([str(exe)], input="\n".join(fens) + "\n", capture_output=True, text=True, check=True).stdout
    boards, cur = {}, None
    for line in out.splitlines():
        key, _, rest = line.partition(" ")
        if key == "fen":
            cur = boards.setdefault(rest, {"moves": []})
        elif key == "mv":
            f, t, p, idx = rest.split()
            cur["moves"].append(((int(f), int(t), p), int(idx)))
        elif key in ("ep_plane", "own_pawns", "done"):
            cur[key] = int(rest)
        elif key == "scalars":
            cur["scalars"] = [float(v) for v in rest.split()]
    return boards


@pytest.fixture(scope="module")
def flat_to_conv():
    m = load_model((GOLDEN / "net_chess_conv_2x32.onnx").read_bytes())
    node = [n for n in m.nodes if n.op == "Gather" and n.outputs == ["policy"]][0]
    const = [n for n in m.nodes if n.op == "Constant" and n.outputs[0] == node.inputs[1]][0]
    table = np.asarray(const.attrs["value"]).astype(np.int64)
    # 1880 flat moves (mod.rs:6-17; their distinctness is checked in test_host_units.py against the reference's table); a queen
    # promotion shares the conv index of the plain pawn move (chess.rs:318-324), so the conv indices are fewer
    assert table.shape == (1880,) and len(set(table.tolist())) == 1880 - 22
    return table


@pytest.mark.parametrize("name", list(PAIRS))
def test_pairs(name, dump, flat_to_conv):
    fen, pairs = PAIRS[name]
    board = dump[fen]
    assert board["done"] == 0
    flat = [idx for _, idx in board["moves"]]
    assert len(set(flat)) == len(flat) and all(0 <= i < 1880 for i in flat)  # test_move_to_index: no duplicate indices
    assert len(board["scalars"]) == 8                                          # test_valid_input_mapping
    by_move = dict(board["moves"])
    by_conv = {int(flat_to_conv[idx]): m for m, idx in board["moves"]}
    for index, move in pairs:
        assert move in by_move, f"{name}: move {move} is not available on the board"  # pairs.rs:367-371
        assert int(flat_to_conv[by_move[move]]) == index, f"{name}: wrong index for move {move}"  # pairs.rs:377
        assert by_conv.get(index) == move, f"{name}: index {index} does not map back to {move}"  # pairs.rs:386-392


def test_known_move_counts(dump):
    assert len(dump[PAIRS["basic_board_mapping"][0]]["moves"]) == 20
    assert len(dump[PAIRS["castles_white"][0]]["moves"]) == 26 and len(dump[PAIRS["castles_black"][0]]["moves"]) == 26


def test_en_passant_plane_marks_the_pushed_pawn(dump):
    """mapping/chess.rs:168 encodes `inner.en_passant()`; in the chess 3.2.0 crate that is the square of the pawn that just
    advanced two ranks (the capture lands on ep_sq.uforward(side_to_move)), seen from the mover: b5 for white's FEN above,
    c4 -> flipped to c5 for black's."""
    assert dump[PAIRS["en_passant_white"][0]]["ep_plane"] == 1 << sq("b5")
    assert dump[PAIRS["en_passant_black"][0]]["ep_plane"] == 1 << sq("c5")
    assert dump[PAIRS["en_passant_black"][0]]["own_pawns"] == 1 << sq("b5")  # black's b4 pawn from black's point of view
    assert dump[PAIRS["queen_distance_white"][0]]["ep_plane"] == 0
