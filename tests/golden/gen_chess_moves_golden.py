"""Writes tests/golden/chess_flat_moves.json from the reference's own table python/lib/mapping/chess_flat_to_move_input.txt
(rows written by ChessStdMapper::encode_mv, rust/kz-core/src/mapping/chess.rs:494-519: from, to, 0, promotion Q, R, B, N, none):
for each of the 1880 policy indices [from, to, promotion] with promotion in "qrbn" or "".  Run in the build container."""
import json
from pathlib import Path

import numpy as np

rows = np.genfromtxt("/root/reference/python/lib/mapping/chess_flat_to_move_input.txt", delimiter=",", dtype=np.int32)[:, :8]
assert rows.shape == (1880, 8) and (rows[:, 3:8].sum(axis=1) == 1).all()
out = [[int(r[0]), int(r[1]), "qrbn"[int(np.argmax(r[3:7]))] if r[7] == 0 else ""] for r in rows]
Path(__file__).with_name("chess_flat_moves.json").write_text(json.dumps(out, separators=(",", ":")))
print("wrote", len(out), "moves;", sum(1 for m in out if m[2]), "promotions")
