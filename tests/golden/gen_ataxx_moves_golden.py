"""Writes tests/golden/ataxx7_moves.json from the reference's own tables (python/lib/mapping/ataxx_index_to_move_input.txt,
ataxx_valid.txt, loaded through python/lib/mapping/mapping.py): for 7x7 ataxx, every policy index -> (is_pass, copy_to,
jump_from, jump_to) and the list of indices that are moves.  Run in the build container (needs /root/reference)."""
import json
import sys
from pathlib import Path

sys.path.insert(0, "/root/reference/python")
from lib.mapping.mapping import ATAXX_INDEX_TO_MOVE_INPUT, ATAXX_VALID_MOVES  # noqa: E402

SIZE = 7
table = ATAXX_INDEX_TO_MOVE_INPUT[SIZE - 2].astype(int)
assert table.shape == (17 * SIZE * SIZE + 1, 4)
out = {"size": SIZE, "valid": [int(v) for v in ATAXX_VALID_MOVES[SIZE - 2]], "index_to_move_input": table.tolist()}
Path(__file__).with_name("ataxx7_moves.json").write_text(json.dumps(out, separators=(",", ":")))
print("wrote", len(out["valid"]), "valid moves of", len(out["index_to_move_input"]), "indices")
