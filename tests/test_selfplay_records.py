"""Self-play record files written by the C++ driver (row N2), read back with the REFERENCE'S OWN loader
(python/lib/data/file.py, position.py -- imported from /root/reference when present) and with an independent parser of
the format of rust/kz-selfplay/src/binary_output.rs:128-297.  The driver runs with the DummyNetwork stand-in
(uniform evaluations, rust/kz-core/src/network/dummy.rs:44-60), so no GPU is needed."""
import json
import struct
import sys
from pathlib import Path

import numpy as np
import pytest

from kzero_b200 import selfplay

REFERENCE_PY = Path("/root/reference/python")
SCALARS = 26


def _run(tmp_path, game, **kw):
    prefix = str(tmp_path / "games_0")
    if game in (selfplay.GAME_GO9, selfplay.GAME_GO9_TERRITORY, selfplay.GAME_CHESS):
        kw.setdefault("max_game_length", 20)  # go games are long: let the length cap end them (max_game_length, generator_alphazero.rs:125)
    cfg = selfplay.default_config(game=game, visits=24, search_batch=4, gpu_batch=32, cpu_threads=2, gpu_threads=1, max_moves=400,
                                  duration_s=30.0, dummy_network=1, output_prefix=prefix, seed=5, **kw)
    r = selfplay.run(None, cfg)
    assert r.games_written > 0 and r.moves_played >= 400
    return prefix, r


def _parse(prefix, bool_count, scalar_count):
    """Independent reader: -> (meta, list of per-position dicts)."""
    meta = json.loads(Path(prefix + ".json").read_text())
    data = Path(prefix + ".bin").read_bytes()
    off = np.frombuffer(Path(prefix + ".off").read_bytes(), dtype="<u8")
    n = meta["position_count"]
    assert off.size == n + meta["game_count"]  # offsets, then one start index per game (binary_output.rs:270)
    ends = list(off[1:n]) + [len(data)]
    out = []
    for start, end in zip(off[:n], ends):
        rec = data[int(start):int(end)]
        s = np.frombuffer(rec[:SCALARS * 4], "<f4")
        pos = SCALARS * 4
        bits = np.frombuffer(rec[pos:pos + (bool_count + 7) // 8], np.uint8)
        pos += (bool_count + 7) // 8
        scal = np.frombuffer(rec[pos:pos + scalar_count * 4], "<f4")
        pos += scalar_count * 4
        k = int(s[8])
        idx = np.frombuffer(rec[pos:pos + 4 * k], "<u4")
        pos += 4 * k
        val = np.frombuffer(rec[pos:pos + 4 * k], "<f4")
        pos += 4 * k
        assert pos == len(rec)  # every record is consumed to the byte (python/lib/data/taker.py:10-11)
        out.append(dict(scalars=s, bits=bits, input_scalars=scal, indices=idx, values=val))
    return meta, out, off[n:]


@pytest.mark.parametrize("game,name,bool_shape,scalar_count,policy_len", [
    (selfplay.GAME_SYNTH_CHESS, "chess", [13, 8, 8], 8, 1880), (selfplay.GAME_ATAXX7, "ataxx-7", [3, 7, 7], 1, 17 * 49 + 1),
    (selfplay.GAME_GO9, "go-9", [4, 9, 9], 6, 82), (selfplay.GAME_GO9_TERRITORY, "go-9", [7, 9, 9], 6, 82),
    (selfplay.GAME_CHESS, "chess", [13, 8, 8], 8, 1880)])
def test_record_files_are_self_consistent(tmp_path, game, name, bool_shape, scalar_count, policy_len):
    prefix, r = _run(tmp_path, game)
    meta, positions, game_starts = _parse(prefix, int(np.prod(bool_shape)), scalar_count)
    assert meta["game"] == name and meta["input_bool_shape"] == bool_shape and meta["input_scalar_count"] == scalar_count
    assert meta["policy_shape"] == [policy_len] and meta["game_count"] == r.games_written and len(meta["scalar_names"]) == SCALARS
    assert meta["includes_terminal_positions"] and meta["includes_game_start_indices"]
    assert abs(sum(meta["root_wdl"]) - 1) < 1e-6
    # the checks of python/lib/data/check.py:9-76: games cover all positions, final-position flags, per-game indices
    pi = 0
    lengths = []
    for g in range(meta["game_count"]):
        assert game_starts[g] == pi
        length = int(positions[pi]["scalars"][2])
        lengths.append(length)
        for k in range(length + 1):
            s = positions[pi]["scalars"]
            assert (int(s[0]), int(s[1]), int(s[2])) == (g, k, length)
            final = k == length
            assert bool(s[5]) == final
            if final:
                assert s[8] == 0 and s[9] == -1 and np.isnan(s[10]) and bool(s[6]) != bool(s[7])
                assert len(positions[pi]["indices"]) == 0
            else:
                p = positions[pi]
                assert s[3] >= 24 and s[4] == 1 and len(p["indices"]) == int(s[8]) > 0
                assert abs(float(p["values"].sum()) - 1) < 1e-3 and int(s[9]) in p["indices"].tolist()
                assert p["indices"].max() < policy_len and len(set(p["indices"].tolist())) == len(p["indices"])
                assert abs(float(s[12:15].sum()) - 1) < 1e-3 and abs(float(s[17:20].sum()) - 1) < 1e-3 and abs(float(s[22:25].sum()) - 1) < 1e-3
                assert s[15] == length + 1 - k  # final_moves_left, binary_output.rs:164
            pi += 1
    assert pi == meta["position_count"] and max(lengths) == meta["max_game_length"] and min(lengths) == meta["min_game_length"]


@pytest.mark.skipif(not REFERENCE_PY.exists(), reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("game,name", [(selfplay.GAME_SYNTH_CHESS, "chess"), (selfplay.GAME_ATAXX7, "ataxx-7"), (selfplay.GAME_GO9, "go-9"),
                                       (selfplay.GAME_CHESS, "chess")])
def test_reference_loader_reads_the_records(tmp_path, game, name):
    """DataFile.open + Position (python/lib/data/file.py:62-130, position.py:34-103) accept the files and agree with the
    independent parser on every field they decode."""
    prefix, r = _run(tmp_path, game)
    sys.path.insert(0, str(REFERENCE_PY))
    try:
        from lib.data.file import DataFile
        from lib.games import Game
    finally:
        sys.path.pop(0)
    ref_game = Game.find(name)
    f = DataFile.open(ref_game, prefix)
    meta, positions, _ = _parse(prefix, int(np.prod(ref_game.input_bool_shape)), ref_game.input_scalar_channels)
    assert f.info.simulation_count == r.games_written and f.info.position_count == len(positions)
    assert f.info.includes_final_positions and f.info.includes_simulation_start_indices
    for pi in list(range(0, len(positions), 7)) + [len(positions) - 1]:
        p = f.load_position(pi)
        mine = positions[pi]
        s = mine["scalars"]
        assert (p.simulation.index, p.move_index, p.simulation.move_count) == (int(s[0]), int(s[1]), int(s[2]))
        assert p.zero_visits == int(s[3]) and p.available_mv_count == int(s[8]) and p.played_mv == int(s[9])
        assert p.is_final == bool(s[5]) and p.is_terminal == bool(s[6])
        assert np.array_equal(p.policy_indices, mine["indices"].astype(np.int32)) and np.array_equal(p.policy_values, mine["values"])
        assert np.array_equal(p.input_scalars, mine["input_scalars"])
        bools = np.unpackbits(mine["bits"], bitorder="little")[:p.input_bools.size].reshape(p.input_bools.shape)
        assert np.array_equal(p.input_bools, bools)
        assert p.final_v in (-1.0, 0.0, 1.0) and abs(p.final_wdl.sum() - 1) < 1e-6
    # the simulations view walks the whole file through the start indices
    sims = [f.simulations[i] for i in range(len(f.simulations))]
    assert sum(sim.position_count for sim in sims) == f.info.position_count


def replay_under_oracle_rules(prefix, name, twin):
    """Replays every game of a record file move by move with the oracle's independent restatement of the rules: at each position
    the recorded input planes and scalars are the twin's encoding, the recorded policy indices are the twin's legal moves in
    order, and the recorded played move leads to the next recorded position.  -> number of moves checked."""
    from oracle import mcts_oracle as mo

    shape = {"ataxx-7": (3, 7, 7), "go-9": (7, 9, 9) if twin == "Go9Territory" else (4, 9, 9), "chess": (13, 8, 8)}[name]
    meta, positions, game_starts = _parse(prefix, int(np.prod(shape)), {"ataxx-7": 1, "go-9": 6, "chess": 8}[name])
    checked = 0
    for g in range(meta["game_count"]):
        first = int(game_starts[g])
        length = int(positions[first]["scalars"][2])
        # the driver reseeds every new game; go's komi and rule set are the seed-dependent parts of a start position and are in the record
        board = getattr(mo, twin).start(0)
        if name == "go-9":
            komi_pov = float(positions[first]["input_scalars"][4]) * 15.0
            board.komi_2 = int(round(2 * komi_pov))  # black moves first: komi from black's side
            board.multi_suicide = int(positions[first]["input_scalars"][5])
        for k in range(length + 1):
            p = positions[first + k]
            planes, scalars = board.encode()
            bools = np.unpackbits(p["bits"], bitorder="little")[:planes.size].reshape(planes.shape)
            assert np.array_equal(bools, planes), (g, k)
            assert np.array_equal(p["input_scalars"], scalars), (g, k, p["input_scalars"], scalars)
            if k == length:  # the final position carries no policy
                assert bool(p["scalars"][5]) and bool(p["scalars"][6]) == board.done()
                break
            assert p["indices"].tolist() == board.moves(), (g, k)
            assert int(p["scalars"][9]) in p["indices"].tolist()         # the played move is one of the available moves
            assert abs(float(p["values"].sum()) - 1.0) < 1e-3, (g, k)  # the visit distribution over them
            board.play(int(p["scalars"][9]))
            checked += 1
        # the recorded result of the game, from every position's side to move (Outcome::Draw when the length cap ended it,
        # binary_output.rs:141-164)
        outcome = board.outcome() if board.done() else 0
        for k in range(length + 1):
            sc = positions[first + k]["scalars"]
            pov = outcome if k % 2 == 0 else -outcome  # player A moves first in all three games
            assert sc[11] == pov and sc[12:15].tolist() == [float(pov > 0), float(pov == 0), float(pov < 0)], (g, k, sc[11:15], pov)
    return checked


@pytest.mark.parametrize("game,name,twin", [(selfplay.GAME_ATAXX7, "ataxx-7", "Ataxx7"), (selfplay.GAME_GO9, "go-9", "Go9"),
                                            (selfplay.GAME_GO9_TERRITORY, "go-9", "Go9Territory"), (selfplay.GAME_CHESS, "chess", "Chess")])
def test_recorded_games_replay_under_the_oracle_rules(tmp_path, game, name, twin):
    """N1 + N2 end to end on the host (DummyNetwork stand-in); tests/test_gpu_selfplay.py does the same with games the GPU played."""
    prefix, r = _run(tmp_path, game)
    assert replay_under_oracle_rules(prefix, name, twin) >= 40


def test_joined_record_files_hold_exactly_the_parts(tmp_path):
    """kzero_b200/record_files.merge (several devices per server): the joined file is the parts one after the other -- every position's
    bytes unchanged except the game id (shifted by the games in front), offsets and per-game start indices consistent, metadata
    recombined."""
    from kzero_b200 import record_files

    parts = []
    for i, seed in enumerate((5, 6, 7)):
        prefix = str(tmp_path / f"part{i}")
        cfg = selfplay.default_config(game=selfplay.GAME_ATAXX7, visits=16, search_batch=4, gpu_batch=32, cpu_threads=2, gpu_threads=1,
                                      max_games=2 + i, duration_s=30.0, dummy_network=1, output_prefix=prefix, seed=seed)
        assert selfplay.run(None, cfg).games_written >= 2 + i
        parts.append(prefix)
    before = [_parse(p, 3 * 49, 1) for p in parts]
    out = str(tmp_path / "joined")
    meta = record_files.merge(parts, out)
    assert not any(Path(p + ext).exists() for p in parts for ext in (".bin", ".off", ".json"))
    jmeta, positions, starts = _parse(out, 3 * 49, 1)
    assert jmeta == meta
    assert jmeta["game_count"] == sum(m["game_count"] for m, _, _ in before) == len(starts)
    assert jmeta["position_count"] == sum(m["position_count"] for m, _, _ in before) == len(positions)
    assert jmeta["max_game_length"] == max(m["max_game_length"] for m, _, _ in before)
    assert jmeta["min_game_length"] == min(m["min_game_length"] for m, _, _ in before)
    total = jmeta["game_count"]
    for i in range(3):
        want = sum(m["root_wdl"][i] * m["game_count"] for m, _, _ in before) / total
        assert abs(jmeta["root_wdl"][i] - want) < 1e-9
    pos_base = game_base = 0
    for m, part_positions, part_starts in before:
        for k, p in enumerate(part_positions):
            q = positions[pos_base + k]
            assert int(q["scalars"][0]) == int(p["scalars"][0]) + game_base
            assert np.array_equal(q["scalars"][1:], p["scalars"][1:], equal_nan=True)
            assert np.array_equal(q["bits"], p["bits"]) and np.array_equal(q["input_scalars"], p["input_scalars"])
            assert np.array_equal(q["indices"], p["indices"]) and np.array_equal(q["values"], p["values"])
        assert [int(s) for s in starts[game_base:game_base + m["game_count"]]] == [int(s) + pos_base for s in part_starts]
        pos_base += m["position_count"]
        game_base += m["game_count"]
    with pytest.raises(ValueError, match="disagree"):
        a, b = str(tmp_path / "a"), str(tmp_path / "b")
        for prefix, game in ((a, selfplay.GAME_ATAXX7), (b, selfplay.GAME_SYNTH_CHESS)):
            cfg = selfplay.default_config(game=game, visits=8, search_batch=4, gpu_batch=32, cpu_threads=1, gpu_threads=1, max_games=1,
                                          duration_s=30.0, dummy_network=1, output_prefix=prefix, seed=3)
            selfplay.run(None, cfg)
        record_files.merge([a, b], str(tmp_path / "bad"))
