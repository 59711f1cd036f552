"""The TCP / JSON control protocol (row N3) exercised with the REFERENCE'S OWN client class
(python/lib/selfplay_client.py, imported from /root/reference when present): StartupSettings, NewSettings,
UseDummyNetwork, two FinishedFile notifications, Stop -- and the files the run leaves behind are read back with the
reference's DataFile loader.  DummyNetwork evaluations: no GPU needed."""
import json
import socket
import sys
import threading
from pathlib import Path

import pytest

from kzero_b200 import selfplay_server

REFERENCE_PY = Path("/root/reference/python")

STARTUP = dict(game="ataxx-7", muzero=False, start_pos="default", first_gen=3, output_folder=None, games_per_gen=6,
               cpu_threads_per_device=2, gpu_threads_per_device=1, gpu_batch_size=32, gpu_batch_size_root=0, search_batch_size=4,
               saved_state_channels=0, eval_random_symmetries=True)
SETTINGS = dict(max_game_length=60, weights=dict(exploration_weight=None, moves_left_weight=None, moves_left_clip=None,
                                                 moves_left_sharpness=None),
                q_mode="wdl+0.0", temperature=1.0, zero_temp_move_count=30, dirichlet_alpha=0.2, dirichlet_eps=0.25,
                search_policy_temperature_root=1.4, search_policy_temperature_child=1.0, search_fpu_root="fixed+0.1",
                search_fpu_child="relative+0", search_virtual_loss_weight=1.0, full_search_prob=0.5, full_iterations=24,
                part_iterations=6, top_moves=100, cache_size=100)


def test_mode_strings_parse_like_the_reference():
    assert selfplay_server.parse_fpu("fixed+0.1") == (0, 0.1) and selfplay_server.parse_fpu("relative-0.25") == (1, -0.25)
    assert selfplay_server.parse_q_mode("value") == (0, 0.0) and selfplay_server.parse_q_mode("wdl") == (1, 0.0)
    assert selfplay_server.parse_q_mode("wdl-0.5") == (1, -0.5)  # step.rs:282-303 round trips
    with pytest.raises(ValueError):
        selfplay_server.parse_fpu("nonsense")


def _serve():
    server = selfplay_server.SelfplayServer(port=0)
    t = threading.Thread(target=server.serve, daemon=True)
    t.start()
    return server, t


@pytest.mark.parametrize("game,length_cap", [("ataxx-7", 60), ("go-9", 30), ("chess", 30)])
def test_protocol_with_raw_socket(tmp_path, game, length_cap):
    server, thread = _serve()
    s = socket.create_connection(("127.0.0.1", server.port))
    f = s.makefile("r")

    def send(m):
        s.sendall((json.dumps(m) + "\n").encode())

    send({"StartupSettings": dict(STARTUP, game=game, output_folder=str(tmp_path))})
    send({"NewSettings": dict(SETTINGS, max_game_length=length_cap)})
    send("UseDummyNetwork")
    assert json.loads(f.readline()) == {"FinishedFile": {"index": 3}}
    assert json.loads(f.readline()) == {"FinishedFile": {"index": 4}}
    send("Stop")
    lines = [json.loads(line) for line in f]
    assert lines[-1] == "Stopped"
    thread.join(timeout=30)
    assert not thread.is_alive()
    for gen in (3, 4):
        meta = json.loads((tmp_path / f"games_{gen}.json").read_text())
        assert meta["game"] == game and meta["game_count"] >= 6 and meta["max_game_length"] <= length_cap


@pytest.mark.skipif(not REFERENCE_PY.exists(), reason="the reference tree is only present in the build container")
def test_reference_client_drives_the_server(tmp_path):
    sys.path.insert(0, str(REFERENCE_PY))
    try:
        from lib.data.file import DataFile
        from lib.games import Game
        from lib.selfplay_client import SelfplayClient, SelfplaySettings, StartupSettings, UctWeights
    finally:
        sys.path.pop(0)
    server, thread = _serve()
    client = SelfplayClient(server.port)
    client.send_startup_settings(StartupSettings(**dict(STARTUP, output_folder=str(tmp_path), first_gen=0)))
    client.send_new_settings(SelfplaySettings(**dict(SETTINGS, weights=UctWeights.default())))
    client.send_dummy_network()
    assert client.wait_for_file() == 0
    assert client.wait_for_file() == 1
    client.send_stop()
    with pytest.raises(RuntimeError, match="stopped"):
        while True:
            client.wait_for_file()
    thread.join(timeout=30)
    f = DataFile.open(Game.find("ataxx-7"), str(tmp_path / "games_0"))
    assert f.info.simulation_count >= 6
    positions = [f.load_position(i) for i in range(f.info.position_count)]
    # full_search_prob = 0.5: both kinds of searches were recorded, with their visit targets (generator_alphazero.rs:88-94)
    kinds = {(p.is_full_search, p.zero_visits >= 24) for p in positions if not p.is_final}
    assert (True, True) in kinds and any(not full for full, _ in kinds)


def test_several_devices_write_one_file_per_generation(tmp_path):
    """`--device a --device b` (server.rs:49-51,316-331): one session per device, the generation's games split between them, the parts
    joined into the single games_<gen> file the loop expects.  DummyNetwork evaluations, so both 'devices' are host threads here; the
    joined file must pass the same checks as a file written by one writer (tests/test_selfplay_records.py) and, when the reference tree
    is present, load with the reference's DataFile."""
    import numpy as np

    from test_selfplay_records import SCALARS, _parse

    server = selfplay_server.SelfplayServer(port=0, devices=[0, 1])
    thread = threading.Thread(target=server.serve, daemon=True)
    thread.start()
    s = socket.create_connection(("127.0.0.1", server.port))
    f = s.makefile("r")
    for m in ({"StartupSettings": dict(STARTUP, games_per_gen=7, output_folder=str(tmp_path))}, {"NewSettings": SETTINGS}, "UseDummyNetwork"):
        s.sendall((json.dumps(m) + "\n").encode())
    assert json.loads(f.readline()) == {"FinishedFile": {"index": 3}}
    assert json.loads(f.readline()) == {"FinishedFile": {"index": 4}}
    s.sendall(b'"Stop"\n')
    assert [json.loads(line) for line in f][-1] == "Stopped"
    thread.join(timeout=30)
    assert not thread.is_alive()
    for gen in (3, 4):
        prefix = str(tmp_path / f"games_{gen}")
        assert not list(tmp_path.glob(f"games_{gen}.dev*"))  # the parts are gone
        meta, positions, starts = _parse(prefix, 3 * 49, 1)
        # (a device may finish a few games past its share -- at most its games in flight: generator threads end games concurrently)
        assert 7 <= meta["game_count"] <= 7 + 2 * 8 and len(starts) == meta["game_count"] and len(meta["scalar_names"]) == SCALARS
        assert abs(sum(meta["root_wdl"]) - 1) < 1e-6
        # python/lib/data/check.py:9-76 on the joined file: games tile the positions, ids count up, per-game indices run 0..length
        pi = 0
        for g, start in enumerate(starts):
            assert int(start) == pi
            length = int(positions[pi]["scalars"][2])
            for k in range(length + 1):
                sc = positions[pi + k]["scalars"]
                assert int(sc[0]) == g and int(sc[1]) == k and int(sc[2]) == length and bool(sc[5]) == (k == length)
            pi += length + 1
        assert pi == meta["position_count"]
        lengths = [int(positions[int(st)]["scalars"][2]) for st in starts]
        assert meta["max_game_length"] == max(lengths) and meta["min_game_length"] == min(lengths)
    if REFERENCE_PY.exists():
        sys.path.insert(0, str(REFERENCE_PY))
        try:
            from lib.data.file import DataFile
            from lib.games import Game
        finally:
            sys.path.pop(0)
        df = DataFile.open(Game.find("ataxx-7"), str(tmp_path / "games_3"))
        assert df.info.simulation_count >= 7
        assert len([df.load_position(i) for i in range(df.info.position_count)]) == df.info.position_count
    assert np.isfinite(meta["hit_move_limit"])


def test_new_settings_take_effect_in_the_middle_of_a_file(tmp_path):
    """The reference applies new settings / a new network as they arrive (commander.rs:13-61, executor.rs:50-65), not at the next file:
    NewSettings sent while generation 3 is being played shows up INSIDE games_3 (both visit targets occur in it), the file still holds
    games_per_gen games, and the FinishedFile sequence is unbroken."""
    from test_selfplay_records import _parse

    server, thread = _serve()
    s = socket.create_connection(("127.0.0.1", server.port))
    f = s.makefile("r")

    def send(m):
        s.sendall((json.dumps(m) + "\n").encode())

    first = dict(SETTINGS, full_search_prob=1.0, full_iterations=24)
    send({"StartupSettings": dict(STARTUP, games_per_gen=300, output_folder=str(tmp_path))})
    send({"NewSettings": first})
    send("UseDummyNetwork")
    import time

    time.sleep(0.3)  # two generations of 300 ataxx games take over a second at these settings: the change arrives inside one of them
    send({"NewSettings": dict(first, full_iterations=40)})
    assert json.loads(f.readline()) == {"FinishedFile": {"index": 3}}
    assert json.loads(f.readline()) == {"FinishedFile": {"index": 4}}
    send("Stop")
    assert [json.loads(line) for line in f][-1] == "Stopped"
    thread.join(timeout=30)
    assert not thread.is_alive()
    visits = {}
    for gen in (3, 4):
        meta, positions, starts = _parse(str(tmp_path / f"games_{gen}"), 3 * 49, 1)
        assert meta["game_count"] >= 300 and len(starts) == meta["game_count"]
        targets = set()
        for p in positions:
            if not bool(p["scalars"][5]):  # not the final position of a game
                v = int(p["scalars"][3])   # zero_visits: the search ran to its target, overshooting by less than a search batch
                targets.add(24 if v < 36 else 40)
        visits[gen] = targets
    assert {24, 40} in visits.values(), visits  # the change arrived inside a file ...
    assert visits[4] == {40} or visits[3] == {24}, visits  # ... and everything after it uses the new target, everything before the old one
