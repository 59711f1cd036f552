"""kzb_selfplay_session_*: the games of a server connection live BETWEEN generations, like the reference's generators, which run across
file boundaries while the collector only rotates the output file (rust/kz-selfplay/src/server/collector.rs:59-116).  A generation
that ends must not drop the games still in flight (dropping them biases the data towards short games).  No GPU
(DummyNetwork stand-in)."""
import json
from pathlib import Path

import pytest

from kzero_b200 import selfplay
from test_selfplay_records import _parse


@pytest.fixture(autouse=True)
def _fresh_stop_flag():
    """The process-wide stop flag stays set until it is cleared (a server that was stopped earlier in this process leaves it set)."""
    from kzero_b200 import _abi

    _abi.lib().kzb_selfplay_clear_stop()
    _abi.lib().kzb_selfplay_clear_interrupt()
    yield
    _abi.lib().kzb_selfplay_clear_stop()
    _abi.lib().kzb_selfplay_clear_interrupt()


def _cfg(prefix, **kw):
    base = dict(game=selfplay.GAME_SYNTH_CHESS, visits=16, search_batch=4, gpu_batch=32, cpu_threads=2, gpu_threads=1, concurrent_games=8,
                max_games=1, duration_s=120.0, dummy_network=2, output_prefix=prefix, seed=11)
    base.update(kw)
    return selfplay.default_config(**base)


def test_games_in_flight_continue_in_the_next_generation(tmp_path):
    moves, lengths = [], []
    with selfplay.Session(selfplay.GAME_SYNTH_CHESS) as session:
        for gen in range(8):
            prefix = str(tmp_path / f"games_{gen}")
            r = session.run(None, _cfg(prefix))
            assert r.games_written >= 1 and r.concurrent_games == 8
            meta, positions, starts = _parse(prefix, 13 * 64, 8)
            assert meta["game_count"] == r.games_written
            lengths.append([int(positions[int(s)]["scalars"][2]) for s in starts])
            moves.append(r.moves_played)
            # every file is complete on its own: per-game position indices run 0..length
            for s, length in zip(starts, lengths[-1]):
                for k in range(length + 1):
                    assert int(positions[int(s) + k]["scalars"][1]) == k
    # a generation cannot have played the games it wrote from their start to their end if the whole generation -- all 8 games
    # together -- played fewer moves than those games are long: such games were carried over from earlier generations
    carried = [gen for gen in range(1, len(moves)) if sum(lengths[gen]) > moves[gen]]
    assert len(carried) >= 2, (moves, lengths)
    # and nothing is invented: the files never hold more positions than were played
    assert sum(sum(lens) for lens in lengths) <= sum(moves)


def test_a_session_keeps_its_startup_settings(tmp_path):
    with selfplay.Session(selfplay.GAME_SYNTH_CHESS) as session:
        session.run(None, _cfg(str(tmp_path / "a")))
        with pytest.raises(Exception, match="cpu_threads"):
            session.run(None, _cfg(str(tmp_path / "b"), cpu_threads=3))
        with pytest.raises(Exception, match="game"):
            session.run(None, _cfg(str(tmp_path / "c"), game=selfplay.GAME_ATAXX7))
        r = session.run(None, _cfg(str(tmp_path / "d"), visits=8))  # Settings may change between generations
        assert r.games_written >= 1


def test_stop_requested_before_a_run_is_not_lost(tmp_path):
    """a Stop that arrives while the network is still being built must end the run; only an explicit clear resets it."""
    from kzero_b200 import _abi

    with selfplay.Session(selfplay.GAME_SYNTH_CHESS) as session:
        _abi.lib().kzb_selfplay_request_stop()
        r = session.run(None, _cfg(str(tmp_path / "a"), max_games=1000, duration_s=60.0))
        assert r.seconds < 5.0 and r.games_written < 1000
        _abi.lib().kzb_selfplay_clear_stop()
        r = session.run(None, _cfg(str(tmp_path / "b")))
        assert r.games_written >= 1


def test_an_interrupt_keeps_the_record_file_open_for_the_next_run(tmp_path):
    """kzb_selfplay_request_interrupt: how a network that arrives in the middle of a generation is put to work at once (the reference's
    executors swap it in between batches, executor.rs:50-65,320-342).  The interrupted run returns with its record file still open in
    the session; the next run -- here with a different stand-in network -- continues the SAME games into the SAME file, which is
    complete only after that run, and holds positions evaluated by both networks, some of them in one and the same game."""
    import threading
    import time

    import numpy as np

    from kzero_b200 import _abi

    lib = _abi.lib()
    lib.kzb_selfplay_clear_interrupt()
    prefix = str(tmp_path / "games_0")
    with selfplay.Session(selfplay.GAME_SYNTH_CHESS) as session:
        timer = threading.Timer(0.4, lib.kzb_selfplay_request_interrupt)
        timer.start()
        t0 = time.perf_counter()
        first = session.run(None, _cfg(prefix, max_games=100000, dummy_network=1))  # DummyNetwork: uniform policy, value 0
        timer.join()
        assert first.interrupted == 1 and time.perf_counter() - t0 < 10.0
        assert first.games_written < 100000 and not Path(prefix + ".json").exists()  # no metadata yet: the file is not finished
        # sticky, like the stop flag: a run that starts while it is set returns at once, its file still open
        again = session.run(None, _cfg(prefix, max_games=100000, dummy_network=1))
        assert again.interrupted == 1 and again.games_written >= first.games_written
        lib.kzb_selfplay_clear_interrupt()
        target = again.games_written + 12
        second = session.run(None, _cfg(prefix, max_games=target, dummy_network=2))  # the pseudo-network: sharp answers
        assert second.interrupted == 0 and second.games_written >= target
    meta, positions, starts = _parse(prefix, 13 * 64, 8)
    assert meta["game_count"] == second.games_written == len(starts)
    uniform = []  # per position (final positions excluded): was its root evaluated by the uniform DummyNetwork?
    game_of = []
    for g, s in enumerate(starts):
        length = int(positions[int(s)]["scalars"][2])
        for k in range(length):
            sc = positions[int(s) + k]["scalars"]
            uniform.append(abs(float(sc[21])) < 1e-9 and abs(float(sc[22]) - 1 / 3) < 1e-6)  # net_v, net_wdl_w
            game_of.append(g)
    uniform, game_of = np.array(uniform), np.array(game_of)
    assert uniform.any() and (~uniform).any()
    mixed = [g for g in range(len(starts)) if uniform[game_of == g].any() and (~uniform[game_of == g]).any()]
    assert mixed, "no game was played under both networks"
    for g in mixed:  # within a game the swap happens once: first the old network's positions, then the new one's
        u = uniform[game_of == g]
        assert not u[int(np.argmin(u)):].any()


def test_a_plain_run_ignores_the_interrupt_flag(tmp_path):
    from kzero_b200 import _abi

    _abi.lib().kzb_selfplay_request_interrupt()
    try:
        r = selfplay.run(None, _cfg(str(tmp_path / "a"), max_games=3))
        assert r.interrupted == 0 and r.games_written >= 3 and Path(str(tmp_path / "a") + ".json").exists()
    finally:
        _abi.lib().kzb_selfplay_clear_interrupt()
