"""Self-play driver on the GPU: generator threads + executor threads feeding the B200 evaluator with ragged batches."""
import pytest

from kzero_b200 import netgen, selfplay

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("game,game_name,gpu_threads", [(selfplay.GAME_SYNTH_CHESS, "chess", 1), (selfplay.GAME_ATAXX7, "ataxx-7", 2)])
def test_selfplay_runs_and_counts_are_consistent(game, game_name, gpu_threads):
    spec = netgen.game_spec(game_name)
    onnx_bytes = netgen.build_onnx(spec, 2, 32, seed=31)
    cfg = selfplay.default_config(game=game, visits=60, search_batch=8, gpu_batch=128, cpu_threads=2, gpu_threads=gpu_threads,
                                  duration_s=1.5, seed=7)
    r = selfplay.run(onnx_bytes, cfg)
    assert r.real_evals > 0 and r.batches > 0 and r.moves_played > 0
    assert r.max_batch <= 128 and r.mean_batch >= 1
    assert r.potential_evals == r.batches * 128  # collector.rs:172-191 "potential"
    # every finished search reached its visit target; each visit is a real eval, a cache hit or a terminal gather
    assert r.root_visits >= r.moves_played * 60
    assert r.concurrent_games == (gpu_threads + 1) * 128 // 8  # server_alphazero.rs:47
    assert r.mcts_nodes_per_s >= r.nn_positions_per_s > 0


@pytest.mark.parametrize("game,name,twin", [(selfplay.GAME_ATAXX7, "ataxx-7", "Ataxx7"), (selfplay.GAME_GO9, "go-9", "Go9"),
                                            (selfplay.GAME_GO9_TERRITORY, "go-9", "Go9Territory"), (selfplay.GAME_CHESS, "chess", "Chess")])
def test_games_the_gpu_played_replay_under_the_oracle_rules(tmp_path, game, name, twin):
    """What the GPU driver PLAYED, not only how much: every recorded game -- searched with a real network on the B200 evaluator,
    ragged batches, several executors -- is replayed under the oracle's independent restatement of the rules (legal move lists,
    encodings, transitions, outcomes), and the file is read by the same checks as the host-only records."""
    from test_selfplay_records import replay_under_oracle_rules

    spec = netgen.game_spec("go-9-territory" if twin == "Go9Territory" else name)  # 13 input channels: GoStdMapper::new(9, true)
    onnx_bytes = netgen.build_onnx(spec, 2, 32, seed=41)
    prefix = str(tmp_path / "games_0")
    cfg = selfplay.default_config(game=game, visits=40, search_batch=8, gpu_batch=64, cpu_threads=2, gpu_threads=2, max_games=6,
                                  max_game_length=30 if game != selfplay.GAME_ATAXX7 else 400, duration_s=60.0, output_prefix=prefix, seed=9)
    r = selfplay.run(onnx_bytes, cfg)
    assert r.games_written > 0 and r.real_evals > 0 and r.batches > 0
    assert replay_under_oracle_rules(prefix, name, twin) >= 40


def test_selfplay_rejects_mismatched_network():
    from kzero_b200.network import KzbError

    onnx_bytes = netgen.build_onnx(netgen.game_spec("ataxx-7"), 1, 16, seed=32)
    cfg = selfplay.default_config(game=selfplay.GAME_SYNTH_CHESS, duration_s=0.2)
    with pytest.raises(KzbError, match="Input shape mismatch"):  # check_graph_shapes, network/common.rs:171-174
        selfplay.run(onnx_bytes, cfg)


def test_server_hot_path_with_a_real_network(tmp_path):
    """N3 end to end on the GPU: StartupSettings -> NewSettings -> NewNetwork(path) -> FinishedFile -> Stop."""
    import json
    import socket
    import threading

    from kzero_b200 import selfplay_server
    from test_selfplay_server import SETTINGS, STARTUP

    onnx_path = tmp_path / "net.onnx"
    onnx_path.write_bytes(netgen.build_onnx(netgen.game_spec("ataxx-7"), 2, 32, seed=33))
    server = selfplay_server.SelfplayServer(port=0)
    thread = threading.Thread(target=server.serve, daemon=True)
    thread.start()
    s = socket.create_connection(("127.0.0.1", server.port))
    f = s.makefile("r")
    for m in ({"StartupSettings": dict(STARTUP, output_folder=str(tmp_path), first_gen=0)}, {"NewSettings": SETTINGS},
              {"NewNetwork": str(onnx_path)}):
        s.sendall((json.dumps(m) + "\n").encode())
    assert json.loads(f.readline()) == {"FinishedFile": {"index": 0}}
    s.sendall(b'"Stop"\n')
    assert [json.loads(line) for line in f][-1] == "Stopped"
    thread.join(timeout=60)
    meta = json.loads((tmp_path / "games_0.json").read_text())
    assert meta["game_count"] >= 6 and meta["position_count"] > meta["game_count"]
