"""C++ unit checks of the self-play host code, compiled with g++ and run here -- no GPU.

  tests/cpp/lru_cache_test.cpp    the flat LRU evaluation cache against a std::list + std::unordered_map model
  tests/cpp/mcts_units_test.cpp   the vectorised uct / tie-aware argmax / visited blocks against their scalar definitions
                                  (node.rs:163-206, kz-util/src/sequence.rs:11-41)
  tests/cpp/go_rules_test.cpp     the 9x9 go restatement: captures, suicide, ko, passes, area scoring, encoding, random playouts
  tests/cpp/chess_perft_test.cpp  the chess move generator: perft counts of five standard positions through the policy-index interface
  tests/cpp/selfplay_tsan_main.cpp  the generator / executor threads of the driver under ThreadSanitizer
"""
import os
import json
import subprocess
from pathlib import Path

import pytest

from kzero_b200 import selfplay

ROOT = Path(__file__).resolve().parent


@pytest.mark.parametrize("name", ["lru_cache_test", "mcts_units_test", "go_rules_test", "chess_perft_test"])
def test_cpp_unit(tmp_path, name):
    exe = tmp_path / name
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), str(ROOT / "cpp" / f"{name}.cpp")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().startswith("ok"), out.stdout + out.stderr


def test_pseudo_network_run_is_sane():
    """dummy_network = 2: sharp, position-dependent answers without a GPU; full batches come from the executor policy
    (a partial batch only when no executor has work in flight)."""
    cfg = selfplay.default_config(game=selfplay.GAME_SYNTH_CHESS, visits=60, search_batch=8, gpu_batch=128, cpu_threads=2,
                                  gpu_threads=2, max_moves=150, duration_s=30.0, dummy_network=2, seed=3)
    r = selfplay.run(None, cfg, device=0)
    assert r.moves_played >= 150 and r.real_evals > 0 and r.batches > 0
    assert r.max_batch <= 128 and r.real_evals <= r.potential_evals
    assert 0.0 < r.cached_evals / (r.real_evals + r.cached_evals) < 0.9  # sharp policies revisit positions: the cache gets hits
    assert abs(r.root_visits / r.moves_played - 60) < 8 + 1  # every move was searched to ~`visits` (overshoot < search_batch)


@pytest.mark.parametrize("env", [{}, {"KZB_SP_SPIN_US": "50", "KZB_SP_PIN_GENERATORS": "1"}])
def test_selfplay_threads_under_tsan(tmp_path, env):
    """The driver's queue / wake-up / record-writer synchronisation, with the evaluator stubbed out (no CUDA linked)."""
    cuda_include = "/usr/local/cuda/include"
    if not os.path.isdir(cuda_include):
        pytest.skip("CUDA headers not found")
    exe = tmp_path / "selfplay_tsan"
    build = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-I", cuda_include, "-o", str(exe),
                            str(ROOT / "cpp" / "selfplay_tsan_main.cpp"), "-lpthread"], capture_output=True, text=True)
    if build.returncode != 0 and "tsan" in build.stderr.lower():
        pytest.skip("ThreadSanitizer runtime not available")
    assert build.returncode == 0, build.stderr[-2000:]
    out = subprocess.run([str(exe), str(tmp_path / "games")], capture_output=True, text=True, timeout=300, env={**os.environ, **env})
    assert out.returncode == 0 and "ThreadSanitizer" not in out.stderr, out.stdout + out.stderr[-4000:]


def test_chess_policy_table_matches_the_reference(tmp_path):
    """Row A7 for chess on the C++ side: the 1880-entry flat move table of chess_game.hpp against the reference's own
    python/lib/mapping/chess_flat_to_move_input.txt (tests/golden/chess_flat_moves.json, gen_chess_moves_golden.py)."""
    exe = tmp_path / "chess_flat_dump"
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", str(exe), str(ROOT / "cpp" / "chess_flat_dump.cpp")], check=True)
    lines = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")[:-1]
    mine = [[int(a), int(b), "" if c == "-" else c] for a, b, c in (line.split() for line in lines)]
    golden = json.loads((ROOT / "golden" / "chess_flat_moves.json").read_text())
    assert len(mine) == 1880 and mine == golden
