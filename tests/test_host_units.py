"""C++ unit checks of the self-play host code, compiled with g++ and run here -- no GPU.

  tests/cpp/lru_cache_test.cpp    the flat LRU evaluation cache against a std::list + std::unordered_map model
  tests/cpp/mcts_units_test.cpp   the vectorised uct / tie-aware argmax / visited lists against their scalar definitions
                                  (node.rs:163-206, kz-util/src/sequence.rs:11-41)
"""
import subprocess
from pathlib import Path

import pytest

from kzero_b200 import selfplay

ROOT = Path(__file__).resolve().parent


@pytest.mark.parametrize("name", ["lru_cache_test", "mcts_units_test"])
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
