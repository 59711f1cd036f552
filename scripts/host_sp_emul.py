#!/usr/bin/env python
"""Host-only emulation of one self-play replica: the whole driver with the pseudo-network and an emulated GPU (one batch
at a time, KZB_SP_DUMMY_LATENCY_US per batch).  Start one per group of cores with taskset to emulate an N-GPU host.

    KZB_SP_DUMMY_LATENCY_US=500 KZB_SP_DUMMY_SERIAL=1 [GAME=chess] taskset -c 0-3 python scripts/host_sp_emul.py [cpu_threads] [gpu_threads] [games] [seconds]
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kzero_b200 import selfplay  # noqa: E402

os.environ.setdefault("KZB_SP_DUMMY_LATENCY_US", "500")
os.environ.setdefault("KZB_SP_DUMMY_SERIAL", "1")
cpu = int(sys.argv[1]) if len(sys.argv) > 1 else 4
gpu = int(sys.argv[2]) if len(sys.argv) > 2 else 3
games = int(sys.argv[3]) if len(sys.argv) > 3 else 0
seconds = float(sys.argv[4]) if len(sys.argv) > 4 else 5.0
game = {"synth": selfplay.GAME_SYNTH_CHESS, "chess": selfplay.GAME_CHESS}[os.environ.get("GAME", "synth")]
cfg = selfplay.default_config(game=game, visits=800, search_batch=16, gpu_batch=1024, cpu_threads=cpu, gpu_threads=gpu,
                              concurrent_games=games, duration_s=seconds, seed=int(os.environ.get("SEED", "1")), dummy_network=2,
                              executor_blocking_sync=1)
r = selfplay.run(None, cfg, device=0)
n = r.real_evals + r.cached_evals
print(f"nodes/s {n / r.seconds:.0f} nn/s {r.real_evals / r.seconds:.0f} mean_batch {r.real_evals / max(r.batches, 1):.0f}")
