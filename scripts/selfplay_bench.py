#!/usr/bin/env python
"""Self-play MCTS nodes/sec on B200 (BASELINE.json configs[3] and the second half of its metric).

    python scripts/selfplay_bench.py [--seconds S] [--visits 800] [--cpu-threads T] [--gpu-threads G] [--game chess|ataxx]
    torchrun --nproc-per-node N scripts/selfplay_bench.py ...      # one replica per GPU, games sharded by replica

Settings default to the reference's production values (python/main/loop_main_alpha.py:24-52: 800 visits, search batch 16,
virtual loss 1, LRU cache 800, Dirichlet 0.03/0.25, root temperature 1.4); the TOPOLOGY is this repo's: gpu batch 1024 (BASELINE.json's
batch; the reference runs 2048), three executor threads and every remaining core as a generator thread (the reference: 1 and 4,
loop_main_alpha.py:24-26) -- all of them are startup settings on both sides.  The net is chess 16x128 random-init
(synthetic, like bench.py).  The game is the chess-SHAPED synthetic game of kzero_b200/csrc/selfplay/games.hpp (--game chess), real chess
(--game chess-real, chess_game.hpp), real 7x7 ataxx with an 8x64 net, or 9x9 go with the 20x256 net.  Prints one JSON line on rank 0:
  nodes/s = (real + cached evals) / s  (the collector's `evals/s: real / cached`, collector.rs:172-191), NN positions/s,
  mean batch and fill of the evaluator calls.
"""
import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kzero_b200 import netgen, replicas, selfplay  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--game", default="chess", choices=["chess", "chess-real", "ataxx", "go"])
ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--visits", type=int, default=800)
ap.add_argument("--search-batch", type=int, default=16)
ap.add_argument("--gpu-batch", type=int, default=1024)
ap.add_argument("--cpu-threads", type=int, default=0, help="0 = host cores / replicas")
ap.add_argument("--gpu-threads", type=int, default=3, help="executor threads, one network instance each")
ap.add_argument("--concurrent-games", type=int, default=0)
ap.add_argument("--pin", action="store_true", help="pin each replica to its own contiguous share of the host cores")
ap.add_argument("--warmup-seconds", type=float, default=0.0, help="play this long on the same games before the timed run (a session keeps them)")
args = ap.parse_args()

ctx = replicas.context_from_env()
dist = None
if ctx.world > 1:
    import torch

    torch.cuda.set_device(ctx.local_rank)
    dist = replicas.init_process_group(ctx, "nccl", torch.device("cuda", ctx.local_rank))
cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)  # honours taskset
share = max(1, cores // ctx.world)  # host cores of this replica
if args.pin and ctx.world > 1 and hasattr(os, "sched_setaffinity"):
    mine = sorted(os.sched_getaffinity(0))[ctx.local_rank * share:(ctx.local_rank + 1) * share]
    os.sched_setaffinity(0, mine)  # threads started from here on inherit it
# plenty of cores: executors spin (lowest latency) on cores of their own; few cores: they sleep on a blocking event and
# every core runs a generator
blocking = share < 12
# three executors: while one scatters answers and the next assembles its batch, a third already has work queued on the
# GPU (measured: 1.62 M nodes/s with 2 executors on 4 cores, 2.03 M with 3 -- profiles/r01d_selfplay_host.md)
gpu_threads = args.gpu_threads
cpu_threads = args.cpu_threads or (share if blocking else share - gpu_threads)
if args.game == "chess":
    spec, depth, channels, game = netgen.game_spec("chess"), 16, 128, selfplay.GAME_SYNTH_CHESS
elif args.game == "chess-real":  # legal chess instead of the chess-shaped synthetic game
    spec, depth, channels, game = netgen.game_spec("chess"), 16, 128, selfplay.GAME_CHESS
elif args.game == "go":  # the net of BASELINE.json configs[2]
    spec, depth, channels, game = netgen.game_spec("go-9"), 20, 256, selfplay.GAME_GO9
else:
    spec, depth, channels, game = netgen.game_spec("ataxx-7"), 8, 64, selfplay.GAME_ATAXX7
onnx_bytes = netgen.build_onnx(spec, depth, channels, seed=0)
cfg = selfplay.default_config(game=game, visits=args.visits, search_batch=args.search_batch, gpu_batch=args.gpu_batch,
                              cpu_threads=cpu_threads, gpu_threads=gpu_threads, concurrent_games=args.concurrent_games,
                              duration_s=args.seconds, seed=replicas.game_seed(ctx, 0), executor_blocking_sync=int(blocking))
if args.warmup_seconds > 0:  # trees allocated and faulted in, caches warm, games at staggered depths: what a long-running server looks like
    cfg.duration_s = args.warmup_seconds
    with selfplay.Session(game) as session:
        session.run(onnx_bytes, cfg, device=ctx.local_rank)
        cfg.duration_s = args.seconds
        replicas.barrier(ctx)
        r = session.run(onnx_bytes, cfg, device=ctx.local_rank)
else:
    replicas.barrier(ctx)
    r = selfplay.run(onnx_bytes, cfg, device=ctx.local_rank)
counts = [r.real_evals, r.cached_evals, r.batches, r.moves_played, r.games_finished]
if dist is not None:
    import torch

    t = torch.tensor(counts, dtype=torch.float64, device="cuda")
    dist.all_reduce(t)  # sums over replicas: the games are disjoint
    counts = [float(v) for v in t.tolist()]
(seconds,) = replicas.max_over_ranks(ctx, [r.seconds], device="cuda" if dist is not None else "cpu")
if ctx.is_root:
    real, cached, batches, moves, games = counts
    print(json.dumps({
        "metric": "self-play MCTS nodes/sec", "value": (real + cached) / seconds, "unit": "nodes/s", "n_gpus": ctx.world,
        "nn_positions_per_s": real / seconds, "cache_hit_rate": cached / max(real + cached, 1),
        "mean_batch": real / max(batches, 1), "batch_fill": real / max(batches * args.gpu_batch, 1), "max_batch_rank0": r.max_batch,
        "moves_per_s": moves / seconds, "games_finished": games, "seconds": seconds, "scaling": "weak",
        "config": {"workload": f"{args.game} self-play, {args.visits} visits, search batch {args.search_batch} with virtual loss, "
                               f"net {depth}x{channels}, gpu batch {args.gpu_batch}",
                   "game": {"chess": "chess-shaped synthetic game (13x8x8 + 8 planes, 1880-move policy, 20-45 legal moves)", "ataxx": "ataxx 7x7",
                            "chess-real": "chess (legal move generation, ChessStdMapper encoding)",
                            "go": "go 9x9 (area scoring, positional superko, cgos or Tromp-Taylor suicide rule per game)"}[args.game],
                   "cpu_threads_per_gpu": cpu_threads, "gpu_threads_per_gpu": gpu_threads, "executor_blocking_sync": bool(blocking), "concurrent_games_per_gpu": r.concurrent_games,
                   "host_cores": cores, "pinned": bool(args.pin), "warmup_seconds": args.warmup_seconds, **replicas.parallelism_note(ctx)},
        "data": "synthetic"}), flush=True)
if dist is not None:
    dist.destroy_process_group()
