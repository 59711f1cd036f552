#!/bin/bash
# round 2, last 8-GPU visit: the loop at N = 8 measured after a warm-up on the same games (what bench.py does now), 512 games in flight per GPU
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
fmt='
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"]), "nn", round(d["nn_positions_per_s"]), "batch", round(d["mean_batch"]), "hit", round(d["cache_hit_rate"], 3), "games/gpu", d["config"]["concurrent_games_per_gpu"], "warm-up", d["config"]["warmup_seconds"])
except Exception as e:
    print("failed", repr(e))'
{
for game in chess-real; do
  echo -n "N=8 $game, 512 games, 3 s warm-up: "
  timeout 50 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
      scripts/selfplay_bench.py --seconds 6 --warmup-seconds 3 --game $game --concurrent-games 512 2> gpurun_out/r02_n8_warm_err.txt | python -c "$fmt"
done
} | tee gpurun_out/r02_n8_warm.txt
