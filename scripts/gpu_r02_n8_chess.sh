#!/bin/bash
# round 2: the self-play loop with REAL chess (bitboard generator) at N = 8 against N = 1 on the same 8-GPU box (32 host cores: 4 per GPU)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
fmt='
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["value"]), "nn", round(d["nn_positions_per_s"]), "batch", round(d["mean_batch"]), "hit", round(d["cache_hit_rate"], 3), "games/gpu", d["config"]["concurrent_games_per_gpu"], "threads", d["config"]["cpu_threads_per_gpu"], d["config"]["gpu_threads_per_gpu"], d["config"]["executor_blocking_sync"], "cores", d["config"]["host_cores"])
except Exception as e:
    print("failed", repr(e))'
run8() {  # label, args
  local label="$1"; shift
  echo -n "N=8 $label: "
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
      scripts/selfplay_bench.py --seconds 6 "$@" 2> gpurun_out/r02_n8_chess_err.txt | python -c "$fmt"
}
{
nproc
KZB_SP_PROFILE=1 run8 "real chess, 384 games (profiled)" --game chess-real --concurrent-games 384
grep -E "gather|apply answers|CPU time" gpurun_out/r02_n8_chess_err.txt | sort | uniq -c | sort -rn | head -12
run8 "synthetic, 384 games" --game chess --concurrent-games 384
run8 "real chess, 512 games" --game chess-real --concurrent-games 512
echo -n "N=1 same box real chess (all cores): "; timeout 100 python scripts/selfplay_bench.py --seconds 6 --game chess-real 2>/dev/null | python -c "$fmt"
echo -n "N=1 same box synthetic (all cores): "; timeout 100 python scripts/selfplay_bench.py --seconds 6 --game chess 2>/dev/null | python -c "$fmt"
echo -n "N=1 same box real chess (4 cores, 384 games): "; timeout 100 taskset -c 0-3 python scripts/selfplay_bench.py --seconds 6 --game chess-real --concurrent-games 384 2>/dev/null | python -c "$fmt"
} | tee gpurun_out/r02_n8_chess_b.txt
tail -3 gpurun_out/r02_n8_chess_err.txt
