#!/bin/bash
# 4-core (taskset) self-play sweep over the topology knobs: generator threads, executor threads, games in flight.
mkdir -p gpurun_out
: > gpurun_out/sp_sweep_4cores.jsonl
for cfg in "4 2 0" "4 3 0" "4 4 0" "3 3 0" "4 3 384"; do
  set -- $cfg
  echo "cpu_threads=$1 gpu_threads=$2 concurrent_games=$3"
  timeout 120 taskset -c 0-3 python scripts/selfplay_bench.py --seconds 6 --cpu-threads $1 --gpu-threads $2 --concurrent-games $3 2>/dev/null | tail -1 | tee -a gpurun_out/sp_sweep_4cores.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   nodes/s %.0f  nn/s %.0f  mean_batch %.0f' % (d['value'], d['nn_positions_per_s'], d['mean_batch']))"
done
echo "all cores, 2 and 3 executors"
for g in 2 3; do timeout 120 python scripts/selfplay_bench.py --seconds 6 --gpu-threads $g 2>/dev/null | tail -1 | tee -a gpurun_out/sp_sweep_allcores.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   nodes/s %.0f  nn/s %.0f  mean_batch %.0f' % (d['value'], d['nn_positions_per_s'], d['mean_batch']))"; done
