#!/bin/bash
# round 2: the self-play loop at N = 8 (32 host cores, 4 per GPU): what moves the last few percent to 7x of one GPU?
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {  # label, env assignments, extra args
  local label="$1"; local envs="$2"; shift 2
  echo -n "$label: "
  env $envs timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
      scripts/selfplay_bench.py --seconds 5 "$@" 2> gpurun_out/r02_n8_sp_err.txt | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'nn', round(d['nn_positions_per_s']), 'batch', round(d['mean_batch']), 'games/gpu', d['config']['concurrent_games_per_gpu'], 'threads', d['config']['cpu_threads_per_gpu'], d['config']['gpu_threads_per_gpu'], d['config']['executor_blocking_sync'])
except Exception as e:
    print('failed', repr(e))"
}
{
run "A default (tree reuse)" "A=1"
run "B malloc THP" "GLIBC_TUNABLES=glibc.malloc.hugetlb=1"
run "C THP + 384 games" "GLIBC_TUNABLES=glibc.malloc.hugetlb=1" --concurrent-games 384
run "D THP + generator spin 30us" "GLIBC_TUNABLES=glibc.malloc.hugetlb=1 KZB_SP_SPIN_US=30"
run "E THP + 2 executors" "GLIBC_TUNABLES=glibc.malloc.hugetlb=1" --gpu-threads 2
run "F THP + 4 executors" "GLIBC_TUNABLES=glibc.malloc.hugetlb=1" --gpu-threads 4
run "G real chess default" "A=1" --game chess-real
run "H real chess THP" "GLIBC_TUNABLES=glibc.malloc.hugetlb=1" --game chess-real
echo -n "N=1 same box: "; timeout 100 python scripts/selfplay_bench.py --seconds 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'nn', round(d['nn_positions_per_s']))"
grep -i hugepages /proc/meminfo | head -3
} | tee gpurun_out/r02_n8_selfplay_ab.txt
