#!/bin/bash
# Self-play scaling on one 8-GPU box: N = 1, 2, 4, 8 replicas (one process per GPU, games sharded by replica).
mkdir -p gpurun_out
lscpu | egrep "Model name|^CPU\(s\)|L2|L3|Thread" > gpurun_out/sp_n8_cpu.txt
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    timeout 180 python scripts/selfplay_bench.py --seconds 8 2>/dev/null | tail -1 > gpurun_out/sp_scale_chess_n$n.json
  else
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29610 + n)) \
      scripts/selfplay_bench.py --seconds 8 2>/dev/null | grep '"metric"' | tail -1 > gpurun_out/sp_scale_chess_n$n.json
  fi
  python - <<PY
import json
d = json.load(open("gpurun_out/sp_scale_chess_n$n.json"))
print("N=$n nodes/s %.0f nn/s %.0f mean_batch %.0f cfg %s" % (d["value"], d["nn_positions_per_s"], d["mean_batch"], {k: d["config"][k] for k in ("cpu_threads_per_gpu", "gpu_threads_per_gpu", "executor_blocking_sync", "host_cores")}))
PY
done
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 \
  scripts/selfplay_bench.py --seconds 6 --game ataxx 2>/dev/null | grep '"metric"' | tail -1 > gpurun_out/sp_scale_ataxx_n8.json; cut -c1-200 gpurun_out/sp_scale_ataxx_n8.json
