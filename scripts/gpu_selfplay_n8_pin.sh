#!/bin/bash
# 8 replicas, with and without pinning each replica to its own 4 cores; generator profile of every rank on stderr.
mkdir -p gpurun_out
for mode in nopin pin; do
  flag=""; [ $mode = pin ] && flag="--pin"
  KZB_SP_PROFILE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 \
    scripts/selfplay_bench.py --seconds 6 $flag > gpurun_out/sp_n8_$mode.out 2> gpurun_out/sp_n8_$mode.err
  grep '"metric"' gpurun_out/sp_n8_$mode.out | tail -1 > gpurun_out/sp_n8_$mode.json
  python - <<PY
import json
d = json.load(open("gpurun_out/sp_n8_$mode.json"))
print("$mode nodes/s %.0f nn/s %.0f mean_batch %.0f" % (d["value"], d["nn_positions_per_s"], d["mean_batch"]))
PY
  grep "thread CPU" gpurun_out/sp_n8_$mode.err | head -3
  grep "gather" gpurun_out/sp_n8_$mode.err | head -2
done
numactl --hardware 2>/dev/null | head -5; lscpu | egrep "NUMA|Socket" | head -6
