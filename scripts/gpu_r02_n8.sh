#!/bin/bash
# round 2: bench.py at N = 8 and N = 1 on the SAME 8-GPU box (the self-play sub-records give the 1 -> 8 curve of config C4)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nproc > gpurun_out/r02_n8_host.txt; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|L3" >> gpurun_out/r02_n8_host.txt; cat /sys/kernel/mm/transparent_hugepage/enabled >> gpurun_out/r02_n8_host.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 50 --warmup 5 --no-comparator --no-cpu-baseline > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
echo "n8 rc=$?"; tail -c 300 gpurun_out/r02_bench_n8.err
timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --no-comparator --no-cpu-baseline > gpurun_out/r02_bench_n1_on8box.json 2> gpurun_out/r02_bench_n1_on8box.err
echo "n1 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_n8.json", "gpurun_out/r02_bench_n1_on8box.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 4))
        for n, o in d.get("other_configs", {}).items(): print("  ", n, round(o["value"]), round(o["ms_per_step"], 3))
        for n, o in d.get("selfplay", {}).items(): print("  selfplay", n, round(o["value"]), "nn", round(o["nn_positions_per_s"]), "batch", round(o["mean_batch"]), "threads", o["cpu_threads_per_gpu"], o["gpu_threads_per_gpu"], o["executor_blocking_sync"], "cores", o["host_cores"])
    except Exception as e:
        print(f, "no line", repr(e))
PY
cat gpurun_out/r02_n8_host.txt
