#!/bin/bash
# go-9 20x256 batch sweep (BASELINE.json configs[2]) with the current default kernels: CUDA-event timing only.
mkdir -p gpurun_out
: > gpurun_out/go9_sweep.jsonl
for b in 64 128 256 512 1024 2048 4096 8192; do
  timeout 120 python scripts/quick_profile.py --game go-9 --depth 20 --channels 256 --batch $b --iters 10 2>&1 | tail -1 >> gpurun_out/go9_sweep.jsonl
done
python - <<'PY'
import json
for l in open("gpurun_out/go9_sweep.jsonl"):
    d = json.loads(l)
    print(d["batch"], round(d["ms_median"], 3), round(d["pos_per_s"]), round(d["tflops"]))
PY
timeout 200 python scripts/quick_profile.py --game go-19 --depth 40 --channels 256 --batch 8192 --iters 3 2>&1 | tail -1 | cut -c1-220 | tee gpurun_out/go19_b8192.json
