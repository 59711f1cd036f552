#!/bin/bash
# Parity and timing of the halo-tile variant of the per-layer conv kernel (KZB_CONV_HALO=1, conv_tch.cu).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 60 -x -k "conv_halo" 2>&1 | tail -6
for h in 0 1; do
  echo "KZB_CONV_HALO=$h"
  KZB_CONV_HALO=$h timeout 60 python scripts/quick_profile.py --game go-9 --depth 20 --channels 256 --batch 2048 --iters 10 2>&1 | tail -1 | cut -c1-230 | tee -a gpurun_out/conv_halo_go9.jsonl
  KZB_CONV_HALO=$h timeout 60 python scripts/quick_profile.py --game go-19 --depth 40 --channels 256 --batch 512 --iters 3 2>&1 | tail -1 | cut -c1-230 | tee -a gpurun_out/conv_halo_go19.jsonl
done
