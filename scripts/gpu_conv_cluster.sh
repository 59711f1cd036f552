#!/bin/bash
# Parity and timing of the 2-CTA weight-multicast variant of the per-layer conv kernel (KZB_CONV_CLUSTER=2).
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -k "conv_cluster" 2>&1 | tail -4
for cl in 1 2; do
  echo "KZB_CONV_CLUSTER=$cl"
  KZB_CONV_CLUSTER=$cl timeout 120 python scripts/quick_profile.py --game go-9 --depth 20 --channels 256 --batch 2048 --iters 10 2>&1 | tail -1 | cut -c1-230 | tee -a gpurun_out/conv_cluster_go9.jsonl
  KZB_CONV_CLUSTER=$cl timeout 120 python scripts/quick_profile.py --game go-19 --depth 40 --channels 256 --batch 512 --iters 3 2>&1 | tail -1 | cut -c1-230 | tee -a gpurun_out/conv_cluster_go19.jsonl
done
