#!/bin/bash
# Parity and small-batch timing of the output-channel split of conv_tch (KZB_CONV_SPLIT=1).
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 60 -x -k "conv_split" 2>&1 | tail -5
: > gpurun_out/conv_split_go9.jsonl
for sp in 0 1; do for b in 64 256; do
  KZB_CONV_SPLIT=$sp timeout 60 python scripts/quick_profile.py --game go-9 --depth 20 --channels 256 --batch $b --iters 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); d['split']=$sp; print(json.dumps(d))" | tee -a gpurun_out/conv_split_go9.jsonl | cut -c1-150
done; done
