#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -2
timeout 60 python scripts/quick_profile.py --game go-9 --depth 20 --channels 256 --batch 4096 --iters 5 2>&1 | tail -1 | cut -c1-200
timeout 60 python scripts/quick_profile.py --game go-19 --depth 40 --channels 256 --batch 512 --iters 3 2>&1 | tail -1 | cut -c1-200
