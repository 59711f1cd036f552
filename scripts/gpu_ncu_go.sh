#!/bin/bash
# One ncu --set full capture of the per-layer conv kernel on the go-9 20x256 net (evidence for the next round's kernel work).
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel" -s 6 -c 2 -f -o gpurun_out/go9_conv_tc \
    python scripts/quick_profile.py --game go-9 --depth 20 --channels 256 --batch 2048 --iters 1 > gpurun_out/go9_ncu.log 2>&1
tail -2 gpurun_out/go9_ncu.log
timeout 120 python scripts/quick_profile.py --game go-9 --depth 20 --channels 256 --batch 2048 --iters 10 2>&1 | tail -1
