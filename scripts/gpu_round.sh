#!/bin/bash
# One GPU visit: parity tests, per-launch timings for the BASELINE configs, one ncu capture of the tower kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -30 > gpurun_out/pytest.log
cat gpurun_out/pytest.log
for cfg in "chess 16 128 1024" "ataxx-7 8 64 256" "go-9 20 256 4096"; do
  set -- $cfg
  timeout 300 python scripts/quick_profile.py --game $1 --depth $2 --channels $3 --batch $4 > gpurun_out/prof_$1.log 2>&1
  tail -1 gpurun_out/prof_$1.log
done
KZB_FORCE_LINEAR=1 timeout 300 python scripts/quick_profile.py --game chess --depth 16 --channels 128 --batch 1024 > gpurun_out/prof_chess_linear.log 2>&1
tail -1 gpurun_out/prof_chess_linear.log
if [ -n "$NCU" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 45 -c 2 -f -o gpurun_out/conv_tc \
     python scripts/quick_profile.py --game chess --depth 16 --channels 128 --batch 1024 --iters 1 > gpurun_out/ncu.log 2>&1
  tail -3 gpurun_out/ncu.log
fi
