#!/bin/bash
# ncu evidence for the hot path (BASELINE.json north_star: tensor-pipe utilisation for the tower, HBM GB/s for encode / heads)
# and the go-9 20x256 batch sweep (configs[2]).  Run under gpurun; digests are made here with scripts/ncu_digest.py.
mkdir -p gpurun_out
# 1. launch list of the bench command
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches2.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list2.log 2>&1
tail -1 gpurun_out/ncu_list2.log | cut -c1-200
# 2. one full capture of each kernel of a chess step (after 3 warm evaluations = 9 launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tower8k|heads8|encode_kc" -s 9 -c 3 -f -o gpurun_out/chess_step \
    python scripts/quick_profile.py --iters 1 > gpurun_out/ncu_chess_step.log 2>&1
tail -2 gpurun_out/ncu_chess_step.log
# 3. go-9 20x256 sweep: plain timing, then tensor-pipe utilisation of the conv kernel under ncu
for b in 64 128 256 512 1024 2048 4096 8192; do
  timeout 300 python scripts/quick_profile.py --game go-9 --depth 20 --channels 256 --batch $b --iters 10 2>&1 | tail -1 > gpurun_out/go9_b$b.json
  cat gpurun_out/go9_b$b.json | cut -c1-220
  timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:conv_tc_kernel -s 132 -c 41 --csv --log-file gpurun_out/go9_ncu_b$b.csv \
      python scripts/quick_profile.py --game go-9 --depth 20 --channels 256 --batch $b --iters 1 > /dev/null 2>&1
done
