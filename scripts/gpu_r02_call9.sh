#!/bin/bash
# round 2, ninth 1-GPU visit: ncu evidence of the shipped code -- the launch list of the bench command (shares vs step_breakdown_ms),
# one `--set full` capture of each kernel of a chess step (DRAM traffic, tensor pipe), the same for one go-9 layer.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_chess_b1024_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-comparator --extras none > gpurun_out/r02_ncu_list.log 2>&1
echo "launch list rc=$?"; tail -1 gpurun_out/r02_ncu_list.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tower8k|heads8|encode_kc" -s 9 -c 3 -f -o gpurun_out/r02_chess_step \
    python scripts/quick_profile.py --iters 1 > gpurun_out/r02_ncu_chess_step.log 2>&1
echo "chess full rc=$?"; tail -2 gpurun_out/r02_ncu_chess_step.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_i2c" -s 50 -c 2 -f -o gpurun_out/r02_go9_i2c_b4096 \
    python scripts/quick_profile.py --game go-9 --depth 20 --channels 256 --batch 4096 --iters 1 > gpurun_out/r02_ncu_go9.log 2>&1
echo "go full rc=$?"; tail -2 gpurun_out/r02_ncu_go9.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
