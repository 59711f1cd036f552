#!/bin/bash
# round 2, last 1-GPU visit: DRAM traffic of a go-19 40x256 layer at batch 8192 (the one `traffic` the bench line still reported as null)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"conv_i2c" -s 50 -c 2 -f -o gpurun_out/r02_go19_i2c_b8192 \
    python scripts/quick_profile.py --game go-19 --depth 40 --channels 256 --batch 8192 --iters 1 > gpurun_out/r02_ncu_go19.log 2>&1
echo "go19 full rc=$?"; tail -2 gpurun_out/r02_ncu_go19.log | cut -c1-200
ls -la gpurun_out/r02_go19_i2c_b8192.ncu-rep
