#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_selfplay.py -m gpu -q --timeout 120 2>&1 | tail -1
KZB_SP_PROFILE=1 timeout 120 taskset -c 0-3 python scripts/selfplay_bench.py --seconds 6 > gpurun_out/q_sp_4cores.json 2> gpurun_out/q_sp_4cores.err; cut -c1-200 gpurun_out/q_sp_4cores.json; grep "kzb selfplay" gpurun_out/q_sp_4cores.err
KZB_SP_PROFILE=1 timeout 120 taskset -c 0-1 python scripts/selfplay_bench.py --seconds 6 --cpu-threads 2 > gpurun_out/q_sp_2cores.json 2> gpurun_out/q_sp_2cores.err; cut -c1-200 gpurun_out/q_sp_2cores.json; grep "thread CPU" gpurun_out/q_sp_2cores.err
