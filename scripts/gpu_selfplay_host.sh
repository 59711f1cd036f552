#!/bin/bash
# Self-play host-efficiency visit: GPU self-play tests, the N=1 bench, and the same bench pinned to 4 host cores
# (the per-GPU share of the 8-GPU box) with the per-section generator profile.
mkdir -p gpurun_out
lscpu | egrep "Model name|^CPU\(s\)|L2|L3|Thread" > gpurun_out/sp_host_cpu.txt; cat /sys/kernel/mm/transparent_hugepage/enabled >> gpurun_out/sp_host_cpu.txt
timeout 600 python -m pytest tests/test_gpu_selfplay.py -m gpu -q --timeout 180 2>&1 | tail -2
timeout 120 python scripts/selfplay_bench.py --seconds 8 2>/dev/null | tail -1 > gpurun_out/sp_chess_n1.json; cut -c1-330 gpurun_out/sp_chess_n1.json
KZB_SP_PROFILE=1 timeout 120 taskset -c 0-3 python scripts/selfplay_bench.py --seconds 8 > gpurun_out/sp_chess_4cores.json 2> gpurun_out/sp_chess_4cores.err; cut -c1-330 gpurun_out/sp_chess_4cores.json; grep "kzb selfplay" gpurun_out/sp_chess_4cores.err
timeout 120 taskset -c 0-3 python scripts/selfplay_bench.py --seconds 8 --game ataxx 2>/dev/null | tail -1 > gpurun_out/sp_ataxx_4cores.json; cut -c1-330 gpurun_out/sp_ataxx_4cores.json
timeout 120 python scripts/selfplay_bench.py --seconds 8 --game ataxx 2>/dev/null | tail -1 > gpurun_out/sp_ataxx_n1.json; cut -c1-330 gpurun_out/sp_ataxx_n1.json
