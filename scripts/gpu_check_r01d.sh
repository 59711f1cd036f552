#!/bin/bash
# Validation visit after the self-play host rework: GPU suite, smoke, the bench line, self-play on all cores and on 4.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -3 | tee gpurun_out/r01d_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r01d_bench.json 2> gpurun_out/r01d_bench.err; cut -c1-260 gpurun_out/r01d_bench.json; tail -2 gpurun_out/r01d_bench.err
timeout 120 python scripts/selfplay_bench.py --seconds 8 2>/dev/null | tail -1 > gpurun_out/r01d_selfplay_chess_n1.json; cut -c1-300 gpurun_out/r01d_selfplay_chess_n1.json
KZB_SP_PROFILE=1 timeout 120 taskset -c 0-3 python scripts/selfplay_bench.py --seconds 8 > gpurun_out/r01d_selfplay_chess_4cores.json 2> gpurun_out/r01d_selfplay_chess_4cores.err; cut -c1-300 gpurun_out/r01d_selfplay_chess_4cores.json; grep "kzb selfplay" gpurun_out/r01d_selfplay_chess_4cores.err
KZB_SP_PROFILE=1 timeout 120 taskset -c 0-1 python scripts/selfplay_bench.py --seconds 8 --cpu-threads 2 > gpurun_out/r01d_selfplay_chess_2cores.json 2> gpurun_out/r01d_selfplay_chess_2cores.err; cut -c1-300 gpurun_out/r01d_selfplay_chess_2cores.json; grep "thread CPU" gpurun_out/r01d_selfplay_chess_2cores.err
timeout 120 python scripts/selfplay_bench.py --seconds 8 --game ataxx 2>/dev/null | tail -1 > gpurun_out/r01d_selfplay_ataxx_n1.json; cut -c1-300 gpurun_out/r01d_selfplay_ataxx_n1.json
