#!/bin/bash
# Last validation visit of the round: GPU suite, smoke, bench line, one self-play line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -2 | tee gpurun_out/final2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err; cut -c1-200 gpurun_out/final2_bench.json; tail -2 gpurun_out/final2_bench.err
timeout 120 python scripts/selfplay_bench.py --seconds 6 2>/dev/null | tail -1 > gpurun_out/final2_selfplay_chess_n1.json; cut -c1-200 gpurun_out/final2_selfplay_chess_n1.json
