#!/bin/bash
# One consolidated GPU visit: parity suite, smoke, bench (both arms), self-play, ncu launch list + full captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -3 | tee gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; cut -c1-300 gpurun_out/final_bench.json; tail -2 gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final_bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/final_bench_ref.json
timeout 120 python scripts/selfplay_bench.py --seconds 8 2>/dev/null | tail -1 > gpurun_out/final_selfplay_chess.json; cut -c1-330 gpurun_out/final_selfplay_chess.json
timeout 120 python scripts/selfplay_bench.py --seconds 8 --game ataxx 2>/dev/null | tail -1 > gpurun_out/final_selfplay_ataxx.json; cut -c1-330 gpurun_out/final_selfplay_ataxx.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tower8k|heads8|encode_kc" -s 9 -c 3 -f -o gpurun_out/final_chess_step \
    python scripts/quick_profile.py --iters 1 > gpurun_out/final_ncu.log 2>&1
tail -1 gpurun_out/final_ncu.log
