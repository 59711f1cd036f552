#!/bin/bash
# round 2, thirteenth 1-GPU visit: validation after the session / server changes (interrupt-and-continue, several devices, warm-up in bench.py)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== GPU suite"
timeout 600 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/r02_pytest_gpu_call13.txt 2>&1; tail -3 gpurun_out/r02_pytest_gpu_call13.txt
echo "== self-play, 4 cores, with and without warm-up"
{
echo -n "4 cores chess-real 512 games, no warm-up: "; timeout 120 taskset -c 0-3 python scripts/selfplay_bench.py --game chess-real --seconds 6 --concurrent-games 512 2>&1 | tail -1 | cut -c1-260
echo -n "4 cores chess-real 512 games, 3 s warm-up: "; timeout 120 taskset -c 0-3 python scripts/selfplay_bench.py --game chess-real --seconds 6 --concurrent-games 512 --warmup-seconds 3 2>&1 | tail -1 | cut -c1-260
} | tee gpurun_out/r02_selfplay_call13.txt
echo "== bench (chess + self-play sub-records only)"
timeout 600 python bench.py --no-comparator --no-cpu-baseline > gpurun_out/r02_bench_call13.json 2> gpurun_out/r02_bench_call13.err; echo "rc=$?"; tail -c 300 gpurun_out/r02_bench_call13.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_call13.json").read().strip().splitlines()[-1])
print("chess value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "tower frac", round(d["roofline"]["frac"], 4))
for n, o in d["selfplay"].items(): print("selfplay", n, round(o["value"]), round(o["nn_positions_per_s"]), round(o["cache_hit_rate"], 3), o["concurrent_games_per_gpu"], o["warmup_seconds"])
PY
