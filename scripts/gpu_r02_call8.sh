#!/bin/bash
# round 2, eighth 1-GPU visit: the bitboard chess generator in the loop (all cores / 4 cores = one GPU's share of an 8-GPU host), go rules with
# superko, where the end-to-end call spends its host time, the whole GPU suite, the default bench line.  Every step is bounded.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== GPU suite"
timeout 600 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/r02_pytest_gpu_call8.txt 2>&1; tail -4 gpurun_out/r02_pytest_gpu_call8.txt
echo "== e2e host-time breakdown (KZB_TRACE=1)"
KZB_TRACE=1 timeout 120 python scripts/e2e_probe.py 2>&1 | tail -4 | tee gpurun_out/r02_e2e_trace.txt
echo "== self-play on one GPU: host cores"; nproc
{
for game in chess chess-real; do
  echo -n "all cores $game: "; timeout 120 python scripts/selfplay_bench.py --game $game --seconds 5 --concurrent-games 384 2>&1 | tail -1 | cut -c1-330
  echo -n "4 cores $game: "; KZB_SP_PROFILE=1 timeout 120 taskset -c 0-3 python scripts/selfplay_bench.py --game $game --seconds 5 --concurrent-games 384 2>&1 | grep -E "cycles / node|CPU time|metric" | cut -c1-330
done
echo -n "4 cores go: "; timeout 120 taskset -c 0-3 python scripts/selfplay_bench.py --game go --seconds 5 --gpu-batch 1024 2>&1 | tail -1 | cut -c1-330
} | tee gpurun_out/r02_selfplay_call8.txt
echo "== bench default"
timeout 900 python bench.py > gpurun_out/r02_bench_default_call8.json 2> gpurun_out/r02_bench_default_call8.err; echo "rc=$?"; tail -c 400 gpurun_out/r02_bench_default_call8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_default_call8.json").read().strip().splitlines()[-1])
print("chess value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "tower frac", round(d["roofline"]["frac"], 4), "sustained", round(d["roofline_sustained"]["frac"], 4), "vs cudnn", round(d["gpu_comparator"]["ours_vs_best_library"], 3))
print(d["step_breakdown_ms"])
for n, o in d["other_configs"].items(): print(n, round(o["value"]), round(o["ms_per_step"], 3), round(o["roofline"]["frac"], 4), round(o["roofline_sustained"]["frac"], 4), round(o["gpu_comparator"]["ours_vs_best_library"], 3))
for n, o in d["selfplay"].items(): print("selfplay", n, round(o["value"]), round(o["nn_positions_per_s"]), o["concurrent_games_per_gpu"])
PY
