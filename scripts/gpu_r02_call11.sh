#!/bin/bash
# round 2, eleventh 1-GPU visit: validation of the shipped code -- the whole GPU suite, the default bench line, real chess / go in the loop on 4 cores
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== GPU suite"
timeout 600 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/r02_pytest_gpu_call11.txt 2>&1; tail -4 gpurun_out/r02_pytest_gpu_call11.txt
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== self-play, one GPU, 4 cores"
{
for game in chess chess-real; do
  echo -n "4 cores $game: "; KZB_SP_PROFILE=1 timeout 120 taskset -c 0-3 python scripts/selfplay_bench.py --game $game --seconds 5 --concurrent-games 384 2>&1 | grep -E "gather|CPU time|metric" | cut -c1-330
done
} | tee gpurun_out/r02_selfplay_call11.txt
echo "== bench default"
timeout 900 python bench.py > gpurun_out/r02_bench_default_call11.json 2> gpurun_out/r02_bench_default_call11.err; echo "rc=$?"; tail -c 400 gpurun_out/r02_bench_default_call11.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_default_call11.json").read().strip().splitlines()[-1])
print("chess value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "tower frac", round(d["roofline"]["frac"], 4), "traffic", d["roofline"]["traffic"], "sustained", round(d["roofline_sustained"]["frac"], 4), "vs cudnn", round(d["gpu_comparator"]["ours_vs_best_library"], 3))
print("k2", d["roofline_k2"]["traffic"], "k3", d["roofline_k3"]["traffic"], d["step_breakdown_ms"])
for n, o in d["other_configs"].items(): print(n, round(o["value"]), round(o["ms_per_step"], 3), round(o["roofline"]["frac"], 4), o["roofline"]["traffic"], round(o["roofline_sustained"]["frac"], 4), round(o["gpu_comparator"]["ours_vs_best_library"], 3))
for n, o in d["selfplay"].items(): print("selfplay", n, round(o["value"]), round(o["nn_positions_per_s"]), o["concurrent_games_per_gpu"])
PY
