#!/bin/bash
# round 2, sixth 1-GPU visit: compute-sanitizer over the new kernels, the whole GPU suite, the default bench line.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== GPU suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02_pytest_gpu.txt 2>&1; tail -4 gpurun_out/r02_pytest_gpu.txt
echo "== sanitizers"
bash scripts/sanitize.sh 2>&1 | tail -30
echo "== chess A/B: PDL off / on (device ms per step, 100 steps, L2 flushed)"
for pdl in 0 1; do
  echo -n "KZB_PDL=$pdl "
  KZB_PDL=$pdl timeout 120 python scripts/quick_profile.py --game chess --depth 16 --channels 128 --batch 1024 --iters 100 2>&1 | tail -1 | cut -c1-220
done | tee gpurun_out/r02_chess_pdl_ab.txt
echo "== bench default"
timeout 1500 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "rc=$?"; tail -c 400 gpurun_out/r02_bench_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_default.json").read().strip().splitlines()[-1])
print("chess value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "tower frac", round(d["roofline"]["frac"], 4), "sustained", round(d["roofline_sustained"]["frac"], 4), "vs cudnn", round(d["gpu_comparator"]["ours_vs_best_library"], 3))
print(d["step_breakdown_ms"])
for n, o in d["other_configs"].items(): print(n, round(o["value"]), round(o["ms_per_step"], 3), round(o["roofline"]["frac"], 4), round(o["roofline_sustained"]["frac"], 4), round(o["gpu_comparator"]["ours_vs_best_library"], 3))
for n, o in d["selfplay"].items(): print("selfplay", n, round(o["value"]), round(o["nn_positions_per_s"]), o["concurrent_games_per_gpu"])
PY
