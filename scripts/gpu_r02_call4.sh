#!/bin/bash
# round 2, fourth GPU visit: whole GPU suite on the cleaned-up kernel set (conv_i2c default, superseded kernels deleted),
# then bench.py with the sub-records for the other BASELINE configs and the self-play loop.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== GPU suite"
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/r02_pytest_gpu.txt 2>&1; tail -5 gpurun_out/r02_pytest_gpu.txt; grep -E "^\[" gpurun_out/r02_pytest_gpu.txt | cut -c1-300
echo "== bench (default: chess + extras)"
timeout 1500 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "rc=$?"; tail -c 800 gpurun_out/r02_bench_default.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_default.json").read().strip().splitlines()[-1])
    def short(x):
        return {k: (round(v, 4) if isinstance(v, float) else v) for k, v in x.items() if not isinstance(v, (dict, list))}
    print("chess", short(d)); print(" roofline", short(d["roofline"]))
    for k in ("roofline_sustained", "roofline_k2", "roofline_k3"):
        print(" ", k, short(d.get(k, {})))
    print(" comparator", {k: short(v) for k, v in d.get("gpu_comparator", {}).get("variants", {}).items()}, d.get("gpu_comparator", {}).get("ours_vs_best_library"))
    for name, o in d.get("other_configs", {}).items():
        print(name, short(o)); print(" roofline", short(o["roofline"])); print(" sustained", short(o.get("roofline_sustained", {})))
        print(" comparator", {k: short(v) for k, v in o.get("gpu_comparator", {}).get("variants", {}).items()}, o.get("gpu_comparator", {}).get("ours_vs_best_library"))
    for name, o in d.get("selfplay", {}).items():
        print("selfplay", name, short(o))
    print("cpu", short(d.get("cpu_baseline", {})))
except Exception as e:
    print("no line:", repr(e))
PY
