#!/bin/bash
# BASELINE.json configs[2]: go-9 20x256, tensor-pipe utilisation sweep over batch 64..8192 with the round-2 kernels.
# Timing: CUDA events, median of 10 whole evaluations, L2 flushed.  Tensor pipe: ncu over the 41 conv_i2c launches of ONE evaluation.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cat > /tmp/one_eval.py <<'PY'
import sys
sys.path.insert(0, '.')
from kzero_b200 import netgen
from kzero_b200.network import B200Network, mapper_for
b = int(sys.argv[1])
spec = netgen.game_spec("go-9")
onnx = netgen.build_onnx(spec, 20, 256, seed=0)
inp = netgen.synthetic_positions(spec, b, seed=1)
import os
os.environ["KZB_NO_GRAPH"] = "1"
with B200Network(mapper_for(spec), onnx, b) as net:
    net.evaluate_packed(*inp)
    net.evaluate_packed(*inp)
PY
: > gpurun_out/r02_go9_sweep.jsonl
for b in 64 128 256 512 1024 2048 4096 8192; do
  timeout 120 python scripts/quick_profile.py --game go-9 --depth 20 --channels 256 --batch $b --iters 10 2>&1 | tail -1 >> gpurun_out/r02_go9_sweep.jsonl
  timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:conv_i2c -s 42 -c 42 --csv --log-file gpurun_out/r02_go9_ncu_b$b.csv python /tmp/one_eval.py $b > /dev/null 2>&1
  echo "b=$b done rc=$?"
done
python - <<'PY'
import csv, json
rows = [json.loads(l) for l in open("gpurun_out/r02_go9_sweep.jsonl")]
out = ["| batch | ms / step | positions/s | TFLOP/s (algorithmic) | of burst 1613.1 | of sustained 1350.8 | tensor pipe % elapsed | tensor pipe % active | conv launches sampled |", "|---|---|---|---|---|---|---|---|---|"]
for d in rows:
    b = d["batch"]
    t = el = ac = 0.0
    n = 0
    try:
        rd = list(csv.reader(l for l in open(f"gpurun_out/r02_go9_ncu_b{b}.csv") if l.startswith('"')))
        hdr = rd[0]
        ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
        idi = hdr.index("ID")
        per = {}
        for r in rd[1:]:
            per.setdefault(r[idi], {})[r[mi]] = float(r[vi].replace(",", ""))
        for k, m in per.items():
            dur = m.get("gpu__time_duration.sum", 0.0)
            t += dur
            el += dur * m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
            ac += dur * m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
            n += 1
    except Exception as e:
        print("ncu parse failed for", b, repr(e))
    out.append(f"| {b} | {d['ms_median']:.3f} | {d['pos_per_s']:,.0f} | {d['tflops']:.0f} | {d['tflops'] / 1613.1:.2f} | {d['tflops'] / 1350.8:.2f} | "
               f"{el / t if t else float('nan'):.1f} | {ac / t if t else float('nan'):.1f} | {n} |")
open("gpurun_out/r02_go9_sweep_table.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
PY
