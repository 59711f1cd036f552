#!/bin/bash
# round 2, first GPU visit: (1) the CTA-pair conv kernel (conv_tchp.cu) for the first time on hardware -- parity, then A/B
# against conv_tch with and without programmatic dependent launch; (2) parity at BASELINE sizes; (3) the library comparator.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pair kernel parity (bounded)"
KZB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x --timeout 120 -k "conv_pair" 2>&1 | tail -4
echo "pair_rc=$?"
echo "== A/B go-9 20x256"
for b in 4096 1024 256 64; do
  for v in "0 0" "0 1" "1 0" "1 1"; do
    set -- $v
    echo -n "b=$b pair=$1 pdl=$2 "
    KZB_CONV_PAIR=$1 KZB_PDL=$2 timeout 120 python scripts/quick_profile.py --game go-9 --depth 20 --channels 256 --batch $b --iters 10 2>&1 | tail -1 | cut -c1-260
  done
done | tee gpurun_out/r02_go9_pair_pdl_ab.txt
echo "== A/B go-19 40x256 b2048"
for v in "0 0" "1 0" "1 1"; do
  set -- $v
  echo -n "pair=$1 pdl=$2 "
  KZB_CONV_PAIR=$1 KZB_PDL=$2 timeout 200 python scripts/quick_profile.py --game go-19 --depth 40 --channels 256 --batch 2048 --iters 3 2>&1 | tail -1 | cut -c1-260
done | tee gpurun_out/r02_go19_pair_pdl_ab.txt
echo "== parity at size"
timeout 900 python -m pytest tests/test_gpu_parity_at_size.py -q --timeout 600 2>&1 | tail -25 | tee gpurun_out/r02_parity_at_size.txt
echo "== bench chess (with comparator), no extras"
timeout 600 python bench.py --steps 50 --warmup 5 --extras none > gpurun_out/r02_bench_chess_call1.json 2> gpurun_out/r02_bench_chess_call1.err; tail -c 600 gpurun_out/r02_bench_chess_call1.err
echo "== bench go9 / go19 (with comparator)"
timeout 600 python bench.py --config go9 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_go9_call1.json 2> gpurun_out/r02_bench_go9_call1.err; tail -c 600 gpurun_out/r02_bench_go9_call1.err
timeout 900 python bench.py --config go19 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_go19_call1.json 2> gpurun_out/r02_bench_go19_call1.err; tail -c 600 gpurun_out/r02_bench_go19_call1.err
python - <<'PY'
import json
for n in ("chess", "go9", "go19"):
    try:
        d = json.loads(open(f"gpurun_out/r02_bench_{n}_call1.json").read().strip().splitlines()[-1])
        print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["roofline"].items() if k in ("achieved", "frac", "kernel")})
        print("   comparator", {k: {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items()} for k, v in d.get("gpu_comparator", {}).get("variants", {}).items()}, d.get("gpu_comparator", {}).get("ours_vs_best_library"))
        for k in ("roofline_sustained", "roofline_k2", "roofline_k3"):
            if k in d: print("  ", k, round(d[k]["achieved"], 1), round(d[k]["frac"], 4))
    except Exception as e:
        print(n, "no line:", e)
PY
