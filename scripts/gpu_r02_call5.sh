#!/bin/bash
# round 2, fifth GPU visit: the new boundary tests (raw weights, go territory planes, per-batch CUDA graphs, GPU-played records),
# host-bound self-play baselines (2 and 4 cores, synthetic and real chess) for the host work that follows.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== new GPU tests"
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -k "raw_weights or territory or cuda_graph or gpu_played or selfplay or server" 2>&1 | tail -4
echo "== whole GPU suite"
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -3
echo "== self-play, host-bound baselines"
for game in chess chess-real; do
  for cores in 2 4; do
    echo -n "game=$game cores=$cores "
    KZB_SP_PROFILE=1 timeout 120 taskset -c 0-$((cores-1)) python scripts/selfplay_bench.py --game $game --seconds 6 2> gpurun_out/r02_sp_${game}_${cores}c.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('value','nn_positions_per_s','mean_batch','cache_hit_rate','moves_per_s')}, d['config']['cpu_threads_per_gpu'], d['config']['gpu_threads_per_gpu'], d['config']['executor_blocking_sync'])"
    grep "kzb selfplay" gpurun_out/r02_sp_${game}_${cores}c.err | tr '\n' ';' | cut -c1-600; echo
  done
done | tee gpurun_out/r02_sp_host_baseline.txt
echo "== go-9 small batches with graphs (e2e through the public call)"
python - <<'PY' | tee gpurun_out/r02_go9_small_batch_e2e.txt
import sys, time
sys.path.insert(0, '.')
import numpy as np
from kzero_b200 import netgen
from kzero_b200.network import B200Network, mapper_for
spec = netgen.game_spec("go-9")
onnx = netgen.build_onnx(spec, 20, 256, seed=0)
for b in (16, 64, 256, 1024):
    inp = netgen.synthetic_positions(spec, b, seed=1)
    with B200Network(mapper_for(spec), onnx, b) as net:
        for _ in range(5): net.evaluate_packed(*inp)
        t0 = time.perf_counter()
        n = 50
        for _ in range(n): net.evaluate_packed(*inp)
        dt = (time.perf_counter() - t0) / n
        net.stage_packed(*inp)
        dev = float(np.median(net.time_staged(20, False)))
    print(f"go-9 20x256 batch {b}: e2e {dt*1e3:.3f} ms/call ({b/dt:,.0f} pos/s), device {dev:.3f} ms")
PY
