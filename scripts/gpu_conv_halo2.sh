#!/bin/bash
# conv_tch as the default: whole GPU suite, compute-sanitizer over small go / linear-chess evaluations, timing.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -3
cat > /tmp/san_halo.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
from kzero_b200 import netgen
from kzero_b200.network import B200Network, mapper_for
for game, depth, ch, n in [("go-9", 2, 64, 5), ("go-9", 1, 256, 3), ("go-19", 1, 64, 2)]:
    spec = netgen.game_spec(game)
    onnx = netgen.build_onnx(spec, depth, ch, seed=1)
    b, s, mi, mo = netgen.synthetic_positions(spec, n, seed=2)
    with B200Network(mapper_for(spec), onnx, n, precision=1) as net:
        v, p = net.evaluate_packed(b, s, mi, mo)
        assert np.isfinite(v).all() and np.isfinite(p).all()
    print("ok", game, ch, flush=True)
PY
for tool in memcheck racecheck; do
  timeout 300 compute-sanitizer --tool $tool python /tmp/san_halo.py > gpurun_out/san_halo_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok" gpurun_out/san_halo_$tool.log | tail -4
done
timeout 60 python scripts/quick_profile.py --game go-9 --depth 20 --channels 256 --batch 4096 --iters 5 2>&1 | tail -1 | cut -c1-230 | tee gpurun_out/conv_halo_go9_b4096.json
