#!/bin/bash
# compute-sanitizer passes over one small evaluation of every kernel family (run under gpurun)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
from kzero_b200 import netgen
from kzero_b200.network import B200Network, mapper_for
# chess: encode_kc + tower8k + heads8 (with programmatic dependent launch); chess 1x256: conv_i2c on an 8x8 board + heads8;
# chess-att: conv_tc 1x1 chunks + attention tail; go-9 / go-19: conv_i2c (3x3 and the 1x1 policy conv, pair tiles past the batch,
# channel split at small batches) + conv_tc head convs + heads_tail
for game, depth, ch, n in [("chess", 2, 128, 20), ("chess", 1, 256, 5), ("chess-att", 1, 64, 6), ("ataxx-7", 2, 64, 9), ("go-9", 2, 64, 5),
                           ("go-9", 1, 256, 40), ("go-19", 1, 64, 3)]:
    spec = netgen.game_spec(game)
    onnx = netgen.build_onnx(spec, depth, ch, seed=1)
    b, s, mi, mo = netgen.synthetic_positions(spec, n, seed=2)
    for prec in (1, 0):
        with B200Network(mapper_for(spec), onnx, n, precision=prec) as net:
            for _ in range(3):  # direct launches, graph capture, graph replay
                v, p = net.evaluate_packed(b, s, mi, mo)
            assert np.isfinite(v).all() and np.isfinite(p).all()
    print("ok", game, flush=True)
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py > gpurun_out/sanitize_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|error" gpurun_out/sanitize_$tool.log | head -12
done
