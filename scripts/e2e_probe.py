#!/usr/bin/env python
"""Where the end-to-end call spends its host time (development aid): KZB_TRACE=1 python scripts/e2e_probe.py"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kzero_b200 import netgen  # noqa: E402
from kzero_b200.network import B200Network, mapper_for  # noqa: E402

spec = netgen.game_spec("chess")
onnx_bytes = netgen.build_onnx(spec, 16, 128, seed=0)
inputs = [netgen.synthetic_positions(spec, 1024, seed=i) for i in range(4)]
net = B200Network(mapper_for(spec), onnx_bytes, 1024)
for i in range(20):
    net.evaluate_packed(*inputs[i % 4])
n = 300
t0 = time.perf_counter()
for i in range(n):
    net.evaluate_packed(*inputs[i % 4])
dt = (time.perf_counter() - t0) / n
net.stage_packed(*inputs[0])
dev = float(np.median(net.time_staged(50, False)))
print(f"e2e {dt * 1e6:.1f} us/call, device-only (no flush) {dev * 1e3:.1f} us, gap {dt * 1e6 - dev * 1e3:.1f} us")
net.close()
