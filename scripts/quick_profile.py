#!/usr/bin/env python
"""Per-launch CUDA-event timing of one staged batch (development aid; bench.py is the contract)."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from kzero_b200 import netgen  # noqa: E402
from kzero_b200.network import B200Network, mapper_for  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--game", default="chess")
ap.add_argument("--depth", type=int, default=16)
ap.add_argument("--channels", type=int, default=128)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--precision", type=int, default=1)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()

spec = netgen.game_spec(args.game)
onnx_bytes = netgen.build_onnx(spec, args.depth, args.channels, seed=0)
bits, scalars, mv_idx, mv_off = netgen.synthetic_positions(spec, args.batch, seed=1)
net = B200Network(mapper_for(spec), onnx_bytes, args.batch, precision=args.precision)
info = net.info()
for _ in range(3):
    net.evaluate_packed(bits, scalars, mv_idx, mv_off)
net.stage_packed(bits, scalars, mv_idx, mv_off)
net.time_staged(3, True)
names, ms = net.profile_staged(True)
for n, m in zip(names, ms):
    print(f"{n:24s} {m * 1000:9.1f} us")
tot = net.time_staged(args.iters, True)
flops = info.flops_per_position * args.batch
print(json.dumps({"game": args.game, "depth": args.depth, "channels": args.channels, "batch": args.batch,
                  "conv_mode": info.conv_mode, "ms_median": float(np.median(tot)), "ms_min": float(tot.min()),
                  "pos_per_s": args.batch / (float(np.median(tot)) * 1e-3),
                  "tflops": flops / (float(np.median(tot)) * 1e-3) / 1e12,
                  "sum_steps_ms": float(ms.sum())}))
