import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
from kzero_b200 import netgen
from kzero_b200.network import B200Network, mapper_for
spec = netgen.game_spec("chess")
onnx_bytes = netgen.build_onnx(spec, 16, 128, seed=0)
bits, scalars, mv_idx, mv_off = netgen.synthetic_positions(spec, 1024, seed=1)
net = B200Network(mapper_for(spec), onnx_bytes, 1024)
net.stage_packed(bits, scalars, mv_idx, mv_off)
net.time_staged(3, True)
names, ms = net.profile_staged(True)
names, ms = net.profile_staged(True)
print(os.environ.get("KZB_DEBUG"), os.environ.get("KZB_TOWER_V1"), dict(zip(names, (ms*1000).round(1))))
