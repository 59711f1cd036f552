"""netgen (synthetic nets in the reference's export structure) and mapping tables, CPU only."""
import numpy as np
import pytest

from helpers import GOLDEN, structure
from kzero_b200 import mapping, netgen
from oracle.graph_exec import OnnxOracle
from oracle.onnx_min import load_model


def _gather_constant(name):
    m = load_model((GOLDEN / f"net_{name}.onnx").read_bytes())
    node = [n for n in m.nodes if n.op == "Gather" and n.outputs == ["policy"]][0]
    if node.inputs[1] in m.initializers:  # registered buffer (attention head, post_act.py:126)
        return m.initializers[node.inputs[1]]
    const = [n for n in m.nodes if n.op == "Constant" and n.outputs[0] == node.inputs[1]][0]
    return np.asarray(const.attrs["value"])


def test_chess_flat_to_conv_matches_reference_export():
    # the Gather constant inside the exported net IS python/lib/mapping/chess_flat_to_conv.txt
    assert np.array_equal(mapping.chess_flat_to_conv(), _gather_constant("chess_conv_2x32"))


def test_chess_flat_to_att_matches_reference_export():
    assert np.array_equal(mapping.chess_flat_to_att(), _gather_constant("chess_att_2x32"))


def test_chess_flat_moves_known_answers():
    mv = mapping.chess_flat_moves_pov()
    assert len(mv) == 1880 and len(set(mv)) == 1880  # tests/mapper/chess/mod.rs:6-17
    # tests/mapper/chess/pairs.rs: queen N distance 1 from a1 is channel 0; knight NNE is channel 56
    assert mapping.chess_conv_channel((0, 8, None)) == 0
    assert mapping.chess_conv_channel((0, 17, None)) == 56
    # queen promotion is a plain queen move north (chess.rs:318-324), under-promotions get 64..72
    assert mapping.chess_conv_channel((48, 56, "q")) == 0
    assert mapping.chess_conv_channel((48, 56, "r")) == 64 + 1 * 3 + 0
    assert mapping.chess_conv_channel((49, 56, "n")) == 64 + 0 * 3 + 2


@pytest.mark.parametrize("fixture,game,depth,ch", [
    ("ataxx7_2x32", "ataxx-7", 2, 32),
    ("chess_conv_2x32", "chess", 2, 32),
    ("go9_2x32", "go-9", 2, 32),
    ("ataxx5_scripted_1x16", "ataxx-5", 1, 16),
])
def test_netgen_structure_equals_reference_export(fixture, game, depth, ch):
    ref = structure()[fixture]
    spec = netgen.game_spec(game)
    data = netgen.build_onnx(spec, depth, ch, seed=1)
    m = load_model(data)
    assert [n.op for n in m.nodes] == ref["ops"]
    assert m.inputs[0][1][1:] == ref["input_shape"]
    assert spec.policy_size == int(np.prod(ref["policy_shape"]))


@pytest.mark.parametrize("game", ["chess", "chess-att", "ataxx-7", "go-9"])
def test_netgen_nets_run_in_oracle(game):
    spec = netgen.game_spec(game)
    data = netgen.build_onnx(spec, 2, 32, seed=2, query_channels=16)
    x = np.random.default_rng(0).standard_normal((3, spec.input_channels, spec.board_size, spec.board_size))
    s, p = OnnxOracle(data).run(x.astype(np.float32))
    assert s.shape == (3, 5) and p.shape == (3, spec.policy_size)
    assert np.isfinite(s).all() and np.isfinite(p).all()


def test_netgen_unfolded_bn_equals_folded():
    spec = netgen.game_spec("ataxx-7")
    x = np.random.default_rng(0).standard_normal((2, 4, 7, 7)).astype(np.float32)
    a = OnnxOracle(netgen.build_onnx(spec, 2, 16, seed=3, fold_bn=True)).run(x)
    b = OnnxOracle(netgen.build_onnx(spec, 2, 16, seed=3, fold_bn=False)).run(x)
    assert np.abs(a[0] - b[0]).max() < 1e-4 and np.abs(a[1] - b[1]).max() < 1e-4


def test_synthetic_positions_layout():
    for game in ["chess", "ataxx-7", "go-9"]:
        spec = netgen.game_spec(game)
        bits, scalars, idx, off = netgen.synthetic_positions(spec, 17, seed=5)
        assert bits.shape == (17, spec.bits_bytes) and scalars.shape == (17, spec.scalar_channels)
        assert off[0] == 0 and off[-1] == idx.shape[0] and (np.diff(off.astype(np.int64)) >= 1).all()
        assert idx.max() < spec.policy_size
        for i in range(17):  # no two legal moves share an index (tests/mapper/mod.rs:60-75)
            seg = idx[off[i]:off[i + 1]]
            assert len(set(seg.tolist())) == len(seg)
