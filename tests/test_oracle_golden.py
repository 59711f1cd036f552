"""Pin the oracle (oracle/) against everything the reference offers for this path. CPU only."""
import numpy as np
import pytest

import oracle
from oracle.graph_exec import OnnxOracle
from helpers import GOLDEN, NET_NAMES, load_net_fixture, structure


# --- rust/kz-core/src/mapping/bit_buffer.rs:112-164: the reference's own unit-test vectors -------------
def _planes_from_bits(bits_list, count):
    bits = np.array([bits_list], dtype=np.uint8)
    out = oracle.expand_planes(bits, np.zeros((1, 0), np.float32), (1, 1, count), 0)
    return out.reshape(-1).astype(int).tolist()


def test_bitbuffer_short():  # bit_buffer.rs:117-124: push(true,false,true) -> storage [0b101]
    assert _planes_from_bits([0b101], 3) == [1, 0, 1]


def test_bitbuffer_edge_length():  # bit_buffer.rs:127-135
    assert _planes_from_bits([0b1111_1111], 8) == [1] * 8
    assert _planes_from_bits([0b1111_1111, 0b1], 9) == [1] * 9


def test_bitbuffer_longer():  # bit_buffer.rs:138-144: bits 1,5,12 set -> [0b0010_0010, 0b1_0000]
    expect = [int(i in (1, 5, 12)) for i in range(16)]
    assert _planes_from_bits([0b0010_0010, 0b1_0000], 16) == expect


def test_bitbuffer_block():  # bit_buffer.rs:157-163: push_block(0b1_0000_0001) -> LE bytes [1,1,0...]
    v = 0b1_0000_0001
    le = list(int(v).to_bytes(8, "little"))
    assert le[:2] == [1, 1]
    planes = _planes_from_bits(le, 64)
    assert planes == [(v >> i) & 1 for i in range(64)]


# --- planes produced by the reference's Python decoder (position.py:94-98,267-271) -------------------
@pytest.mark.parametrize("game", ["chess", "ataxx-7", "go-9", "ataxx-3", "go-19", "ttt", "arimaa-split"])
def test_expand_planes_vs_reference_python(game):
    d = np.load(GOLDEN / f"planes_{game}.npz")
    out = oracle.expand_planes(d["bits"], d["scalars"], tuple(int(v) for v in d["bool_shape"]),
                               int(d["scalar_count"]))
    # bit-exact, including the f32 scalars (compare bit patterns so -0.0/NaN could not hide)
    assert out.dtype == np.float32
    assert np.array_equal(out.view(np.uint32), d["planes"].view(np.uint32))


# --- network numerics: PyTorch fp32 on the reference's own model classes (save_onnx.py:94-102) --------
@pytest.mark.parametrize("name", NET_NAMES)
def test_graph_exec_vs_reference_pytorch(name):
    onnx_bytes, x, scalars, policy = load_net_fixture(name)
    net = OnnxOracle(onnx_bytes)
    out_s, out_p = net.run(x)
    assert out_s.shape == scalars.shape
    out_p = out_p.reshape(policy.shape)
    assert np.abs(out_s - scalars).max() <= 1e-4
    assert np.abs(out_p - policy).max() <= 1e-4


def test_export_structure_matches_survey():
    s = structure()
    ops = s["chess_conv_2x32"]["ops"]
    # Conv -> D x [Conv Relu Conv Relu Add] -> BatchNormalization -> heads
    assert ops[:12] == ["Conv"] + ["Conv", "Relu", "Conv", "Relu", "Add"] * 2 + ["BatchNormalization"]


# --- decode_output restatement (network/common.rs:16-114) -------------------------------------------
def test_decode_output_semantics():
    rng = np.random.default_rng(1)
    b, p = 4, 50
    scalars = rng.standard_normal((b, 5)).astype(np.float32)
    logits = rng.standard_normal((b, p)).astype(np.float32) * 3
    counts = [7, 0, 50, 1]  # includes a terminal board (common.rs:77 -> empty policy)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
    idx = np.concatenate([rng.permutation(p)[:c] for c in counts]).astype(np.uint32)
    vals, probs = oracle.decode_output(scalars, logits, idx, off)
    assert np.allclose(vals[:, 0], np.tanh(scalars[:, 0]), atol=1e-6)
    assert np.allclose(vals[:, 1:4].sum(axis=1), 1.0, atol=1e-6)
    assert np.array_equal(vals[:, 4], scalars[:, 4])  # moves_left passes through (common.rs:61)
    for i, c in enumerate(counts):
        seg = probs[off[i]:off[i + 1]]
        assert seg.shape[0] == c
        if c:
            ref = logits[i, idx[off[i]:off[i + 1]]].astype(np.float64)
            ref = np.exp(ref - ref.max())
            ref /= ref.sum()
            assert np.allclose(seg, ref, atol=1e-6)
            assert abs(seg.sum() - 1) < 1e-5


def test_decode_output_nan_panics_like_reference():
    scalars = np.zeros((1, 5), np.float32)
    logits = np.full((1, 4), np.nan, np.float32)
    with pytest.raises(FloatingPointError):  # common.rs:110 assert!(sum > 0.0)
        oracle.decode_output(scalars, logits, np.array([0, 1], np.uint32), np.array([0, 2], np.uint32))
