"""Shared test helpers: golden-file readers (reference formats) and synthetic position generators."""
import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"

NET_NAMES = ["ataxx7_2x32", "chess_conv_2x32", "chess_att_2x32", "go9_2x32", "ataxx5_scripted_1x16"]


def structure():
    return json.loads((GOLDEN / "export_structure.json").read_text())


def load_net_fixture(name):
    """-> (onnx_bytes, x [b,C,H,W], scalars [b,5], policy [b,P]) from the reference's check-file format
    (python/lib/save_onnx.py:94-102): 1 byte batch, raw f32 inputs, raw f32 outputs."""
    info = structure()[name]
    onnx_bytes = (GOLDEN / f"net_{name}.onnx").read_bytes()
    raw = (GOLDEN / f"net_{name}.bin").read_bytes()
    b = raw[0]
    assert b == info["batch"]
    data = np.frombuffer(raw[1:], dtype="<f4")
    n_in = b * int(np.prod(info["input_shape"]))
    p = int(np.prod(info["policy_shape"]))
    assert data.size == n_in + b * 5 + b * p
    x = data[:n_in].reshape(b, *info["input_shape"])
    scalars = data[n_in:n_in + b * 5].reshape(b, 5)
    policy = data[n_in + b * 5:].reshape(b, p)
    return onnx_bytes, x, scalars, policy
