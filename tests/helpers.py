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


def bf16_round(a):
    """f32 array rounded to the nearest bf16 (ties to even) and widened back -- what a bf16 operand holds."""
    import torch

    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def bf16_emulation(onnx_bytes, residual_bf16=True):
    """The oracle's op-by-op interpreter with the ARITHMETIC of a bf16 tensor-core path: every Conv reads bf16-rounded
    activations and weights and accumulates in f32; the residual stream (every Add) is stored as bf16 when
    `residual_bf16` -- the product's layout -- or kept in f32.  Separates 'the kernel computes something else' from
    'bf16 operands cannot be closer than this to the f32 oracle'.  Conv via torch's CPU kernels (f32 accumulate)."""
    from oracle.graph_exec import OnnxOracle

    class Emulation(OnnxOracle):
        def _run_node(self, node, env):
            if node.op == "Conv":
                env = dict(env)
                env[node.inputs[0]] = bf16_round(env[node.inputs[0]])
                env[node.inputs[1]] = bf16_round(env[node.inputs[1]])
                return super()._run_node(node, env)
            out = super()._run_node(node, env)
            if node.op == "Add" and residual_bf16:
                out = [bf16_round(out[0])]
            return out

    return Emulation(onnx_bytes, conv_backend="torch")


def raw_weights_from_netgen(spec, depth, weights):
    """kzero_b200.network.RawWeights (the input of kzb_net_create) from the initializers netgen.build_onnx reports through
    `weights_out` (conv-policy heads: chess conv + gather, ataxx, go)."""
    from kzero_b200 import mapping
    from kzero_b200.network import RawWeights

    a = spec.area
    conv = lambda k: (weights[f"w{k}"], weights[f"b{k}"])  # noqa: E731
    fc = lambda k: (weights[f"fc_w{k}"], weights[f"fc_b{k}"])  # noqa: E731
    k = 2 * depth + 1
    gamma, beta, mean, var = (weights[f"final_bn.{n}"].astype(np.float64) for n in ("weight", "bias", "running_mean", "running_var"))
    scale = gamma / np.sqrt(var + np.float64(np.float32(1e-5)))
    affine = (scale.astype(np.float32), (beta - scale * mean).astype(np.float32))
    extra = {}
    if spec.head == "chess_conv":
        src = mapping.chess_flat_to_conv()
    elif spec.head == "ataxx":
        src = np.concatenate([np.arange(spec.policy_size - 1), [-1]])
    elif spec.head == "go":
        src = np.concatenate([np.arange(a), [-2]])
        extra = dict(extra_conv=conv(k + 6), extra_fc=fc(k + 7))
    else:
        raise KeyError(spec.head)
    return RawWeights(spec.board_size, conv(1), [conv(i) for i in range(2, k + 1)], affine, conv(k + 1), fc(k + 2), fc(k + 3),
                      conv(k + 4), conv(k + 5), src, **extra)
