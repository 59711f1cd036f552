"""Op-by-op f32 interpreter of an ONNX graph -- ORACLE / TEST INFRASTRUCTURE ONLY.

Restates what the reference's CPU path does: `CPUNetwork::evaluate_batch`
(rust/kz-core/src/network/cpu.rs:33-76) hands the NCHW f32 input to
`kn_graph::cpu::cpu_eval_graph_exec` (kn-graph 0.7.3, un-vendored crates.io dependency pinned in
rust/Cargo.toml:48 / rust/Cargo.lock:1342-1345), which evaluates the graph value by value in f32.
kn-graph's source is not in the reference tree, so this follows the ONNX operator definitions for
exactly the ops the reference's exporter emits, with the heavy loops (Conv,
Gemm) in plain C (kz_oracle.c).

No BN folding, no fusion: every node is evaluated as written, which is what makes it an
independent check of the product's pattern matcher + folded/fused kernels.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from . import conv2d, gemm
from .onnx_min import Model, load_model


class OnnxOracle:
    def __init__(self, onnx_bytes: bytes, conv_backend: str = "c"):
        """conv_backend "c": the straight C/OpenMP loops of kz_oracle.c (the restatement of the reference's naive CPU
        executor); "torch": the same op-by-op interpreter with Conv / Gemm handed to PyTorch's CPU kernels (oneDNN) --
        a stronger CPU baseline that bench.py reports beside the port, never used as the parity oracle."""
        assert conv_backend in ("c", "torch")
        self.conv_backend = conv_backend
        self.model: Model = load_model(onnx_bytes)
        assert len(self.model.inputs) == 1, "reference nets have exactly one input (network/common.rs:166-168)"
        self.input_name, self.input_shape = self.model.inputs[0]
        self.output_names = [n for n, _ in self.model.outputs]

    # ------------------------------------------------------------------ ops
    def _run_node(self, node, env: Dict[str, np.ndarray]):
        op = node.op
        a = node.attrs
        x = [env[i] if i != "" else None for i in node.inputs]
        if op == "Conv":
            k = a["kernel_shape"]
            pads = a.get("pads", [0, 0, 0, 0])
            assert a.get("group", 1) == 1 and list(a.get("strides", [1, 1])) == [1, 1]
            assert list(a.get("dilations", [1, 1])) == [1, 1]
            assert k[0] == k[1] and len(set(pads)) == 1
            if self.conv_backend == "torch":
                import warnings

                import torch

                with torch.no_grad(), warnings.catch_warnings():
                    warnings.simplefilter("ignore")  # read-only numpy views of the initializers are only read
                    y = torch.nn.functional.conv2d(torch.from_numpy(np.ascontiguousarray(x[0])), torch.from_numpy(np.ascontiguousarray(x[1])),
                                                   torch.from_numpy(np.ascontiguousarray(x[2])) if len(x) > 2 else None, padding=pads[0])
                return [y.numpy()]
            return [conv2d(x[0], x[1], x[2] if len(x) > 2 else None, pads[0])]
        if op == "Relu":
            return [np.maximum(x[0], np.float32(0))]
        if op == "Add":
            return [(x[0] + x[1]).astype(np.result_type(x[0], x[1]))]
        if op == "Sub":
            return [x[0] - x[1]]
        if op == "Mul":
            return [x[0] * x[1]]
        if op == "Div":
            return [(x[0] / x[1]).astype(x[0].dtype)]
        if op == "Sqrt":
            return [np.sqrt(x[0])]
        if op == "Tanh":
            return [np.tanh(x[0])]
        if op == "BatchNormalization":
            eps = np.float32(a.get("epsilon", 1e-5))
            scale, bias, mean, var = x[1:5]
            shp = (1, -1) + (1,) * (x[0].ndim - 2)
            y = (x[0] - mean.reshape(shp)) / np.sqrt(var.reshape(shp) + eps) * scale.reshape(shp) + bias.reshape(shp)
            return [y.astype(np.float32)]
        if op == "Flatten":
            axis = a.get("axis", 1)
            s = x[0].shape
            return [x[0].reshape(int(np.prod(s[:axis], dtype=np.int64)), -1)]
        if op == "Gemm":
            assert a.get("transA", 0) == 0
            return [gemm(x[0], x[1], x[2] if len(x) > 2 else None, bool(a.get("transB", 0)),
                         float(a.get("alpha", 1.0)), float(a.get("beta", 1.0)))]
        if op == "MatMul":
            return [np.matmul(x[0], x[1]).astype(np.float32)]
        if op == "Gather":
            return [np.take(x[0], x[1].astype(np.int64), axis=a.get("axis", 0))]
        if op == "Concat":
            return [np.concatenate(x, axis=a["axis"])]
        if op == "Constant":
            return [np.asarray(a["value"])]
        if op == "Identity":
            return [x[0]]
        if op == "Shape":
            return [np.array(x[0].shape, dtype=np.int64)]
        if op == "ConstantOfShape":
            v = a.get("value")
            fill = np.asarray(v).reshape(-1)[0] if v is not None else np.float32(0)
            return [np.full(tuple(int(d) for d in x[0]), fill, dtype=np.asarray(fill).dtype)]
        if op == "Unsqueeze":
            y = x[0]
            for ax in sorted(a["axes"]):
                y = np.expand_dims(y, ax)
            return [y]
        if op == "Squeeze":
            return [np.squeeze(x[0], axis=tuple(a["axes"])) if "axes" in a else np.squeeze(x[0])]
        if op == "Reshape":
            shape = [int(d) for d in x[1]]
            shape = [x[0].shape[i] if d == 0 else d for i, d in enumerate(shape)]
            return [x[0].reshape(shape)]
        if op == "Transpose":
            return [np.transpose(x[0], a.get("perm"))]
        if op == "Slice":
            starts, ends = x[1], x[2]
            axes = x[3] if len(x) > 3 and x[3] is not None else np.arange(len(starts))
            steps = x[4] if len(x) > 4 and x[4] is not None else np.ones(len(starts), dtype=np.int64)
            sl = [slice(None)] * x[0].ndim
            for s, e, ax, st in zip(starts, ends, axes, steps):
                sl[int(ax)] = slice(int(s), int(min(e, np.iinfo(np.int64).max)), int(st))
            return [x[0][tuple(sl)]]
        if op == "Cast":
            to = {1: np.float32, 6: np.int32, 7: np.int64, 9: np.bool_}[a["to"]]
            return [x[0].astype(to)]
        if op == "Softmax":
            ax = a.get("axis", 1)
            s = x[0].shape
            flat = x[0].reshape(int(np.prod(s[:ax], dtype=np.int64)), -1)
            e = np.exp(flat - flat.max(axis=1, keepdims=True))
            return [(e / e.sum(axis=1, keepdims=True)).reshape(s).astype(np.float32)]
        raise NotImplementedError(f"oracle: ONNX op {op}")

    def run(self, x: np.ndarray, keep_all: bool = False):
        """x: [batch, C, H, W] f32 -> list of outputs in graph order (reference: scalars, policy)."""
        env: Dict[str, np.ndarray] = dict(self.model.initializers)
        env[self.input_name] = np.ascontiguousarray(x, dtype=np.float32)
        for node in self.model.nodes:
            outs = self._run_node(node, env)
            for name, val in zip(node.outputs, outs):
                env[name] = val
        if keep_all:
            return env
        return [env[n] for n in self.output_names]

    def op_counts(self) -> Dict[str, int]:
        c: Dict[str, int] = {}
        for n in self.model.nodes:
            c[n.op] = c.get(n.op, 0) + 1
        return c


def evaluate(onnx_bytes: bytes, x: np.ndarray) -> List[np.ndarray]:
    return OnnxOracle(onnx_bytes).run(x)
