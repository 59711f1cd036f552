"""Minimal ONNX protobuf reader (no `onnx` package) -- ORACLE / TEST INFRASTRUCTURE ONLY.

Independent of the product's C++ reader (kzero_b200/csrc/onnx_reader.cpp) on purpose: the two
implementations cross-check each other in tests/.

Follows the ONNX wire format as exported by the reference (python/lib/save_onnx.py:111-119,
opset 10).  Field numbers are from onnx.proto3: ModelProto.graph=7; GraphProto.node=1,
initializer=5, input=11, output=12; NodeProto.input=1, output=2, name=3, op_type=4, attribute=5;
AttributeProto.name=1, f=2, i=3, s=4, t=5, floats=7, ints=8; TensorProto.dims=1, data_type=2,
float_data=4, int32_data=5, int64_data=7, name=8, raw_data=9.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf: bytes):
    """Yield (field_number, wire_type, value) for one message; value is int or bytes."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, v


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(v, wt) -> List[int]:
    if wt == 0:
        return [_signed64(v)]
    out = []
    pos = 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(_signed64(x))
    return out


_DTYPES = {1: np.float32, 6: np.int32, 7: np.int64, 9: np.bool_, 11: np.float64}


def parse_tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype = 1
    name = ""
    raw = None
    floats: List[float] = []
    ints: List[int] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims += _packed_varints(v, wt)
        elif fno == 2:
            dtype = v
        elif fno == 4:
            if wt == 5:
                floats.append(struct.unpack("<f", v)[0])
            else:
                floats += list(np.frombuffer(v, dtype="<f4"))
        elif fno in (5, 7):
            ints += _packed_varints(v, wt)
        elif fno == 8:
            name = v.decode()
        elif fno == 9:
            raw = bytes(v)
    np_dtype = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(np_dtype).newbyteorder("<")).astype(np_dtype)
    elif dtype == 1:
        arr = np.array(floats, dtype=np.float32)
    else:
        arr = np.array(ints, dtype=np_dtype)
    return name, arr.reshape(dims)


@dataclass
class Node:
    op: str
    inputs: List[str]
    outputs: List[str]
    attrs: Dict[str, object] = field(default_factory=dict)
    name: str = ""


def _parse_attr(buf: bytes):
    name = ""
    val = None
    ints: List[int] = []
    floats: List[float] = []
    have_ints = have_floats = False
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = v.decode()
        elif fno == 2:
            val = struct.unpack("<f", v)[0]
        elif fno == 3:
            val = _signed64(v)
        elif fno == 4:
            val = bytes(v)
        elif fno == 5:
            val = parse_tensor(v)[1]
        elif fno == 7:
            have_floats = True
            if wt == 5:
                floats.append(struct.unpack("<f", v)[0])
            else:
                floats += list(np.frombuffer(v, dtype="<f4"))
        elif fno == 8:
            have_ints = True
            ints += _packed_varints(v, wt)
    if have_ints:
        val = ints
    elif have_floats:
        val = floats
    return name, val


def _parse_node(buf: bytes) -> Node:
    node = Node("", [], [])
    for fno, wt, v in _fields(buf):
        if fno == 1:
            node.inputs.append(v.decode())
        elif fno == 2:
            node.outputs.append(v.decode())
        elif fno == 3:
            node.name = v.decode()
        elif fno == 4:
            node.op = v.decode()
        elif fno == 5:
            k, a = _parse_attr(v)
            node.attrs[k] = a
    return node


def _parse_value_info(buf: bytes):
    """-> (name, shape) where shape entries are int or str (dim_param) or None."""
    name = ""
    shape = None
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = v.decode()
        elif fno == 2:  # TypeProto
            for f2, _, v2 in _fields(v):
                if f2 == 1:  # tensor_type
                    for f3, _, v3 in _fields(v2):
                        if f3 == 2:  # shape
                            shape = []
                            for f4, _, v4 in _fields(v3):
                                if f4 == 1:  # dim
                                    d = None
                                    for f5, _, v5 in _fields(v4):
                                        if f5 == 1:
                                            d = _signed64(v5)
                                        elif f5 == 2:
                                            d = v5.decode()
                                    shape.append(d)
    return name, shape


@dataclass
class Model:
    nodes: List[Node]
    initializers: Dict[str, np.ndarray]
    inputs: List[Tuple[str, list]]
    outputs: List[Tuple[str, list]]


def load_model(data: bytes) -> Model:
    graph = None
    for fno, wt, v in _fields(data):
        if fno == 7:
            graph = v
    if graph is None:
        raise ValueError("no graph in ModelProto")
    nodes: List[Node] = []
    inits: Dict[str, np.ndarray] = {}
    inputs = []
    outputs = []
    for fno, wt, v in _fields(graph):
        if fno == 1:
            nodes.append(_parse_node(v))
        elif fno == 5:
            n, arr = parse_tensor(v)
            inits[n] = arr
        elif fno == 11:
            inputs.append(_parse_value_info(v))
        elif fno == 12:
            outputs.append(_parse_value_info(v))
    inputs = [(n, s) for n, s in inputs if n not in inits]
    return Model(nodes, inits, inputs, outputs)
