"""C-ABI surface and host-side logic that needs no GPU: the library loads, exports every symbol the
header declares, recognises the reference's exports, rejects what it must, and refuses to run without CUDA."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from helpers import GOLDEN, NET_NAMES, load_net_fixture, structure
from kzero_b200 import _abi, netgen
from kzero_b200.build import build
from kzero_b200.network import B200Network, KzbError, inspect_onnx, mapper_for

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module", autouse=True)
def _built():
    build()


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "kzb200.h").read_text()
    declared = set(re.findall(r"\b(kzb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_abi.SYMBOLS), declared ^ set(_abi.SYMBOLS)
    lib = ctypes.CDLL(str(_abi.lib_path()))
    for name in declared:
        assert hasattr(lib, name), f"libkzb200.so does not export {name}"


@pytest.mark.parametrize("name", NET_NAMES)
def test_inspect_recognises_reference_exports(name):
    onnx_bytes, x, _, policy = load_net_fixture(name)
    info = inspect_onnx(onnx_bytes)
    s = structure()[name]
    assert [info.input_channels, info.board_h, info.board_w] == s["input_shape"]
    assert info.policy_len == policy.shape[1]
    depth = s["ops"].count("Add")
    assert info.depth == depth
    assert info.channels in (16, 32)


def test_inspect_flops_match_survey_formula():
    # chess 16x128 = 610.45 MFLOP / position, ataxx-7 8x64 = 58.57
    chess = inspect_onnx(netgen.build_onnx(netgen.game_spec("chess"), 16, 128))
    assert abs(chess.flops_per_position / 1e6 - 610.45) < 0.05
    ataxx = inspect_onnx(netgen.build_onnx(netgen.game_spec("ataxx-7"), 8, 64))
    assert abs(ataxx.flops_per_position / 1e6 - 58.57) < 0.05


def test_inspect_accepts_unfolded_bn():
    info = inspect_onnx(netgen.build_onnx(netgen.game_spec("ataxx-7"), 3, 16, fold_bn=False))
    assert info.depth == 3 and info.channels == 16


def test_inspect_attention_head_flops_and_size():
    """The attention policy head (post_act.py:115-141) with Q = channels, as supervised_main_alpha.py:76 builds it:
    conv_bulk C->2Q on 64 squares, conv_under C->3Q on one rank, [64,Q]x[Q,88] product."""
    c = q = 128
    info = inspect_onnx(netgen.build_onnx(netgen.game_spec("chess-att"), 2, c, query_channels=q))
    assert info.policy_len == 1880
    tower = 2 * 64 * 9 * 21 * c + 2 * 2 * (2 * 64 * 9 * c * c)
    scalar = 2 * 64 * c * 4 + 2 * 256 * 32 + 2 * 32 * 5
    att = 2 * 64 * c * 2 * q + 2 * 8 * c * 3 * q + 2 * 64 * 88 * q
    assert info.flops_per_position == pytest.approx(tower + scalar + att, rel=1e-9)


def test_inspect_rejects_unknown_policy_head_with_message():
    """A graph whose policy output is not one of the reference's heads must fail with a message, not mis-evaluate."""
    onnx_bytes, *_ = load_net_fixture("chess_conv_2x32")
    broken = onnx_bytes.replace(b"Gather", b"Gathex")  # same length: the policy head's last op becomes unknown
    assert broken != onnx_bytes
    with pytest.raises(KzbError, match="policy head|unsupported"):
        inspect_onnx(broken)


def test_inspect_rejects_garbage():
    with pytest.raises(KzbError):
        inspect_onnx(b"\x00\x01\x02 definitely not onnx")
    with pytest.raises(KzbError):
        inspect_onnx(netgen.build_onnx(netgen.game_spec("chess"), 1, 16)[:2000])


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly (never fall back to the oracle or any CPU path)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert _abi.lib().kzb_device_count() == 0
    spec = netgen.game_spec("ataxx-7")
    with pytest.raises(KzbError, match="no CUDA device|CPU fallback"):
        B200Network(mapper_for(spec), netgen.build_onnx(spec, 1, 16), 4)


def test_product_does_not_import_oracle():
    for path in (ROOT / "kzero_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cpp", ".hpp", ".cuh", ".h"):
            text = path.read_text()
            assert "import oracle" not in text and "from oracle" not in text and "kz_oracle" not in text, path


def test_net_create_from_raw_weights_checks_shapes_before_touching_a_device():
    """kzb_net_create (weights the caller already holds, so the reference's load_graph stays untouched): shape errors are reported
    with a message, and a well-formed spec gets as far as the device -- which is missing here."""
    import numpy as np
    import pytest

    from helpers import raw_weights_from_netgen
    from kzero_b200 import netgen
    from kzero_b200.network import B200Network, KzbError, mapper_for

    spec = netgen.game_spec("go-9")
    weights = {}
    netgen.build_onnx(spec, 2, 32, seed=3, weights_out=weights)
    raw = raw_weights_from_netgen(spec, 2, weights)
    assert (raw.spec.input_channels, raw.spec.channels, raw.spec.depth, raw.spec.policy_len, raw.spec.has_extra) == (10, 32, 2, 82, 1)
    with pytest.raises(KzbError, match="no CUDA device|CUDA"):  # everything host-side was accepted
        B200Network(mapper_for(spec), raw, 4)
    bad = dict(weights)
    bad["w3"] = np.zeros((32, 31, 3, 3), np.float32)
    with pytest.raises(KzbError, match="block conv: expected a 3x3 conv over 32 channels"):
        B200Network(mapper_for(spec), raw_weights_from_netgen(spec, 2, bad), 4)
    raw = raw_weights_from_netgen(spec, 2, weights)
    raw.spec.policy_src[5] = 81 * 1 + 7  # beyond the 1-channel policy map
    with pytest.raises(KzbError, match="policy_src entry"):
        B200Network(mapper_for(spec), raw, 4)


def test_three_output_graphs_are_rejected_with_a_message():
    """The legacy 3-output graph form (value [B], wdl [B, 3], policy; rust/kz-core/src/network/common.rs:43-50,181-196) is not
    produced by the reference's exporter (save_onnx.py:111-119 writes `scalars`, `policy`) and is not supported here: loading one
    fails at inspect / create time with a message that says so -- it never evaluates to something else."""
    import pytest

    from kzero_b200 import netgen
    from kzero_b200.network import KzbError, inspect_onnx

    spec = netgen.game_spec("ataxx-7")
    onnx_bytes = netgen.build_onnx(spec, 1, 16, seed=1, legacy_three_outputs=True)
    with pytest.raises(KzbError, match="3-output"):
        inspect_onnx(onnx_bytes)


def test_header_is_plain_c_and_struct_layouts_match_the_ctypes_mirror(tmp_path):
    """The boundary is a C ABI: include/kzb200.h must compile as C99 (and as C++), a C program must link against the built library
    and call it, and every struct the Python mirror declares must have the size -- and its last field the offset -- the C compiler
    gives it (a field added on one side only would silently shift everything behind it)."""
    import ctypes
    import subprocess

    from kzero_b200 import _abi
    from kzero_b200.build import LIB

    pairs = [("kzb_net_info", _abi.NetInfo), ("kzb_conv_weights", _abi.ConvWeights), ("kzb_fc_weights", _abi.FcWeights),
             ("kzb_net_spec", _abi.NetSpecC), ("kzb_selfplay_config", _abi.SelfplayConfig), ("kzb_selfplay_stats", _abi.SelfplayStats),
             ("kzb_mcts_trace_out", _abi.MctsTraceOut)]
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "kzb200.h"', 'int main(void) {']
    for cname, mirror in pairs:
        last = mirror._fields_[-1][0]
        lines.append(f'    printf("{cname} %zu %zu\\n", sizeof({cname}), offsetof({cname}, {last}));')
    lines += ['    printf("devices %d\\n", kzb_device_count());', '    printf("error [%s]\\n", kzb_last_error());', '    return 0;', '}']
    src = tmp_path / "abi_probe.c"
    src.write_text("\n".join(lines) + "\n")
    include = str(ROOT / "include")
    exe = tmp_path / "abi_probe"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", include, "-o", str(exe), str(src), str(LIB),
                    f"-Wl,-rpath,{LIB.parent}"], check=True)
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", "-I", include, str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    got = {line.split()[0]: tuple(int(v) for v in line.split()[1:3]) for line in out if line.startswith("kzb_")}
    for cname, mirror in pairs:
        last = mirror._fields_[-1][0]
        assert got[cname] == (ctypes.sizeof(mirror), getattr(mirror, last).offset), (cname, got[cname])
    assert any(line.startswith("devices ") for line in out)  # the call went through the C ABI (0 devices here, or an error code)
