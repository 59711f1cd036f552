"""include/kzb200.hpp -- the C++ mirror of kz-core's `Network` interface over the C ABI (the reference's host side is compiled code;
INTEGRATION.md holds the same shim in Rust) -- compiled with g++ and driven by tests/cpp/network_mirror_test.cpp the way the
reference's executor drives `CudaNetwork`: boards in, one `ZeroEvaluation` per board out.

Without a GPU the program must fail loudly while it creates the network (no CPU fallback); on a B200 its evaluations must equal, bit
for bit, what the Python mirror returns for the same packed records, and its errors must be the reference's panics."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from kzero_b200 import netgen
from kzero_b200.build import LIB, build

ROOT = Path(__file__).resolve().parent.parent


def _compile(tmp_path):
    build()
    exe = tmp_path / "network_mirror_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-o", str(exe), str(ROOT / "tests" / "cpp" / "network_mirror_test.cpp"), str(LIB),
                    f"-Wl,-rpath,{LIB.parent}"], check=True)
    return exe


def test_cpp_mirror_compiles_and_fails_loudly_without_a_device(tmp_path):
    import torch

    exe = _compile(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: the GPU test runs the program")
    spec = netgen.game_spec("ataxx-7")
    onnx = tmp_path / "net.onnx"
    onnx.write_bytes(netgen.build_onnx(spec, 1, 16, seed=2))
    out = subprocess.run([str(exe), str(onnx), "ataxx-7", "3", "1", "4"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 3, out.stdout + out.stderr
    lines = out.stdout.strip().split("\n")
    assert lines[0] == "devices 0"
    assert lines[-1].startswith("error ") and ("no CUDA device" in lines[-1] or "CPU fallback" in lines[-1])


@pytest.mark.gpu
@pytest.mark.parametrize("game", ["chess", "go-9", "ataxx-7"])
def test_cpp_mirror_equals_the_python_mirror(tmp_path, game):
    from kzero_b200.network import B200Network, mapper_for

    exe = _compile(tmp_path)
    spec = netgen.game_spec(game)
    onnx_bytes = netgen.build_onnx(spec, 2, 32, seed=17)
    onnx = tmp_path / "net.onnx"
    onnx.write_bytes(onnx_bytes)
    n, max_batch = 12, 16
    out = subprocess.run([str(exe), str(onnx), game, str(n), "5", str(max_batch)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().split("\n")
    assert lines[1] == f"max_batch_size {max_batch}"
    boards, i = [], 2
    while lines[i].startswith("board "):
        rec = {"done": int(lines[i].split()[3])}
        for key, line in zip(("bits", "scalars", "indices", "values", "policy"), lines[i + 1:i + 6]):
            assert line.split()[0] == key
            rec[key] = line.split()[1:]
        boards.append(rec)
        i += 6
    assert len(boards) == n
    tail = dict(line.split(" ", 1) for line in lines[i:])
    assert tail["single_equals_row"] == "1"
    assert tail["too_many"].startswith("error batch size 17 exceeds max_batch_size 16")  # cudnn.rs:58
    assert tail["empty"] == "0"
    assert tail["mismatch"].startswith("error Input shape mismatch between graph and mapper")  # common.rs:171-174
    # the same records through the Python mirror: the C++ shim must have sent exactly these and returned exactly that
    bits = np.array([[int(v) for v in b["bits"]] for b in boards], np.uint8)
    scalars = np.array([[float(v) for v in b["scalars"]] for b in boards], np.float32)
    mv_idx = np.array([int(v) for b in boards for v in b["indices"]], np.uint32)
    mv_off = np.concatenate([[0], np.cumsum([len(b["indices"]) for b in boards])]).astype(np.uint32)
    with B200Network(mapper_for(spec), onnx_bytes, max_batch) as net:
        values, probs = net.evaluate_packed(bits, scalars, mv_idx, mv_off)
    for k, b in enumerate(boards):
        policy = np.array([float(v) for v in b["policy"]], np.float32)
        assert len(policy) == len(b["indices"]) and (len(policy) == 0) == bool(b["done"])
        assert np.array_equal(policy, probs[mv_off[k]:mv_off[k + 1]])
        assert np.array_equal(np.array([float(v) for v in b["values"]], np.float32), values[k])
        if len(policy):
            assert abs(float(policy.sum()) - 1.0) < 1e-4
            assert len(set(b["indices"])) == len(b["indices"]) and max(int(v) for v in b["indices"]) < spec.policy_size
    assert any(len(b["policy"]) > 0 for b in boards)
