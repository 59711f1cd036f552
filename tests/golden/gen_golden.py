#!/usr/bin/env python
"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Runs only in the build container (needs /root/reference); the fixtures it writes are committed and
are what travels to the GPU box.  Re-run with:  python tests/golden/gen_golden.py

What it produces (all from the reference's own Python code, imported from /root/reference/python):
  net_<name>.onnx + net_<name>.bin   small random-init nets built from the reference's model classes
        (python/lib/model/post_act.py) and exported with the reference's own export arguments
        (python/lib/save_onnx.py:111-119); the .bin is the reference's own check-file format
        (save_onnx.py:94-102): 1 byte batch size, raw f32 inputs, raw f32 PyTorch outputs.
  planes_<game>.npz                   packed (bits, scalars) records and the f32 planes the reference's
        Python decoder produces for them (python/lib/data/position.py:94-98 unpackbits little +
        position.py:267-271 write_input).
  export_structure.json               op sequence of each export (pins kzero_b200/netgen.py's structure)
"""
import io
import json
import sys
import types
import warnings
from pathlib import Path

import numpy as np

REF = "/root/reference/python"
sys.path.insert(0, REF)

import torch  # noqa: E402
from torch import nn  # noqa: E402
import torch.onnx._internal.torchscript_exporter.onnx_proto_utils as _opu  # noqa: E402

# the legacy exporter's last step imports the (absent) `onnx` package only to splice onnxscript
# functions; there are none in these models
_opu._add_onnxscript_fn = lambda b, c: b

from lib.games import Game  # noqa: E402
from lib.data.position import write_input  # noqa: E402
from lib.mapping.mapping import CHESS_FLAT_TO_CONV  # noqa: E402
from lib.model.post_act import (ResTower, ScalarHead, PredictionHeads, AtaxxConvPolicyHead,  # noqa: E402
                                ConvPolicyHead, AttentionPolicyHead, conv2d)

OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(OUT.parent.parent))
from oracle.onnx_min import load_model  # noqa: E402


class LegacyChessConvPolicyHead(ConvPolicyHead):
    """conv1x1 -> relu -> conv1x1(73) -> flatten -> Gather(CHESS_FLAT_TO_CONV): post_act.py:69-88.
    HEAD's __init__ asserts policy_shape == (73*64,) (post_act.py:60) so it cannot be constructed for
    `chess` (policy 1880) directly; this subclass only bypasses that assert, forward() is inherited."""

    def __init__(self, game: Game, channels: int):
        nn.Module.__init__(self)
        self.extra_moves = 0
        self.seq = nn.Sequential(conv2d(channels, channels, 1), nn.ReLU(),
                                 conv2d(channels, game.policy_conv_channels, 1))
        self.flatten_indices = CHESS_FLAT_TO_CONV


def perturb_bn(net: nn.Module, gen: torch.Generator):
    """Make BN folding non-trivial (cf. python/main/write_test_networks.py:26-37, which trains a few
    steps for the same reason).  Distributions: SURVEY.md 8(d)."""
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d):
            c = m.num_features
            m.running_mean.copy_(torch.randn(c, generator=gen) * 0.1)
            m.running_var.copy_(torch.rand(c, generator=gen) + 0.5)
            if m.affine:
                m.weight.data.copy_(torch.rand(c, generator=gen) + 0.5)
                m.bias.data.copy_(torch.randn(c, generator=gen) * 0.1)


def export(net: nn.Module, game: Game, name: str, batch: int, structure: dict):
    net.eval()
    x = torch.randn(batch, *game.full_input_shape)
    with torch.no_grad():
        outs = net(x)
    f = io.BytesIO()
    names_in, names_out = ["input"], ["scalars", "policy"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.onnx.export(model=net, args=(x,), f=f, input_names=names_in, output_names=names_out,
                          dynamic_axes={k: {0: "batch_size"} for k in names_in + names_out},
                          opset_version=10, dynamo=False)
    (OUT / f"net_{name}.onnx").write_bytes(f.getvalue())
    with open(OUT / f"net_{name}.bin", "wb") as fb:  # save_onnx.py:94-102
        fb.write(batch.to_bytes(1, byteorder="little", signed=False))
        fb.write(x.numpy().tobytes())
        for o in outs:
            fb.write(o.cpu().detach().numpy().tobytes())
    model = load_model(f.getvalue())
    structure[name] = {
        "game": game.name,
        "input_shape": list(game.full_input_shape),
        "policy_shape": list(game.policy_shape),
        "batch": batch,
        "ops": [n.op for n in model.nodes],
        "onnx_bytes": len(f.getvalue()),
    }
    print(name, len(f.getvalue()), "bytes", {o: structure[name]["ops"].count(o) for o in set(structure[name]["ops"])})


def make_nets():
    gen = torch.Generator().manual_seed(0)
    torch.manual_seed(0)
    structure = {}
    depth, ch = 2, 32

    g = Game.find("ataxx-7")
    net = PredictionHeads(ResTower(depth, g.full_input_channels, ch), ScalarHead(g.board_size, ch, 4, 32),
                          AtaxxConvPolicyHead(g, ch))
    perturb_bn(net, gen)
    export(net, g, "ataxx7_2x32", 5, structure)

    g = Game.find("chess")
    net = PredictionHeads(ResTower(depth, g.full_input_channels, ch), ScalarHead(g.board_size, ch, 4, 32),
                          LegacyChessConvPolicyHead(g, ch))
    perturb_bn(net, gen)
    export(net, g, "chess_conv_2x32", 4, structure)

    net = PredictionHeads(ResTower(depth, g.full_input_channels, ch), ScalarHead(g.board_size, ch, 4, 32),
                          AttentionPolicyHead(g, ch, 16))
    perturb_bn(net, gen)
    export(net, g, "chess_att_2x32", 3, structure)

    g = Game.find("go-9")
    net = PredictionHeads(ResTower(depth, g.full_input_channels, ch), ScalarHead(g.board_size, ch, 4, 32),
                          ConvPolicyHead(g, ch, extra_moves=1))
    perturb_bn(net, gen)
    export(net, g, "go9_2x32", 3, structure)

    # a scripted module, the form the training loop actually exports (python/lib/loop.py:198,299)
    g = Game.find("ataxx-5")
    net = PredictionHeads(ResTower(1, g.full_input_channels, 16), ScalarHead(g.board_size, 16, 4, 32),
                          AtaxxConvPolicyHead(g, 16))
    perturb_bn(net, gen)
    net.eval()
    net = torch.jit.script(net)
    export(net, g, "ataxx5_scripted_1x16", 2, structure)

    (OUT / "export_structure.json").write_text(json.dumps(structure, indent=1))


def make_planes():
    """Packed records -> planes via the reference's Python decode (position.py:94-98, 267-271)."""
    rng = np.random.default_rng(0)
    for name in ["chess", "ataxx-7", "go-9", "ataxx-3", "go-19", "ttt", "arimaa-split"]:
        game = Game.find(name)
        n = 6
        cb, h, w = game.input_bool_shape
        cs = game.input_scalar_channels
        bool_count = cb * h * w
        nbytes = (bool_count + 7) // 8
        bits = rng.integers(0, 256, size=(n, nbytes), dtype=np.uint8)
        # one all-zero and one all-one record
        bits[0] = 0
        bits[1] = 255
        scalars = rng.standard_normal((n, cs)).astype(np.float32)
        scalars[0] = 0
        planes = np.zeros((n, *game.full_input_shape), dtype=np.float32)
        for i in range(n):
            bool_buffer = np.unpackbits(bits[i], bitorder="little")  # position.py:95
            pos = types.SimpleNamespace(
                input_bools=bool_buffer[:bool_count].reshape(*game.input_bool_shape),  # position.py:96
                input_scalars=scalars[i],
            )
            target = torch.zeros(*game.full_input_shape)
            write_input(game, target, pos)  # position.py:267-271
            planes[i] = target.numpy()
        np.savez_compressed(OUT / f"planes_{name}.npz", bits=bits, scalars=scalars, planes=planes,
                            bool_shape=np.array(game.input_bool_shape), scalar_count=np.array(cs))
        print("planes", name, planes.shape)


if __name__ == "__main__":
    make_nets()
    make_planes()
