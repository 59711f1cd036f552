"""Self-play record files (`<prefix>.bin / .off / .json`, rust/kz-selfplay/src/binary_output.rs:128-297) on the host side: joining the
files several devices wrote for one generation into the single file the training loop expects.

The reference's collector receives finished simulations from the generators of EVERY device and appends them to one output
(collector.rs:59-116); this driver runs one session per device, each with its own writer, so a generation's files are joined here:
`.bin` records are concatenated, position offsets are shifted by the bytes in front of them, per-game start indices by the positions
in front of them, the `game_id` scalar of every position (first f32 of its record, binary_output.rs:168-171) by the games in front of
it, and the metadata is recombined (counts summed, length extremes, game-weighted means of `root_wdl` / `hit_move_limit`).
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Sequence

import numpy as np


def merge(parts: Sequence[str], out_prefix: str, remove_parts: bool = True) -> dict:
    """parts: prefixes of finished record files of the same game; -> metadata of the joined file `<out_prefix>.{bin,off,json}`."""
    metas = [json.loads(Path(p + ".json").read_text()) for p in parts]
    for m in metas[1:]:
        for key in ("game", "input_bool_shape", "input_scalar_count", "policy_shape", "scalar_names"):
            if m[key] != metas[0][key]:
                raise ValueError(f"record files disagree on {key}: {m[key]} vs {metas[0][key]}")
    byte_base = pos_base = game_base = 0
    offsets, starts = [], []
    tmp = out_prefix + ".bin.tmp"
    with open(tmp, "wb") as out:
        for p, m in zip(parts, metas):
            n, g = int(m["position_count"]), int(m["game_count"])
            data = np.fromfile(p + ".bin", dtype=np.uint8)
            off = np.fromfile(p + ".off", dtype="<u8")
            if off.size != n + g:
                raise ValueError(f"{p}.off holds {off.size} entries, expected {n} offsets + {g} game starts")
            pos_off = off[:n].astype(np.int64)
            if game_base and n:
                idx = pos_off[:, None] + np.arange(4)
                ids = data[idx].copy().view("<f4").reshape(n) + np.float32(game_base)
                data[idx] = ids.astype("<f4").view(np.uint8).reshape(n, 4)
            out.write(data.tobytes())
            offsets.append(off[:n] + np.uint64(byte_base))
            starts.append(off[n:] + np.uint64(pos_base))
            byte_base += int(data.size)
            pos_base += n
            game_base += g
    with open(out_prefix + ".off.tmp", "wb") as f:
        f.write(np.concatenate(offsets + starts).astype("<u8").tobytes())
    games = [int(m["game_count"]) for m in metas]
    total = max(sum(games), 1)
    played = [m for m in metas if m["game_count"]]
    meta = dict(metas[0])
    meta["game_count"] = sum(games)
    meta["position_count"] = pos_base
    meta["max_game_length"] = max((m["max_game_length"] for m in played), default=-1)
    meta["min_game_length"] = min((m["min_game_length"] for m in played), default=-1)
    meta["root_wdl"] = [sum(m["root_wdl"][i] * g for m, g in zip(metas, games)) / total for i in range(3)]
    meta["hit_move_limit"] = sum(m["hit_move_limit"] * g for m, g in zip(metas, games)) / total
    Path(out_prefix + ".json.tmp").write_text(json.dumps(meta, indent=2) + "\n")
    for ext in (".bin", ".off", ".json"):  # the .json appears last: its presence says the file is complete (like the writer's rename)
        os.replace(out_prefix + ext + ".tmp", out_prefix + ext)
    if remove_parts:
        for p in parts:
            for ext in (".bin", ".off", ".json"):
                os.remove(p + ext)
    return meta
