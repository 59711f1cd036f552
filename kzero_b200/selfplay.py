"""Python host mirror of the self-play driver in libkzb200.so (include/kzb200.h, `kzb_selfplay_*`, `kzb_mcts_trace`).

Mirrors what the reference's self-play server runs per device (rust/kz-selfplay/src/server/server_alphazero.rs:32-124):
generator threads (generator_alphazero.rs:23-260) feeding executor threads (executor.rs:27-146); settings named like
the reference's `StartupSettings` / `Settings` (protocol.rs:11-110, python/lib/selfplay_client.py).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _abi
from ._abi import SelfplayConfig, SelfplayStats

GAME_SYNTH_CHESS = 0
GAME_ATAXX7 = 1
GAME_GO9 = 2
GAME_GO9_TERRITORY = 4  # go-9 with the three territory planes (GoStdMapper::new(9, true), server.rs:193)
GAME_CHESS = 3  # real chess; GAME_SYNTH_CHESS is the chess-shaped synthetic game the B200 self-play numbers were taken with


def default_config(**overrides) -> SelfplayConfig:
    """The reference's production settings (python/main/loop_main_alpha.py:24-52), with keyword overrides."""
    cfg = SelfplayConfig()
    _abi.lib().kzb_selfplay_default_config(ctypes.byref(cfg))
    names = {f[0] for f in SelfplayConfig._fields_}
    for k, v in overrides.items():
        if k not in names:
            raise KeyError(k)
        if k == "output_prefix" and isinstance(v, str):
            v = v.encode()
        setattr(cfg, k, v)
    return cfg


@dataclass
class SelfplayResult:
    seconds: float
    real_evals: int
    cached_evals: int
    potential_evals: int
    batches: int
    max_batch: int
    games_finished: int
    moves_played: int
    root_visits: int
    concurrent_games: int
    games_written: int = 0
    interrupted: int = 0  # a session run that returned on kzb_selfplay_request_interrupt with its record file still open

    @property
    def nn_positions_per_s(self) -> float:  # "real evals/s" of collector.rs:172-191
        return self.real_evals / self.seconds

    @property
    def mcts_nodes_per_s(self) -> float:  # real + cached evals per second
        return (self.real_evals + self.cached_evals) / self.seconds

    @property
    def mean_batch(self) -> float:
        return self.real_evals / max(self.batches, 1)

    @property
    def cache_hit_rate(self) -> float:
        return self.cached_evals / max(self.real_evals + self.cached_evals, 1)


def run(onnx_bytes: Optional[bytes], config: SelfplayConfig, device: int = 0, precision: int = 1) -> SelfplayResult:
    """onnx_bytes may be None when config.dummy_network is set (uniform evaluations, no GPU)."""
    stats = SelfplayStats()
    _abi.lib().kzb_selfplay_clear_stop()  # a stand-alone run starts fresh; the server manages the flag itself (kzb_selfplay_request_stop)
    _abi.check(_abi.lib().kzb_selfplay_run(device, onnx_bytes, len(onnx_bytes) if onnx_bytes else 0, precision, ctypes.byref(config),
                                           ctypes.byref(stats)))
    return SelfplayResult(**{f[0]: getattr(stats, f[0]) for f in SelfplayStats._fields_})


class Session:
    """The concurrent games of one server connection, kept alive between runs (kzb_selfplay_session_*): every `run` plays until
    config.max_games more games have finished and returns; games in flight continue in the next run, with that run's network and
    settings -- like the reference's generators, which run across file boundaries (collector.rs:59-116)."""

    def __init__(self, game: int):
        self._handle = ctypes.c_void_p()
        _abi.check(_abi.lib().kzb_selfplay_session_create(game, ctypes.byref(self._handle)))

    def run(self, onnx_bytes: Optional[bytes], config: SelfplayConfig, device: int = 0, precision: int = 1) -> SelfplayResult:
        stats = SelfplayStats()
        _abi.check(_abi.lib().kzb_selfplay_session_run(self._handle, device, onnx_bytes, len(onnx_bytes) if onnx_bytes else 0, precision,
                                                       ctypes.byref(config), ctypes.byref(stats)))
        return SelfplayResult(**{f[0]: getattr(stats, f[0]) for f in SelfplayStats._fields_})

    def close(self) -> None:
        if self._handle:
            _abi.lib().kzb_selfplay_session_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


@dataclass
class TraceResult:
    child_visits: np.ndarray
    child_moves: np.ndarray
    child_policy: np.ndarray
    root_values: np.ndarray
    root_visits: int
    tree_nodes: int
    evals: int


def mcts_trace(config: SelfplayConfig, game_seed: int, plies: int, eval_kind: int, capacity: int = 4096) -> TraceResult:
    """Host-only search trace (no GPU): see kzb_mcts_trace in include/kzb200.h."""
    visits = np.zeros(capacity, np.uint64)
    moves = np.zeros(capacity, np.uint32)
    policy = np.zeros(capacity, np.float32)
    out = _abi.MctsTraceOut()
    out.capacity = capacity
    out.child_visits = visits.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
    out.child_moves = moves.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
    out.child_policy = policy.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    _abi.check(_abi.lib().kzb_mcts_trace(ctypes.byref(config), game_seed, plies, eval_kind, ctypes.byref(out)))
    n = out.n_children
    return TraceResult(visits[:n].copy(), moves[:n].copy(), policy[:n].copy(), np.array(list(out.root_values), np.float32),
                       int(out.root_visits), int(out.tree_nodes), int(out.evals))
