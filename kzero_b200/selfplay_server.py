"""Self-play server speaking the reference's TCP / JSON control protocol ("next" row N3), so that the
UNMODIFIED Python training loop (python/lib/loop.py via python/lib/selfplay_client.py:93-132) can drive the B200 driver.

    python -m kzero_b200.selfplay_server [--port 63105] [--device 0 --device 1 ...]

Mirrors `selfplay_server_main` (rust/kz-selfplay/src/server/server.rs:45-101): bind 127.0.0.1:<port>, accept ONE client,
read one `StartupSettings` line, then
  commander  (commander.rs:13-61)   NewSettings / NewNetwork(path) / WaitForNewNetwork / UseDummyNetwork / Stop
  collector  (collector.rs:15-116)  `{"FinishedFile": {"index": gen}}` after every `games_per_gen` games written to
                                    `<output_folder>/games_<gen>.{bin,off,json}`, `"Stopped"` at the end
Messages are one JSON value per line, externally tagged like serde's enums (protocol.rs:30-84).

Games live in a session (kzb_selfplay_session_*): a generation is one run of it, which returns after `games_per_gen` more games
have been written; the games still in flight at that point -- boards, trees, caches, recorded positions -- continue in the next
generation, so long games are never dropped (the reference's generators run across file boundaries the same way).  A new
network or new settings take effect at once, under the running games and in the middle of a file, like the reference's
(executor.rs:50-65,320-342): the commander interrupts the running session runs (kzb_selfplay_request_interrupt), which return with
their record files still open, and the next runs continue the same games into the same files with what arrived.  Differences that are deliberate (documented in DESIGN.md): `eval_random_symmetries`, `start_pos`, `top_moves`,
`saved_state_channels` and `gpu_batch_size_root` are accepted and ignored (muzero-only or above the boundary); `chess`,
`ataxx-7` and `go-9` are served with this repo's restated rules (`chess-synth` is the chess-shaped synthetic game the
B200 self-play numbers were taken with).  Several `--device` flags (server.rs:49-51,316-331): one session per device, each with its own
generator / executor threads and its own games; a generation's `games_per_gen` games are split evenly between the devices, each device
writes its share, and the parts are joined into the one `games_<gen>` file the loop expects (record_files.merge; the reference's single
collector takes games from whichever device finishes them, so there a faster device contributes more).
"""
from __future__ import annotations

import argparse
import json
import os
import socket
import threading
from pathlib import Path
from typing import Optional

from . import _abi, record_files, selfplay

DEFAULT_PORT = 63105  # server.rs:38


def parse_fpu(text: str):
    """FpuMode::from_str, rust/kz-core/src/zero/step.rs:212-226: "fixed+0.1" / "relative-0.2" -> (relative, value)."""
    for prefix, relative in (("fixed", 0), ("relative", 1)):
        if text.startswith(prefix):
            return relative, float(text[len(prefix):])
    raise ValueError(f"invalid fpu mode {text!r}")


def parse_q_mode(text: str):
    """QMode::from_str, step.rs:255-271: "value" | "wdl" | "wdl+0.0" -> (wdl, draw_score)."""
    if text == "value":
        return 0, 0.0
    if text.startswith("wdl"):
        rest = text[3:]
        return 1, float(rest) if rest else 0.0
    raise ValueError(f"invalid q mode {text!r}")


def game_id(name: str) -> int:
    """Game::parse + the per-game dispatch of server.rs:103-199, for the games this driver bundles."""
    if name == "chess":
        return selfplay.GAME_CHESS
    if name == "chess-synth":  # the chess-shaped synthetic game the B200 self-play numbers were taken with
        return selfplay.GAME_SYNTH_CHESS
    if name in ("ataxx", "ataxx-7"):
        return selfplay.GAME_ATAXX7
    if name == "go-9":  # the 4-plane encoding python/lib/games.py declares (what the Python loader and trainer read)
        return selfplay.GAME_GO9
    if name == "go-9-territory":  # the 7-plane encoding the reference's Rust server constructs (server.rs:193)
        return selfplay.GAME_GO9_TERRITORY
    raise ValueError(f"game {name!r} is not available in this driver (chess, ataxx-7, go-9, go-9-territory and the chess-shaped synthetic game 'chess-synth' are)")


def config_from(startup: dict, settings: dict, seed: int) -> _abi.SelfplayConfig:
    """StartupSettings (protocol.rs:11-28) + Settings (protocol.rs:86-112) -> kzb_selfplay_config."""
    w = settings.get("weights") or {}
    fpu_root_rel, fpu_root = parse_fpu(settings["search_fpu_root"])
    fpu_child_rel, fpu_child = parse_fpu(settings["search_fpu_child"])
    q_wdl, draw_score = parse_q_mode(settings["q_mode"])
    kw = dict(
        game=game_id(startup["game"]), visits=int(settings["full_iterations"]), part_iterations=int(settings["part_iterations"]),
        full_search_prob=float(settings["full_search_prob"]), search_batch=int(startup["search_batch_size"]),
        gpu_batch=int(startup["gpu_batch_size"]), cpu_threads=int(startup["cpu_threads_per_device"]),
        gpu_threads=int(startup["gpu_threads_per_device"]),
        max_game_length=int(settings["max_game_length"]) if settings.get("max_game_length") is not None else 2 ** 31 - 1,
        cache_size=int(settings["cache_size"]), zero_temp_move_count=int(settings["zero_temp_move_count"]),
        temperature=float(settings["temperature"]), dirichlet_alpha=float(settings["dirichlet_alpha"]),
        dirichlet_eps=float(settings["dirichlet_eps"]), policy_temperature_root=float(settings["search_policy_temperature_root"]),
        policy_temperature_child=float(settings["search_policy_temperature_child"]), fpu_root=fpu_root, fpu_root_relative=fpu_root_rel,
        fpu_child=fpu_child, fpu_child_relative=fpu_child_rel, virtual_loss=float(settings["search_virtual_loss_weight"]),
        q_mode_wdl=q_wdl, draw_score=draw_score, max_games=int(startup["games_per_gen"]), duration_s=1e9, seed=seed)
    for name in ("exploration_weight", "moves_left_weight", "moves_left_clip", "moves_left_sharpness"):
        if w.get(name) is not None:  # Weights::to_uct, protocol.rs:122-132: None -> UctWeights::default
            kw[name] = float(w[name])
    return selfplay.default_config(**kw)


class SelfplayServer:
    def __init__(self, port: int = DEFAULT_PORT, device: int = 0, devices=None):
        self.port = port
        self.devices = list(devices) if devices else [device]
        self.lock = threading.Condition()
        self.settings: Optional[dict] = None
        self.network = None  # None: wait (WaitForNewNetwork); "dummy"; or ONNX bytes
        self.stop = False
        self.listener = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.listener.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.listener.bind(("127.0.0.1", port))  # server.rs:58
        self.listener.listen(1)
        self.port = self.listener.getsockname()[1]

    # -- commander.rs:13-61 -----------------------------------------------------------------------
    def _commander(self, reader):
        for line in reader:
            line = line.strip()
            if not line:
                continue
            cmd = json.loads(line)
            with self.lock:
                if cmd == "Stop":
                    self.stop = True
                    _abi.lib().kzb_selfplay_request_stop()
                elif cmd == "WaitForNewNetwork":
                    self.network = None
                    _abi.lib().kzb_selfplay_request_interrupt()
                elif cmd == "UseDummyNetwork":
                    self.network = "dummy"
                    _abi.lib().kzb_selfplay_request_interrupt()
                elif isinstance(cmd, dict) and "NewSettings" in cmd:
                    self.settings = cmd["NewSettings"]
                    _abi.lib().kzb_selfplay_request_interrupt()
                elif isinstance(cmd, dict) and "NewNetwork" in cmd:
                    self.network = Path(cmd["NewNetwork"]).read_bytes()  # load_graph, server_alphazero.rs:126-128
                    # put it to work at once, under the running games and into the running file (executor.rs:50-65,320-342)
                    _abi.lib().kzb_selfplay_request_interrupt()
                elif isinstance(cmd, dict) and "StartupSettings" in cmd:
                    raise RuntimeError("Already received startup settings")  # commander.rs:30
                else:
                    raise ValueError(f"unknown command {cmd!r}")
                self.lock.notify_all()
                if self.stop:
                    return
        with self.lock:  # client went away
            self.stop = True
            _abi.lib().kzb_selfplay_request_stop()
            self.lock.notify_all()

    # -- server.rs:45-101 + collector.rs:15-116 --------------------------------------------------------
    def serve(self):
        conn, _ = self.listener.accept()
        reader = conn.makefile("r")
        first = json.loads(reader.readline())
        startup = first["StartupSettings"]  # server.rs:64
        if startup.get("muzero"):
            raise ValueError("MuZero is not working in the reference either (Readme.md:73) and is not built here")
        os.makedirs(startup["output_folder"], exist_ok=True)
        threading.Thread(target=self._commander, args=(reader,), daemon=True).start()

        def send(message):
            conn.sendall((json.dumps(message) + "\n").encode())

        gen = int(startup["first_gen"])
        sessions = [selfplay.Session(game_id(startup["game"])) for _ in self.devices]
        n_dev = len(self.devices)
        games_per_gen = int(startup["games_per_gen"])
        if games_per_gen < n_dev:
            raise ValueError(f"games_per_gen = {games_per_gen} cannot be split over {n_dev} devices")
        quotas = [games_per_gen // n_dev + (1 if d < games_per_gen % n_dev else 0) for d in range(n_dev)]
        file_done = [False] * n_dev  # devices whose share of the current generation is written
        totals = [0, 0, 0, 0.0]      # games, moves, nodes, seconds of the current generation (over its runs)
        try:
            while True:
                with self.lock:
                    while not self.stop and (self.settings is None or self.network is None):
                        self.lock.wait()
                    if self.stop:
                        break
                    settings, network = dict(self.settings), self.network
                    # under the lock the commander also takes: a Stop (or a newer network) can no longer slip between this check and the run
                    _abi.lib().kzb_selfplay_clear_stop()
                    _abi.lib().kzb_selfplay_clear_interrupt()
                out_prefix = str(Path(startup["output_folder"]) / f"games_{gen}")
                results, errors = [None] * n_dev, []

                def run_device(d):
                    try:
                        cfg = config_from(startup, settings, seed=int(startup["first_gen"]) * 1000 + d)
                        cfg.max_games = quotas[d]
                        cfg.output_prefix = (out_prefix if n_dev == 1 else f"{out_prefix}.dev{d}").encode()
                        if network == "dummy":
                            cfg.dummy_network = 1
                        results[d] = sessions[d].run(None if network == "dummy" else network, cfg, device=self.devices[d])
                    except Exception as e:  # noqa: BLE001 -- reported below, after the other devices have been stopped
                        errors.append(e)
                        _abi.lib().kzb_selfplay_request_stop()

                todo = [d for d in range(n_dev) if not file_done[d]]
                if len(todo) == 1:
                    run_device(todo[0])
                else:
                    threads = [threading.Thread(target=run_device, args=(d,)) for d in todo]
                    for t in threads:
                        t.start()
                    for t in threads:
                        t.join()
                if errors:
                    raise errors[0]
                if self.stop:
                    break
                for d in todo:
                    r = results[d]
                    totals[1] += r.moves_played
                    totals[2] += r.real_evals + r.cached_evals
                    file_done[d] = not r.interrupted
                totals[3] += max(results[d].seconds for d in todo)
                if not all(file_done):
                    # a new network / new settings arrived in the middle of the generation: the runs returned with their files open and
                    # the next round continues them -- same games, same files -- with what arrived
                    continue
                if n_dev > 1:
                    record_files.merge([f"{out_prefix}.dev{d}" for d in range(n_dev)], out_prefix)
                games = json.loads(Path(out_prefix + ".json").read_text())["game_count"]
                print(f"generation {gen}: {games} games, {totals[1]} moves, {totals[2] / max(totals[3], 1e-9):,.0f} nodes/s on {n_dev} device(s)",
                      flush=True)
                send({"FinishedFile": {"index": gen}})  # ServerUpdate::FinishedFile, protocol.rs:80-84
                gen += 1
                file_done = [False] * n_dev
                totals = [0, 0, 0, 0.0]
        finally:
            for session in sessions:
                session.close()
            try:
                send("Stopped")
            except OSError:
                pass
            conn.close()
            self.listener.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--port", type=int, default=DEFAULT_PORT)
    ap.add_argument("--device", type=int, action="append", help="may be given several times (server.rs:49-51); default: device 0")
    args = ap.parse_args()
    SelfplayServer(args.port, devices=args.device or [0]).serve()


if __name__ == "__main__":
    main()
