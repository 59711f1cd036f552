#!/usr/bin/env python
"""bench.py -- NN positions/sec of the self-play inference hot path on B200 (contract: see the task prompt).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config chess|ataxx|go9]

A "step" = one pass of the hot path over one batch of synthetic positions: packed (bits, scalars) ->
plane encoding -> conv tower -> heads -> legal-move-masked softmax.  Default workload = BASELINE.json
configs[1]: chess 8x8 ResNet 16x128, kz-core chess mapping (21 planes in, 1880-move policy), batch 1024, bf16
tensor-core path.  Multi-GPU = independent replicas sharded by game (no collective on the data path), so
--gpus N under torchrun is weak scaling: every rank evaluates its own batches.

Printed JSON (one line, rank 0):
  value      positions/s, whole job, device-timed (CUDA events on the network's stream), inputs resident in HBM
  e2e        positions/s through the public call (kzb_eval_packed via B200Network.evaluate_packed) with HOST
             buffers: H2D of the packed batch and D2H of values + legal-move probabilities inside the timed region
  roofline   dominant kernel (the conv tower) algorithmic TFLOP/s vs the measured bf16 peak
  cpu_baseline  the oracle (CPU restatement of the reference's CPU executor) on a bounded sample, host cores
`--impl reference` times that CPU restatement (the reference's own crates cannot be built here: no Rust
toolchain, kn-graph is an un-vendored crates.io dependency) on all host threads, same workload and metric.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from kzero_b200 import netgen  # noqa: E402

CONFIGS = {
    # BASELINE.json configs[1] -- the configuration the metric is quoted on
    "chess": dict(game="chess", depth=16, channels=128, batch=1024,
                  workload="chess 8x8 ResNet 16x128 (kz-core chess mapping: 21x8x8 planes, 1880-move policy, "
                           "conv policy head), batch 1024"),
    # BASELINE.json configs[0] -- the reference's own CPU-runnable case
    "ataxx": dict(game="ataxx-7", depth=8, channels=64, batch=256,
                  workload="ataxx 7x7 ResNet 8x64, batch 256"),
    "go9": dict(game="go-9", depth=20, channels=256, batch=4096,
                workload="go 9x9 ResNet 20x256, batch 4096"),
}
N_INPUT_SETS = 4  # distinct synthetic batches rotated through the e2e loop


def measured_peaks():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        d = json.loads(path.read_text())
        return dict(tflops_sustained=float(d["bf16_tflops_sustained"]), tflops_burst=float(d["bf16_tflops"]),
                    hbm_gbs=float(d["hbm_gbs"]), source="MEASURED_PEAKS.json")
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return dict(tflops_sustained=1400.0, tflops_burst=1590.0, hbm_gbs=6650.0, source="fallback")


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: an NVML polling thread (a 50-step chess run
    lasts tens of milliseconds, far below nvidia-smi's -lms resolution), nvidia-smi as the fallback."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
               "hw_power_brake_slowdown": 0x80}

    def __init__(self, device: int):
        self.device = device
        self.samples = []  # (sm_mhz, power_w, reasons bitmask)
        self.sm_max = None
        self.stop_flag = threading.Event()
        self.thread = None
        self.source = None

    def _handle(self):
        import pynvml
        import torch

        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(self.device).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:  # noqa: BLE001
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = self.device
            if visible:
                try:
                    index = int(visible.split(",")[self.device])
                except ValueError:
                    pass
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)

    def start(self):
        try:
            nv, h = self._handle()
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.source = "nvml"

            def poll():
                while not self.stop_flag.is_set():
                    try:
                        sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                        pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                        rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                        self.samples.append((sm, pw, rs))
                    except Exception:  # noqa: BLE001
                        pass
                    time.sleep(0.002)

            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.source = None

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader,nounits",
                                  "-i", str(self.device)], capture_output=True, text=True, timeout=20).stdout.strip().split(",")
            return float(out[0]), float(out[1]), float(out[2])
        except Exception:  # noqa: BLE001
            return None

    def stop(self):
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.samples:
            one = self._smi_once()
            if one is None:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source available"]}
            return {"sm_mhz": one[0], "sm_max_mhz": one[1], "reasons": [], "samples": 1, "power_w_max": one[2],
                    "source": "nvidia-smi (single query right after the timed region)"}
        sm = [s[0] for s in self.samples]
        power = [s[1] for s in self.samples]
        busy = [s for s, pw in zip(sm, power) if pw >= 0.5 * max(power)] or sm
        mask = 0
        for s in self.samples:
            mask |= s[2]
        reasons = sorted(name for name, bit in self.REASONS.items() if mask & bit)
        return {"sm_mhz": float(np.median(busy)), "sm_min_mhz": float(min(busy)), "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(sm), "power_w_max": float(max(power)), "source": "nvml, polled every ~2 ms during the timed regions"}


class CpuRestatement:
    """The oracle (expand planes -> ONNX graph in f32 -> decode_output) as a timed CPU implementation of the path."""

    def __init__(self, onnx_bytes, spec, threads: int, conv_backend: str = "c"):
        import oracle
        from oracle.graph_exec import OnnxOracle

        self.oracle = oracle
        self.spec = spec
        oracle.set_threads(threads)
        if conv_backend == "torch":
            import torch

            torch.set_num_threads(threads)
        self.net = OnnxOracle(onnx_bytes, conv_backend=conv_backend)

    def run(self, n: int, seed: int) -> float:
        """seconds for one pass over n synthetic positions"""
        spec = self.spec
        bits, scalars, mv_idx, mv_off = netgen.synthetic_positions(spec, n, seed=seed)
        t0 = time.perf_counter()
        planes = self.oracle.expand_planes(bits, scalars, (spec.bool_channels, spec.board_size, spec.board_size),
                                           spec.scalar_channels)
        s, p = self.net.run(planes)
        self.oracle.decode_output(s, p, mv_idx, mv_off)
        return time.perf_counter() - t0

    def sample_size(self, batch: int, seconds_target: float, threads: int) -> int:
        probe_n = max(2, min(batch, threads))
        self.run(probe_n, 99)  # first pass pays thread start-up and page faults
        dt = self.run(probe_n, 100)
        return int(max(probe_n, min(batch, seconds_target / max(dt / probe_n, 1e-9))))


def cpu_restatement_rate(cfg, onnx_bytes, spec, seconds_target: float, threads: int, conv_backend: str = "c"):
    """positions/s of the oracle on `threads` host threads over a bounded sample of the workload."""
    cpu = CpuRestatement(onnx_bytes, spec, threads, conv_backend)
    n = cpu.sample_size(cfg["batch"], seconds_target, threads)
    dt = cpu.run(n, 101)
    return n / dt, n, dt


def run_reference(args, cfg, spec, onnx_bytes):
    """--impl reference: the reference's CPU implementation of the path (restated; see module docstring).
    Every step is a bounded sample of the workload sized so that the whole run ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    total = args.steps + args.warmup
    per_step_seconds = max(0.25, min(20.0, 150.0 / total))
    cpu = CpuRestatement(onnx_bytes, spec, threads)
    sample = cpu.sample_size(cfg["batch"], per_step_seconds, threads)
    rates, t_steps = [], []
    for i in range(total):
        dt = cpu.run(sample, 200 + i)
        if i >= args.warmup:
            rates.append(sample / dt)
            t_steps.append(dt)
    value = float(np.mean(rates))
    desc = (f"{sample} positions per step of the same workload (bounded sample; rate is per position so it "
            f"extrapolates linearly to batch {cfg['batch']})")
    line = {
        "impl": "reference", "metric": "NN positions/sec", "value": value, "unit": "positions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean(t_steps) * 1e3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "batch_per_gpu": cfg["batch"], "sample_positions_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "positions/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "positions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of kn-graph's CPU executor path (oracle/): the reference's Rust crates cannot be "
                "built in this image (no cargo/rustc; kn-graph 0.7.3 is not vendored)",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="chess", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    cfg = CONFIGS[args.config]
    spec = netgen.game_spec(cfg["game"])
    onnx_bytes = netgen.build_onnx(spec, cfg["depth"], cfg["channels"], seed=0)

    if args.impl == "reference":
        run_reference(args, cfg, spec, onnx_bytes)
        return

    import torch

    from kzero_b200 import replicas
    from kzero_b200.network import B200Network, PRECISION_BF16, mapper_for

    ctx = replicas.context_from_env()
    rank, world, local_rank = ctx.rank, ctx.world, ctx.local_rank
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = replicas.init_process_group(ctx, "nccl", torch.device("cuda", local_rank))

    def barrier():
        replicas.barrier(ctx, torch.cuda.synchronize)

    batch = cfg["batch"]
    net = B200Network(mapper_for(spec), onnx_bytes, batch, device=local_rank, precision=PRECISION_BF16)
    info = net.info()
    inputs = [netgen.synthetic_positions(spec, batch, seed=replicas.game_seed(ctx, i)) for i in range(N_INPUT_SETS)]

    # ---- device-resident throughput ("value"): K steps, each timed with CUDA events on the net's stream,
    #      L2 flushed (256 MiB memset) before every step outside the timed region
    for i in range(args.warmup):
        net.evaluate_packed(*inputs[i % N_INPUT_SETS])
    net.stage_packed(*inputs[0])
    net.time_staged(args.warmup, True)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    step_ms = net.time_staged(args.steps, True)
    torch.cuda.synchronize()
    wall_value = time.perf_counter() - wall0
    barrier()

    # ---- per-launch durations for the roofline of the dominant kernel (same staged batch, same stream)
    tower_ms, all_ms = [], []
    for _ in range(10):
        names, ms = net.profile_staged(True)
        all_ms.append(ms)
        tower_ms.append(sum(m for n, m in zip(names, ms) if n == "tower8" or n.startswith("conv_first") or n.startswith("block")))
    tower_launches = sum(1 for n in names if n == "tower8" or n.startswith("conv_first") or n.startswith("block"))
    share = {n: float(np.mean([m[i] for m in all_ms])) for i, n in enumerate(names)}

    # ---- end to end through the public call, host buffers in, host buffers out
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        net.evaluate_packed(*inputs[i % N_INPUT_SETS])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    # ---- informational: the same call from TWO executor threads with one network instance each, the reference's
    #      `gpu_threads_per_device = 2` topology (rust/kz-selfplay/src/server/server_alphazero.rs:89-121): one thread's
    #      host work and PCIe copies overlap the other's kernels.  Not the headline: the reference's own settings use
    #      one executor thread per device (python/main/loop_main_alpha.py:25).
    net2 = B200Network(mapper_for(spec), onnx_bytes, batch, device=local_rank, precision=PRECISION_BF16)
    for i in range(3):
        net2.evaluate_packed(*inputs[i % N_INPUT_SETS])

    def executor_thread(n, offset, count):
        for i in range(count):
            n.evaluate_packed(*inputs[(i + offset) % N_INPUT_SETS])

    half = (args.steps + 1) // 2
    threads = [threading.Thread(target=executor_thread, args=(n, k, half)) for k, n in enumerate((net, net2))]
    barrier()
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    e2e2_s = time.perf_counter() - t0
    barrier()
    net2.close()
    clocks = sampler.stop()

    dev_s = float(step_ms.sum()) * 1e-3
    dev_s, e2e_s, wall_value, e2e2_s = replicas.max_over_ranks(ctx, [dev_s, e2e_s, wall_value, e2e2_s], device="cuda")

    if rank == 0:
        peaks = measured_peaks()
        a = spec.area
        tower_flops = batch * (2.0 * a * 9 * spec.input_channels * cfg["channels"]
                               + cfg["depth"] * 2 * (2.0 * a * 9 * cfg["channels"] ** 2))
        tower_s = float(np.mean(tower_ms)) * 1e-3
        achieved = tower_flops / tower_s / 1e12
        bits, scalars, mv_idx, mv_off = inputs[0]
        h2d = int((batch + 1) * 4 + scalars.nbytes + bits.nbytes + mv_idx.nbytes)
        d2h = int(16 + batch * 5 * 4 + mv_idx.nbytes)
        traffic = None
        tpath = ROOT / "profiles" / "tower_dram_traffic.json"
        if tpath.exists() and args.config == "chess":
            traffic = json.loads(tpath.read_text()).get("dram_bytes_per_launch")
        line = {
            "metric": "NN positions/sec", "value": replicas.job_throughput(ctx, batch, args.steps, dev_s), "unit": "positions/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_s / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": cfg["workload"], "batch_per_gpu": batch, **replicas.parallelism_note(ctx),
                       "l2": "flushed before every timed step (256 MiB memset, untimed)", "conv_mode": int(info.conv_mode),
                       "flops_per_position": float(info.flops_per_position)},
            "e2e": {"value": replicas.job_throughput(ctx, batch, args.steps, e2e_s), "unit": "positions/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3,
                    "inputs": f"{N_INPUT_SETS} distinct synthetic batches rotated, host numpy buffers"},
            "e2e_two_executor_threads": {"value": replicas.job_throughput(ctx, batch, 2 * half, e2e2_s), "unit": "positions/s",
                                         "ms_per_step": e2e2_s / (2 * half) * 1e3,
                                         "note": "informational: 2 executor threads x 1 network instance each per GPU "
                                                 "(the reference's gpu_threads_per_device = 2 topology)"},
            "gpu_launches": int(net.launches_per_eval() * args.steps),
            "launches_per_step": int(net.launches_per_eval()),
            "tflops_whole_step": float(info.flops_per_position) * batch * args.steps / dev_s / 1e12,
            # the tower kernel is timed alone (one ~0.45 ms launch between L2 flushes, the GPU idles in between), so the
            # denominator is the BURST cuBLAS figure; the sustained-loop figure is reported beside it
            "roofline": {"bound": "tensor", "kernel": ("tower8_kernel" if os.environ.get("KZB_TOWER_V1") == "1" else "tower8k_kernel") if "tower8" in share else "conv_tc_kernel",
                         "achieved": achieved, "peak": peaks["tflops_burst"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tflops_burst"], "frac_of_sustained_peak": achieved / peaks["tflops_sustained"],
                         "peak_source": peaks["source"] + " (bf16_tflops, burst: kernel timed in isolation)", "traffic": traffic,
                         "launches_per_step": tower_launches, "avg_launch_ms": tower_s * 1e3 / max(tower_launches, 1),
                         "algorithmic_flops_per_launch": tower_flops / max(tower_launches, 1)},
            "step_breakdown_ms": share,
            "clocks": clocks,
            "wall_ms_per_step_incl_flush": wall_value / args.steps * 1e3,
        }
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, n, dt = cpu_restatement_rate(cfg, onnx_bytes, spec, 15.0, threads)
            line["cpu_baseline"] = {"value": rate, "unit": "positions/s", "cores": threads, "kind": "port",
                                    "sample": f"{n} positions of the same workload in {dt:.1f} s (oracle/: f32 ONNX graph "
                                              "interpreter with C/OpenMP conv loops + plane expansion + decode_output)"}
            # informational: the same interpreter with Conv handed to PyTorch's CPU kernels (oneDNN), a much stronger CPU arm
            rate_t, n_t, dt_t = cpu_restatement_rate(cfg, onnx_bytes, spec, 8.0, threads, "torch")
            line["cpu_baseline"]["torch_onednn"] = {"value": rate_t, "unit": "positions/s", "cores": threads,
                                                    "sample": f"{n_t} positions in {dt_t:.1f} s"}
        print(json.dumps(line), flush=True)
    net.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
