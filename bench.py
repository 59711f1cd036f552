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
    # BASELINE.json configs[4]
    "go19": dict(game="go-19", depth=40, channels=256, batch=8192,
                 workload="go 19x19 ResNet 40x256, batch 8192"),
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


def gpu_comparator(cfg, spec, weights, device, iters, variants=("bf16_channels_last", "tf32_nchw", "fp32_nchw")):
    """LIBRARY baseline on the same box (SURVEY.md 2c / 8(d) "the kernel to beat"): the conv tower of the same net -- same
    weights, same batch -- on torch + cuDNN, the stand-in for the reference's cuDNN executor schedule
    (cudnnConvolutionBiasActivationForward per layer, rust/kz-core/src/network/cudnn.rs:73, docs/conv_bn_sm_flow.svg): one fused
    conv+bias+ReLU call per layer (torch.cudnn_convolution_relu; plain conv2d + relu_ where cuDNN refuses the fusion) and a
    separate residual add (the reference cannot fuse it either: the ReLU comes before the add), captured in a CUDA graph,
    cudnn.benchmark autotuned.  Timed exactly like the product's tower: CUDA events around one replay, L2 flushed before each.
    None of this repo's kernels run here and nothing here runs inside the product's timed regions."""
    import torch
    import torch.nn.functional as F

    depth, ch, batch, size = cfg["depth"], cfg["channels"], cfg["batch"], spec.board_size
    flops = batch * (2.0 * spec.area * 9 * spec.input_channels * ch + depth * 2 * (2.0 * spec.area * 9 * ch * ch))
    out = {"kind": "library baseline, not this repo's kernels", "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
           "what": "conv tower only (2*depth+1 conv3x3 + bias (+ReLU) per layer, residual adds), CUDA graph, cudnn.benchmark",
           "algorithmic_flops": flops, "variants": {}}
    torch.backends.cudnn.benchmark = True
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    for tag in variants:
        dtype = torch.bfloat16 if tag.startswith("bf16") else torch.float32
        fmt = torch.channels_last if tag.endswith("channels_last") else torch.contiguous_format
        tf32 = tag.startswith("tf32")
        if tag == "fp32_nchw" and flops > 20e12:
            continue  # strict fp32 on CUDA cores: seconds per tower at the go sizes
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            ws = [(torch.from_numpy(weights[f"w{k}"]).to(device=device, dtype=dtype).contiguous(memory_format=fmt),
                   torch.from_numpy(weights[f"b{k}"]).to(device=device, dtype=dtype)) for k in range(1, 2 * depth + 2)]
            x0 = torch.randn(batch, spec.input_channels, size, size, device=device, dtype=dtype).contiguous(memory_format=fmt)
            fused = {"ok": True}

            def conv_relu(x, w, b):
                if fused["ok"]:
                    try:
                        return torch.cudnn_convolution_relu(x, w, b, (1, 1), (1, 1), (1, 1), 1)
                    except Exception:  # noqa: BLE001
                        fused["ok"] = False
                return F.conv2d(x, w, b, padding=1).relu_()

            def tower():
                x = F.conv2d(x0, ws[0][0], ws[0][1], padding=1)
                for d in range(depth):
                    t = conv_relu(x, *ws[1 + 2 * d])
                    y = conv_relu(t, *ws[2 + 2 * d])
                    x = x.add_(y)
                return x

            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.no_grad(), torch.cuda.stream(side):
                for _ in range(3):
                    tower()
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize(device)
            graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(graph):
                y = tower()
            ms = []
            for _ in range(iters + 2):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                graph.replay()
                e1.record()
                e1.synchronize()
                ms.append(e0.elapsed_time(e1))
            ms = ms[2:]
            med = float(np.median(ms))
            out["variants"][tag] = {"ms_per_tower": med, "ms_min": float(min(ms)), "tflops_algorithmic": flops / (med * 1e-3) / 1e12,
                                    "positions_per_s_tower_only": batch / (med * 1e-3), "fused_conv_bias_relu": bool(fused["ok"]), "iters": iters}
            del graph, y, ws, x0
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out["variants"][tag] = {"error": repr(e)[:200]}
    torch.backends.cudnn.allow_tf32 = True
    return out


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


TOWER_STEP = lambda n: n == "tower8" or n.startswith("conv_first") or n.startswith("block")  # noqa: E731


def measure(name, args, ctx, steps, warmup, with_two_threads, sampler=None):
    """One configuration on this rank's GPU -> dict of raw measurements (times are max over ranks where they feed `value`)."""
    import torch

    from kzero_b200 import replicas
    from kzero_b200.network import B200Network, PRECISION_BF16, mapper_for

    cfg = CONFIGS[name]
    spec = netgen.game_spec(cfg["game"])
    weights = {}
    onnx_bytes = netgen.build_onnx(spec, cfg["depth"], cfg["channels"], seed=0, weights_out=weights)
    local_rank, batch = ctx.local_rank, cfg["batch"]

    def barrier():
        replicas.barrier(ctx, torch.cuda.synchronize)

    net = B200Network(mapper_for(spec), onnx_bytes, batch, device=local_rank, precision=PRECISION_BF16)
    info = net.info()
    n_sets = N_INPUT_SETS if batch * spec.policy_size < (1 << 24) else 2
    inputs = [netgen.synthetic_positions(spec, batch, seed=replicas.game_seed(ctx, i)) for i in range(n_sets)]

    # ---- device-resident throughput ("value"): K steps, each timed with CUDA events on the net's stream,
    #      L2 flushed (256 MiB memset) before every step outside the timed region
    for i in range(warmup):
        net.evaluate_packed(*inputs[i % n_sets])
    net.stage_packed(*inputs[0])
    net.time_staged(warmup, True)
    barrier()
    if sampler is not None:
        sampler.start()
    wall0 = time.perf_counter()
    step_ms = net.time_staged(steps, True)
    torch.cuda.synchronize()
    wall_value = time.perf_counter() - wall0
    barrier()

    # ---- per-launch durations for the rooflines (same staged batch, same stream, CUDA events around every launch)
    all_ms = []
    for _ in range(10 if batch * spec.area <= (1 << 19) else 3):
        names, ms = net.profile_staged(True)
        all_ms.append(ms)
    share = {}
    for i, n in enumerate(names):
        share[n] = share.get(n, 0.0) + float(np.mean([m[i] for m in all_ms]))
    tower_ms = sum(v for n, v in share.items() if TOWER_STEP(n))
    tower_launches = sum(1 for n in names if TOWER_STEP(n))

    # ---- end to end through the public call, host buffers in, host buffers out
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        net.evaluate_packed(*inputs[i % n_sets])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    # ---- sustained: the same staged step back to back, no L2 flush, for >= 2 s -- the regime the self-play loop runs in
    #      (power cap, lower SM clock); compared with the SUSTAINED cuBLAS figure
    sustained = None
    if args.sustained_seconds > 0:
        per = float(np.median(step_ms)) * 1e-3
        chunk = max(10, int(0.25 / per))
        sus_sampler = ClockSampler(local_rank)
        net.time_staged(chunk, False)
        sus_sampler.start()
        t0 = time.perf_counter()
        sus_ms = []
        while time.perf_counter() - t0 < args.sustained_seconds:
            sus_ms.append(net.time_staged(chunk, False))
        sus_clocks = sus_sampler.stop()
        sus_ms = np.concatenate(sus_ms)
        tail = sus_ms[len(sus_ms) // 2:]  # second half: clocks have settled
        sustained = {"ms_per_step": float(np.mean(tail)), "steps": int(len(sus_ms)), "seconds": float(sus_ms.sum() * 1e-3), "clocks": sus_clocks}

    # ---- informational: the same call from TWO executor threads with one network instance each, the reference's
    #      `gpu_threads_per_device = 2` topology (rust/kz-selfplay/src/server/server_alphazero.rs:89-121): one thread's
    #      host work and PCIe copies overlap the other's kernels.  Not the headline: the reference's own settings use
    #      one executor thread per device (python/main/loop_main_alpha.py:25).
    e2e2_s, half = None, (steps + 1) // 2
    if with_two_threads:
        net2 = B200Network(mapper_for(spec), onnx_bytes, batch, device=local_rank, precision=PRECISION_BF16)
        for i in range(3):
            net2.evaluate_packed(*inputs[i % n_sets])

        def executor_thread(n, offset, count):
            for i in range(count):
                n.evaluate_packed(*inputs[(i + offset) % n_sets])

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
    clocks = sampler.stop() if sampler is not None else None
    launches = net.launches_per_eval()
    net.close()

    comparator = None
    if not args.no_comparator and ctx.is_root:
        big = batch * spec.area * cfg["channels"] > (1 << 27)
        comparator = gpu_comparator(cfg, spec, weights, torch.device("cuda", local_rank), iters=3 if big else 10,
                                    variants=("bf16_channels_last", "tf32_nchw") if big else ("bf16_channels_last", "tf32_nchw", "fp32_nchw"))
    barrier()

    dev_s = float(step_ms.sum()) * 1e-3
    times = [dev_s, e2e_s, wall_value, e2e2_s if e2e2_s is not None else 0.0, sustained["ms_per_step"] if sustained else 0.0]
    dev_s, e2e_s, wall_value, e2e2_s_max, sus_ms_max = replicas.max_over_ranks(ctx, times, device="cuda")
    bits, scalars, mv_idx, mv_off = inputs[0]
    return dict(name=name, cfg=cfg, spec=spec, onnx_bytes=onnx_bytes, info=info, batch=batch, steps=steps, warmup=warmup, dev_s=dev_s, e2e_s=e2e_s,
                wall_value=wall_value, e2e2_s=e2e2_s_max if e2e2_s is not None else None, half=half, share=share, tower_ms=tower_ms,
                tower_launches=tower_launches, launches=launches, clocks=clocks, sustained=sustained, sus_ms=sus_ms_max,
                comparator=comparator, n_sets=n_sets,
                h2d=int((batch + 1) * 4 + scalars.nbytes + bits.nbytes + mv_idx.nbytes), d2h=int(16 + batch * 5 * 4 + mv_idx.nbytes),
                moves=int(mv_idx.size))


def rooflines(m, peaks):
    """roofline of the dominant kernel (the conv tower, tensor-bound) and of K2 / K3 (HBM-bound), from CUDA-event launch times."""
    cfg, spec, batch, share = m["cfg"], m["spec"], m["batch"], m["share"]
    a, c = spec.area, cfg["channels"]
    tower_flops = batch * (2.0 * a * 9 * spec.input_channels * c + cfg["depth"] * 2 * (2.0 * a * 9 * c * c))
    tower_s = m["tower_ms"] * 1e-3
    achieved = tower_flops / tower_s / 1e12
    if "tower8" in share:
        kernel = "tower8k_kernel"
    else:
        kernel = "conv_tc_kernel" if os.environ.get("KZB_NO_I2C") == "1" else "conv_i2c_kernel"
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (scripts/ncu_digest.py); only
    # for the workload the capture was taken on
    tpath = ROOT / "profiles" / "dram_traffic.json"
    table = json.loads(tpath.read_text()) if tpath.exists() else {}

    def traffic_of(kern):
        e = table.get(f"{kern}/{m['name']}")
        return e.get("dram_bytes_per_launch") if e else None
    traffic = traffic_of(kernel)
    # the tower is timed launch by launch between L2 flushes with the GPU idle in between: burst denominator
    tower = {"bound": "tensor", "kernel": kernel, "achieved": achieved, "peak": peaks["tflops_burst"], "unit": "TFLOP/s",
             "frac": achieved / peaks["tflops_burst"],
             "peak_source": peaks["source"] + " (bf16_tflops, burst: launches timed one by one between L2 flushes)", "traffic": traffic,
             "launches_per_step": m["tower_launches"], "avg_launch_ms": m["tower_ms"] / max(m["tower_launches"], 1),
             "algorithmic_flops_per_launch": tower_flops / max(m["tower_launches"], 1)}
    out = {"roofline": tower}
    if m["sustained"]:
        step_flops = float(m["info"].flops_per_position) * batch
        sus = step_flops / (m["sus_ms"] * 1e-3) / 1e12
        out["roofline_sustained"] = {"bound": "tensor", "what": "whole step (encode + tower + heads) back to back for >= "
                                     f"{m['sustained']['seconds']:.1f} s, no L2 flush", "achieved": sus, "peak": peaks["tflops_sustained"],
                                     "unit": "TFLOP/s", "frac": sus / peaks["tflops_sustained"], "ms_per_step": m["sus_ms"],
                                     "peak_source": peaks["source"] + " (bf16_tflops_sustained)", "clocks": m["sustained"]["clocks"]}
    # K2: packed record in, bf16 planes out (136 B + C_in*A*2 B per chess position)
    enc_bytes = batch * (spec.bits_bytes + spec.scalar_channels * 4 + spec.input_channels * a * 2)
    if share.get("encode"):
        gbs = enc_bytes / (share["encode"] * 1e-3) / 1e9
        out["roofline_k2"] = {"bound": "hbm", "kernel": "encode_kc_kernel" if "tower8" in share else "encode_nhwc_kernel", "achieved": gbs,
                              "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "launch_ms": share["encode"],
                              "algorithmic_bytes_per_launch": enc_bytes,
                              "traffic": traffic_of("encode_kc_kernel" if "tower8" in share else "encode_nhwc_kernel"),
                              "note": "a 10 us launch: latency-bound, not bandwidth-bound"}
    # K3: tower output in (C*A*2 B per position), values + legal-move probabilities out
    head_ms = sum(v for n, v in share.items() if not TOWER_STEP(n) and n != "encode")
    head_bytes = batch * (c * a * 2 + 5 * 4) + m["moves"] * 8
    if head_ms > 0:
        gbs = head_bytes / (head_ms * 1e-3) / 1e9
        out["roofline_k3"] = {"bound": "hbm", "kernel": "heads8_kernel" if "heads8" in share else "head convs + heads_tail_kernel",
                              "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "launch_ms": head_ms,
                              "algorithmic_bytes_per_launch": head_bytes, "traffic": traffic_of("heads8_kernel") if "heads8" in share else None}
    return out


def selfplay_record(ctx, game_name, seconds, dist):
    """Config C4 (BASELINE.json configs[3]): the full self-play loop -- generator threads running 800-visit searches with
    virtual loss, executor threads batching their requests into the evaluator -- on this rank's GPU with this rank's share
    of the host cores.  nodes/s = (real + cached evals) / s, the collector's line (collector.rs:172-191)."""
    from kzero_b200 import replicas, selfplay

    game = {"chess-synthetic": selfplay.GAME_SYNTH_CHESS, "chess": selfplay.GAME_CHESS}[game_name]
    spec = netgen.game_spec("chess")
    onnx_bytes = netgen.build_onnx(spec, 16, 128, seed=0)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    share = max(1, cores // ctx.world)
    blocking = share < 12
    gpu_threads = 3
    cpu_threads = share if blocking else share - gpu_threads
    # games in flight per GPU: the reference's formula for three executors, (3 + 1) * 1024 / 16 = 256 (server_alphazero.rs:47), when the
    # replica has plenty of host cores (the loop is GPU-bound; more games only cost cache: 2.04 M -> 1.97 M nodes/s at 384 on a 16-core
    # box); twice as many when it has few (executors sleep on events): the extra games hide the executor <-> generator wake-up latency
    # of eight processes sharing one host -- synthetic game 256 -> 320 -> 384 games: 6.88x -> 7.07x -> 7.15x of one GPU at N = 8 with 4
    # cores per GPU (profiles/r02_n8_selfplay_ab.txt), real chess 384 -> 512: 6.40x -> 6.57x (profiles/r02_n8_chess_b.txt)
    concurrent_games = 512 if blocking else 256
    cfg = selfplay.default_config(game=game, visits=800, search_batch=16, gpu_batch=1024, cpu_threads=cpu_threads, gpu_threads=gpu_threads,
                                  concurrent_games=concurrent_games, duration_s=seconds, seed=replicas.game_seed(ctx, 1),
                                  executor_blocking_sync=int(blocking))
    replicas.barrier(ctx)
    r = selfplay.run(onnx_bytes, cfg, device=ctx.local_rank)
    counts = [r.real_evals, r.cached_evals, r.batches, r.moves_played, r.games_finished]
    if dist is not None:
        import torch

        t = torch.tensor(counts, dtype=torch.float64, device="cuda")
        dist.all_reduce(t)  # sums over replicas: the games are disjoint
        counts = [float(v) for v in t.tolist()]
    (secs,) = replicas.max_over_ranks(ctx, [r.seconds], device="cuda" if dist is not None else "cpu")
    real, cached, batches, moves, games = counts
    return {"metric": "self-play MCTS nodes/sec", "value": (real + cached) / secs, "unit": "nodes/s", "n_gpus": ctx.world,
            "nn_positions_per_s": real / secs, "cache_hit_rate": cached / max(real + cached, 1), "mean_batch": real / max(batches, 1),
            "moves_per_s": moves / secs, "games_finished": games, "seconds": secs, "scaling": "weak",
            "game": {"chess-synthetic": "chess-shaped synthetic game (13x8x8 + 8 planes, 1880-move policy, 20-45 legal moves)",
                     "chess": "chess (legal move generation, ChessStdMapper encoding)"}[game_name],
            "settings": "800 visits, search batch 16 with virtual loss, LRU cache 800, net chess 16x128, gpu batch 1024",
            "concurrent_games_per_gpu": int(r.concurrent_games),
            "host_cores": cores, "cpu_threads_per_gpu": cpu_threads, "gpu_threads_per_gpu": gpu_threads,
            "executor_blocking_sync": bool(blocking)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="chess", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-comparator", action="store_true", help="skip the torch + cuDNN library baseline of the tower")
    ap.add_argument("--extras", default="auto", choices=["auto", "none", "all"],
                    help="sub-records for the other BASELINE.json configs (go9, go19, self-play); auto = all for the default config")
    ap.add_argument("--selfplay-seconds", type=float, default=6.0)
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        spec = netgen.game_spec(cfg["game"])
        run_reference(args, cfg, spec, netgen.build_onnx(spec, cfg["depth"], cfg["channels"], seed=0))
        return

    import torch

    from kzero_b200 import replicas

    ctx = replicas.context_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(ctx.local_rank)
    dist = replicas.init_process_group(ctx, "nccl", torch.device("cuda", ctx.local_rank))
    peaks = measured_peaks()

    m = measure(args.config, args, ctx, args.steps, args.warmup, with_two_threads=True, sampler=ClockSampler(ctx.local_rank))
    batch, spec = m["batch"], m["spec"]
    line = {
        "metric": "NN positions/sec", "value": replicas.job_throughput(ctx, batch, m["steps"], m["dev_s"]), "unit": "positions/s",
        "n_gpus": ctx.world, "steps": m["steps"], "warmup": m["warmup"], "ms_per_step": m["dev_s"] / m["steps"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": cfg["workload"], "batch_per_gpu": batch, **replicas.parallelism_note(ctx),
                   "l2": "flushed before every timed step (256 MiB memset, untimed)", "conv_mode": int(m["info"].conv_mode),
                   "flops_per_position": float(m["info"].flops_per_position)},
        "e2e": {"value": replicas.job_throughput(ctx, batch, m["steps"], m["e2e_s"]), "unit": "positions/s", "h2d_bytes_per_step": m["h2d"],
                "d2h_bytes_per_step": m["d2h"], "ms_per_step": m["e2e_s"] / m["steps"] * 1e3,
                "inputs": f"{m['n_sets']} distinct synthetic batches rotated, host numpy buffers"},
        "e2e_two_executor_threads": {"value": replicas.job_throughput(ctx, batch, 2 * m["half"], m["e2e2_s"]), "unit": "positions/s",
                                     "ms_per_step": m["e2e2_s"] / (2 * m["half"]) * 1e3,
                                     "note": "informational: 2 executor threads x 1 network instance each per GPU "
                                             "(the reference's gpu_threads_per_device = 2 topology)"},
        "gpu_launches": int(m["launches"] * m["steps"]),
        "launches_per_step": int(m["launches"]),
        "tflops_whole_step": float(m["info"].flops_per_position) * batch * m["steps"] / m["dev_s"] / 1e12,
        **rooflines(m, peaks),
        "step_breakdown_ms": m["share"],
        "clocks": m["clocks"],
        "wall_ms_per_step_incl_flush": m["wall_value"] / m["steps"] * 1e3,
    }
    if m["comparator"] is not None:
        best = min((v["ms_per_tower"] for v in m["comparator"]["variants"].values() if "ms_per_tower" in v), default=None)
        m["comparator"]["ours_tower_ms"] = m["tower_ms"]
        m["comparator"]["ours_vs_best_library"] = (best / m["tower_ms"]) if best else None
        line["gpu_comparator"] = m["comparator"]

    extras = args.extras == "all" or (args.extras == "auto" and args.config == "chess")
    if extras:
        # the other BASELINE.json configurations, measured the same way with fewer steps (a go-19 step is ~0.25 s)
        other = {}
        for name, steps in (("go9", 20), ("go19", 5)):
            sub_args = argparse.Namespace(**{**vars(args), "sustained_seconds": min(args.sustained_seconds, 2.0)})
            o = measure(name, sub_args, ctx, steps, 3, with_two_threads=False, sampler=ClockSampler(ctx.local_rank))
            rec = {"workload": CONFIGS[name]["workload"], "value": replicas.job_throughput(ctx, o["batch"], o["steps"], o["dev_s"]),
                   "unit": "positions/s", "n_gpus": ctx.world, "steps": o["steps"], "warmup": o["warmup"], "ms_per_step": o["dev_s"] / o["steps"] * 1e3,
                   "e2e": {"value": replicas.job_throughput(ctx, o["batch"], o["steps"], o["e2e_s"]), "unit": "positions/s",
                           "h2d_bytes_per_step": o["h2d"], "d2h_bytes_per_step": o["d2h"]},
                   "launches_per_step": int(o["launches"]), **rooflines(o, peaks), "clocks": o["clocks"]}
            if o["comparator"] is not None:
                best = min((v["ms_per_tower"] for v in o["comparator"]["variants"].values() if "ms_per_tower" in v), default=None)
                o["comparator"]["ours_tower_ms"] = o["tower_ms"]
                o["comparator"]["ours_vs_best_library"] = (best / o["tower_ms"]) if best else None
                rec["gpu_comparator"] = o["comparator"]
            other[name] = rec
        line["other_configs"] = other
        if args.selfplay_seconds > 0:
            line["selfplay"] = {g: selfplay_record(ctx, g, args.selfplay_seconds, dist) for g in ("chess-synthetic", "chess")}

    if ctx.is_root:
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, n, dt = cpu_restatement_rate(cfg, m["onnx_bytes"], spec, 15.0, threads)
            line["cpu_baseline"] = {"value": rate, "unit": "positions/s", "cores": threads, "kind": "port",
                                    "sample": f"{n} positions of the same workload in {dt:.1f} s (oracle/: f32 ONNX graph "
                                              "interpreter with C/OpenMP conv loops + plane expansion + decode_output)"}
            # informational: the same interpreter with Conv handed to PyTorch's CPU kernels (oneDNN), a much stronger CPU arm
            rate_t, n_t, dt_t = cpu_restatement_rate(cfg, m["onnx_bytes"], spec, 8.0, threads, "torch")
            line["cpu_baseline"]["torch_onednn"] = {"value": rate_t, "unit": "positions/s", "cores": threads,
                                                    "sample": f"{n_t} positions in {dt_t:.1f} s"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
