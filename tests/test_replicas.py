"""The N>1 path on CPU: world_size 2 over gloo.  Replicas are sharded by game with no data-path collective
(SURVEY.md 8(e)); torch.distributed only brackets the timed region (barrier) and reduces the timings (MAX)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from kzero_b200 import netgen, replicas

    ctx = replicas.context_from_env()
    assert (ctx.rank, ctx.world, ctx.is_root) == (rank, world, rank == 0)
    dist = replicas.init_process_group(ctx, "gloo")
    try:
        # each replica draws its own games: same spec, disjoint seeds, nothing exchanged
        spec = netgen.game_spec("ataxx-7")
        seeds = [replicas.game_seed(ctx, i) for i in range(3)]
        bits, scalars, mv_idx, mv_off = netgen.synthetic_positions(spec, 16, seed=seeds[0])
        replicas.barrier(ctx)
        # pretend measurements: rank 1 is the slow replica
        dev_s, e2e_s = (0.010, 0.020) if rank == 0 else (0.015, 0.018)
        dev_max, e2e_max = replicas.max_over_ranks(ctx, [dev_s, e2e_s])
        replicas.barrier(ctx)
        value = replicas.job_throughput(ctx, 16, 5, dev_max)
        np.savez(Path(out_dir) / f"rank{rank}.npz", seeds=np.array(seeds), bits=bits, dev_max=dev_max, e2e_max=e2e_max,
                 value=value)
    finally:
        dist.destroy_process_group()


def test_two_replicas_over_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in range(world))
    # sharded by game: disjoint seed ranges and different positions on each replica
    assert not set(r0["seeds"].tolist()) & set(r1["seeds"].tolist())
    assert not np.array_equal(r0["bits"], r1["bits"])
    # every rank sees the MAX over ranks, element-wise
    for r in (r0, r1):
        assert float(r["dev_max"]) == pytest.approx(0.015)
        assert float(r["e2e_max"]) == pytest.approx(0.020)
        # whole-job aggregate: both replicas' positions over the slowest replica's time (weak scaling)
        assert float(r["value"]) == pytest.approx(2 * 16 * 5 / 0.015)


def test_single_replica_needs_no_process_group(monkeypatch):
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    from kzero_b200 import replicas

    ctx = replicas.context_from_env()
    assert ctx.world == 1 and ctx.is_root
    assert replicas.init_process_group(ctx, "gloo") is None
    replicas.barrier(ctx)
    assert replicas.max_over_ranks(ctx, [1.5, 2.5]) == [1.5, 2.5]
    assert replicas.job_throughput(ctx, 1024, 10, 0.005) == pytest.approx(1024 * 10 / 0.005)
    with pytest.raises(ValueError):
        replicas.game_seed(ctx, 1000)


def test_reference_arm_runs_on_rank0_only(tmp_path):
    """bench.py --impl reference under torchrun: rank 0 alone prints the line, other ranks exit 0 without work."""
    import json
    import subprocess

    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "3", "--config", "ataxx"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env["RANK"] = "0"
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "3", "--config", "ataxx"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "positions/s"
