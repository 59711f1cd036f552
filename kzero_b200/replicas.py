"""Multi-GPU plumbing of the hot path: independent replicas sharded by game, no data-path collective.

The reference shards self-play by game: every device gets its own generator pool, job channel and executor
threads (`for device in devices { spawn_device_threads(...) }`, rust/kz-selfplay/src/server/server.rs:316-331;
generator ids are `concurrent_games * device_id + local_id`, server_alphazero.rs:66).  A position's evaluation
depends only on that position and the read-only weights, so there is nothing to exchange between GPUs.

Here that is one process per GPU (torchrun): each rank owns a full weight replica and evaluates its own games.
`torch.distributed` is used for exactly two things, both off the data path: the barrier that brackets the timed
region and the MAX-over-ranks reduction of the measured times.  The backend is NCCL on GPUs and gloo on CPU
(tests/test_replicas.py runs this file with world_size 2 on gloo).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Sequence


@dataclass(frozen=True)
class ReplicaContext:
    rank: int
    world: int
    local_rank: int

    @property
    def is_root(self) -> bool:
        return self.rank == 0


def context_from_env() -> ReplicaContext:
    """RANK / WORLD_SIZE / LOCAL_RANK as torchrun exports them; a plain `python bench.py` is one replica."""
    return ReplicaContext(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                          int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(ctx: ReplicaContext, backend: str, device=None):
    """Join the job's process group (no-op for a single replica).  Rendezvous on 127.0.0.1 unless told otherwise."""
    if ctx.world == 1:
        return None
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    kwargs = {}
    if device is not None and backend == "nccl":
        kwargs["device_id"] = device
    dist.init_process_group(backend, rank=ctx.rank, world_size=ctx.world, **kwargs)
    return dist


def game_seed(ctx: ReplicaContext, game_index: int, games_per_replica: int = 1000) -> int:
    """Global id of the `game_index`-th batch of games owned by this replica: disjoint ranges per rank, the
    same rule as the reference's `concurrent_games * device_id + local_id`."""
    if not 0 <= game_index < games_per_replica:
        raise ValueError("game_index out of range for this replica")
    return games_per_replica * ctx.rank + game_index


def barrier(ctx: ReplicaContext, sync_device=None) -> None:
    if ctx.world > 1:
        import torch.distributed as dist

        dist.barrier()
    if sync_device is not None:
        sync_device()


def max_over_ranks(ctx: ReplicaContext, values: Sequence[float], device="cpu") -> list:
    """Element-wise MAX of per-rank measurements (the slowest replica defines the job's time)."""
    import torch

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if ctx.world > 1:
        import torch.distributed as dist

        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def job_throughput(ctx: ReplicaContext, positions_per_replica_step: int, steps: int, seconds_max_over_ranks: float) -> float:
    """Whole-job positions/s: every replica processed `positions_per_replica_step * steps` positions of its own
    games (weak scaling: per-GPU work is fixed) in at most `seconds_max_over_ranks`."""
    return ctx.world * positions_per_replica_step * steps / seconds_max_over_ranks


def parallelism_note(ctx: ReplicaContext) -> Dict[str, str]:
    return {"parallelism": f"replicas x{ctx.world} (sharded by game, no collective on the data path)"}
