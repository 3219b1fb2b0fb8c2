"""
Multi-GPU plumbing: one process per GPU, envs partitioned across ranks, no collective on
the data path (envs never interact -- every `FireSimulation` is a separate object in the
reference).  `torch.distributed` is used only for the rendezvous, the barriers around timed
regions and the max / sum reductions of scalars (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Tuple


def env_shard(total_envs: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of envs owned by `rank`: (first env, count); sizes differ by at most 1."""
    if not (0 <= rank < world) or total_envs < 0:
        raise ValueError(f"env_shard: rank {rank} of {world}, {total_envs} envs")
    base, extra = divmod(total_envs, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def row_slabs(total_rows: int, world: int) -> List[Tuple[int, int]]:
    """Row slabs (first row, rows) of a single grid split across `world` GPUs (slab mode)."""
    return [env_shard(total_rows, world, r) for r in range(world)]


@dataclass
class RankContext:
    rank: int = 0
    world: int = 1
    local_rank: int = 0
    backend: Optional[str] = None

    @classmethod
    def from_env(cls, backend: Optional[str] = None, device_id=None) -> "RankContext":
        """Reads RANK / WORLD_SIZE / LOCAL_RANK (torchrun) and joins the process group if world > 1."""
        ctx = cls(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                  int(os.environ.get("LOCAL_RANK", "0")), backend)  # fmt: skip
        if ctx.world > 1:
            import torch.distributed as dist

            if not dist.is_initialized():
                kw = {"device_id": device_id} if device_id is not None else {}
                dist.init_process_group(backend or "nccl", **kw)
        return ctx

    def _tensor(self, value: float):
        import torch

        dev = "cuda" if (self.backend or "nccl") == "nccl" else "cpu"
        return torch.tensor([float(value)], dtype=torch.float64, device=dev)

    def barrier(self, host: bool = False) -> None:
        """All ranks meet.  host=True: on a gloo group, i.e. without putting a collective kernel on the GPUs
        (an NCCL barrier right before a timed region costs every rank some tens of microseconds before its next
        kernel starts, which a sub-millisecond region would report as a scaling loss)."""
        if self.world > 1:
            import torch.distributed as dist

            if host and (self.backend or "nccl") == "nccl":
                if getattr(self, "_host_group", None) is None:
                    self._host_group = dist.new_group(backend="gloo")
                dist.barrier(group=self._host_group)
            else:
                dist.barrier()

    def max(self, value: float) -> float:
        """Max over ranks (timings are reported as the slowest rank's)."""
        if self.world == 1:
            return float(value)
        import torch.distributed as dist

        t = self._tensor(value)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, value: float) -> float:
        if self.world == 1:
            return float(value)
        import torch.distributed as dist

        t = self._tensor(value)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def close(self) -> None:
        if self.world > 1:
            import torch.distributed as dist

            if dist.is_initialized():
                dist.destroy_process_group()


def aggregate_throughput(ctx: RankContext, cells_this_rank: int, steps: int, seconds_this_rank: float) -> float:
    """Whole-job cell-updates/s: all ranks' cells over the slowest rank's time."""
    total_cells = ctx.sum(float(cells_this_rank))
    return total_cells * steps / ctx.max(seconds_this_rank)
