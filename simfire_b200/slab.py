"""
Single-huge-grid mode (BASELINE config 5): one H x W grid split in horizontal slabs, one slab
per engine.  The sweep kernel of a slab reads the row above / below it directly from the
neighbour slab's state plane -- peer device memory over NVLink when the neighbour is another
GPU (CUDA IPC handle opened in this process), a plain pointer when it is an engine on the same
GPU -- so there is no halo copy.  Per step the slabs only have to agree on two things:

    barrier            every slab finished step t-1 (its ignitions are visible)
    sweep              on every slab
    OR of flags        any_live / any_cand per env across slabs (all-reduce MAX, 32 B per env)
    eval               on every slab

Two layouts: all slabs in this process (`world == 1`; used by the single-GPU tests) or one slab
per rank (`torch.distributed`, NCCL collectives ordered on the engine's stream).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .engine import FireEngine
from .sharding import RankContext, row_slabs


class SlabGrid:
    def __init__(self, total_H: int, W: int, planes: Dict[str, np.ndarray], *, n_slabs: Optional[int] = None,
                 ctx: Optional[RankContext] = None, E: int = 1, device: int = 0, sync: str = "p2p",
                 **engine_kwargs) -> None:  # fmt: skip
        """
        `planes`: the eight static (total_H, W) planes of the WHOLE grid (each rank slices its
        rows).  With a multi-rank `ctx` there is one slab per rank; otherwise `n_slabs` slabs
        are created on `device`.
        """
        self.ctx = ctx if ctx is not None else RankContext()
        self.total_H, self.W, self.E = int(total_H), int(W), int(E)
        self.distributed = self.ctx.world > 1
        # per-step agreement between ranks: "p2p" = mailboxes in peer memory, polled by tiny
        # kernels (no host or NCCL round trip); "nccl" = two small all-reduces per step
        self.sync = sync
        world = self.ctx.world if self.distributed else int(n_slabs or 1)
        self.slabs = row_slabs(self.total_H, world)
        mine = [self.ctx.rank] if self.distributed else list(range(world))
        self.engines: List[FireEngine] = []
        self.my_slabs = mine
        single = world == 1  # one slab = the whole grid: an ordinary engine (TMA sweep, no coordination)
        for i in mine:
            y0, h = self.slabs[i]
            eng = FireEngine(h, W, E, device=device, slab_y0=0 if single else y0,
                             slab_total_H=0 if single else self.total_H, **engine_kwargs)  # fmt: skip
            eng.set_static({k: np.broadcast_to(np.asarray(v, dtype=np.float64), (self.total_H, W))[y0 : y0 + h]
                            for k, v in planes.items()})  # fmt: skip
            self.engines.append(eng)
        self._opened: List[int] = []
        self._connect()
        self._flags = [e.flags_tensors() for e in self.engines]
        self._steps = 0
        if self.distributed:
            import torch

            self._stream = torch.cuda.ExternalStream(self.engines[0].stream, device=torch.device("cuda", device))
            self._token = torch.zeros(1, dtype=torch.int32, device=f"cuda:{device}")

    # -- wiring --------------------------------------------------------------------------------
    def _connect(self) -> None:
        geo = [e.state_device() for e in self.engines]  # (ptr, plane, pitch, cell_bytes)
        if not self.distributed:
            for k, eng in enumerate(self.engines):
                top = bottom = (0, 0)
                if k > 0:
                    ptr, plane, pitch, cb = geo[k - 1]
                    top = (ptr + (self.slabs[k - 1][1] - 1) * pitch * cb, plane)
                if k + 1 < len(self.engines):
                    ptr, plane, pitch, cb = geo[k + 1]
                    bottom = (ptr, plane)
                eng.set_halo(top[0], top[1], bottom[0], bottom[1])
            return
        import torch.distributed as dist

        eng = self.engines[0]
        ptr, plane, pitch, cb = geo[0]
        r, world = self.ctx.rank, self.ctx.world
        info = [None] * world  # (ipc handle, plane cells, pitch cells, cell bytes, rows, mailbox offset)
        dist.all_gather_object(info, (eng.ipc_export(), plane, pitch, cb, self.slabs[r][1], eng.slab_mailbox()[1]))
        # map the state planes we need: the two neighbours (halo rows) and, for the peer-memory
        # handshakes, every slab (its mailbox sits behind its state plane)
        need = {q for q in (r - 1, r + 1) if 0 <= q < world}
        if self.sync == "p2p":
            need |= set(range(world)) - {r}
        bases = {}
        for q in sorted(need):
            bases[q] = eng.ipc_open(info[q][0])
            self._opened.append(bases[q])
        top = bottom = (0, 0)
        if r > 0:
            _, pl, pi, c, rows, _ = info[r - 1]
            top = (bases[r - 1] + (rows - 1) * pi * c, pl)
        if r + 1 < world:
            bottom = (bases[r + 1], info[r + 1][1])
        eng.set_halo(top[0], top[1], bottom[0], bottom[1])
        if self.sync == "p2p":
            boxes = [eng.slab_mailbox()[0] if q == r else bases[q] + info[q][5] for q in range(world)]
            eng.slab_connect(r, world, boxes)
        self.ctx.barrier()

    # -- mutations (coordinates are those of the whole grid) ---------------------------------------
    def reset(self, positions, envs: Optional[Sequence[int]] = None) -> None:
        for e in self.engines:
            e.reset(positions, envs=envs)
        self._sync_all()

    def apply_points(self, points) -> None:
        for e in self.engines:
            e.apply_points(points)
        self._sync_all()

    def _sync_all(self) -> None:
        for e in self.engines:
            e.synchronize()
        self.ctx.barrier()

    # -- stepping ------------------------------------------------------------------------------------
    def step(self, n: int = 1) -> None:
        if self.distributed:
            self._step_distributed(n)
        else:
            self._step_local(n)

    def _step_local(self, n: int) -> None:
        import torch

        if len(self.engines) == 1:
            self.engines[0].step(n)
            self._steps += n
            return
        for _ in range(n):
            par = self._steps % 2
            for e in self.engines:
                e.step_sweep()
            for e in self.engines:
                e.synchronize()
            flags = [f[par] for f in self._flags]
            merged = torch.stack(flags).amax(dim=0)
            for f in flags:
                f.copy_(merged)
            torch.cuda.synchronize()
            for e in self.engines:
                e.step_eval()
            for e in self.engines:
                e.synchronize()
            self._steps += 1

    def _step_distributed(self, n: int) -> None:
        import torch
        import torch.distributed as dist

        eng = self.engines[0]
        if self.sync == "p2p":
            eng.step_slab(n)  # enqueues everything; the slabs meet in peer memory
            self._steps += n
            eng.synchronize()
            return
        with torch.cuda.stream(self._stream):
            for _ in range(n):
                par = self._steps % 2
                dist.all_reduce(self._token)  # every slab has finished the previous step
                eng.step_sweep()
                dist.all_reduce(self._flags[0][par], op=dist.ReduceOp.MAX)
                eng.step_eval()
                self._steps += 1
        eng.synchronize()

    def step_timed(self, n: int) -> float:
        """n steps bracketed by CUDA events on this rank's engine stream; milliseconds."""
        import torch

        if not self.distributed:
            if len(self.engines) == 1:
                self._steps += n
                return self.engines[0].step_timed(n)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            self._step_local(n)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self._stream):
            a.record()
        if self.sync == "p2p":
            self.engines[0].step_slab(n)
            self._steps += n
        else:
            self._step_distributed(n)
        with torch.cuda.stream(self._stream):
            b.record()
        b.synchronize()
        self.engines[0].synchronize()  # also surfaces a handshake time-out
        return a.elapsed_time(b)

    # -- results ------------------------------------------------------------------------------------------
    def local_fire_map(self) -> np.ndarray:
        """int8 [E, rows held by this process, W]."""
        return np.concatenate([e.fire_map() for e in self.engines], axis=1)

    def fire_map(self) -> np.ndarray:
        """int8 [E, total_H, W] on every rank (gathers the slabs)."""
        local = self.local_fire_map()
        if not self.distributed:
            return local
        import torch.distributed as dist

        parts = [None] * self.ctx.world
        dist.all_gather_object(parts, local)
        return np.concatenate(parts, axis=1)

    def status(self):
        return self.engines[0].status()

    def close(self) -> None:
        self._sync_all()
        for p in self._opened:
            self.engines[0].ipc_close(p)
        self._opened = []
        for e in self.engines:
            e.close()
        self.engines = []
