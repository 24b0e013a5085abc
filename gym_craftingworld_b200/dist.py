"""Multi-GPU plumbing: one process per GPU, worlds sharded by global index, no data-path collective.

Worlds are independent (one reference env object = one world, ``ray.py:119-141``), so rank ``g`` of ``G`` simply owns
the global ids ``[g*N/G, (g+1)*N/G)``; Philox streams are keyed by the GLOBAL id, so a world's trajectory does not
depend on ``G``.  The only exchange is the episode-statistics vector (24 x int64): one SUM all-reduce every
``every`` steps, issued on a side stream behind an event so the step/render stream never waits on NCCL.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Contiguous slice of ``total`` worlds owned by ``rank`` (sizes differ by at most one)."""
    lo = total * rank // world
    hi = total * (rank + 1) // world
    return lo, hi - lo


class StatsReducer:
    """Periodic all-reduce of ``env.stats`` (backend-agnostic: NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, stats: torch.Tensor, every: int = 128, group=None, inline: bool = False):
        """``inline=True`` issues the all-reduce on the caller's stream instead of a side stream (it then costs its
        ~20-30 us latency once per ``every`` steps but never competes with the env kernel for SM slots)."""
        self.stats, self.every, self.group, self.inline = stats, int(every), group, bool(inline)
        self.global_stats = torch.zeros_like(stats)
        self._steps = 0
        self._side = torch.cuda.Stream(device=stats.device) if stats.is_cuda else None
        self._work = None

    def step(self):
        """Call once per env step; launches the reduction every ``every`` steps."""
        self._steps += 1
        if self._steps % self.every == 0:
            self.reduce_async()

    def reduce_async(self):
        if not (dist.is_available() and dist.is_initialized()):
            self.global_stats.copy_(self.stats)
            return
        if self._side is None:                                    # CPU tensors (gloo): asynchronous work handle
            self.global_stats.copy_(self.stats)
            self._work = dist.all_reduce(self.global_stats, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            return
        if self.inline:                                           # on the caller's stream, ordered with the env launches
            self.global_stats.copy_(self.stats, non_blocking=True)
            dist.all_reduce(self.global_stats, op=dist.ReduceOp.SUM, group=self.group)
            return
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.stats.device))
        with torch.cuda.stream(self._side):
            self._side.wait_event(ready)
            self.global_stats.copy_(self.stats, non_blocking=True)
            dist.all_reduce(self.global_stats, op=dist.ReduceOp.SUM, group=self.group)

    def total(self) -> torch.Tensor:
        """Global statistics vector; sums the device replicas when the reducer was given ``env.stats_raw``."""
        g = self.wait()
        return g.sum(dim=0) if g.dim() == 2 else g

    def wait(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self._side is not None and not self.inline:
            torch.cuda.current_stream(self.stats.device).wait_stream(self._side)
        return self.global_stats
