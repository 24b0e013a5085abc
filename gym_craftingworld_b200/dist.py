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


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int, sysfs: str = "/sys"):
    """Restrict this process (and the threads it creates later: the host transport's workers) to the CPUs of the NUMA node the GPU
    hangs off, when the machine has more than one node and the process is allowed to run there.  The host-buffer API polls
    and patches mapped pinned memory the GPU writes over PCIe; from the wrong socket every such access crosses the
    inter-socket link.  Returns the CPU set bound to, or None when nothing was changed (single node, unknown topology,
    empty intersection with the current affinity)."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device_index)
        pci = f"{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
        with open(os.path.join(sysfs, "bus/pci/devices", pci, "numa_node")) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(os.path.join(sysfs, f"devices/system/node/node{node}/cpulist")) as f:
            node_cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        target = node_cpus & allowed
        if not target or target == allowed:
            return None
        os.sched_setaffinity(0, target)
        return target
    except Exception:  # noqa: BLE001  (no sysfs, no such attribute: leave the affinity alone)
        return None


class StatsReducer:
    """Periodic all-reduce of the episode statistics (backend-agnostic: NCCL on GPUs, gloo in the CPU tests).

    ``stats`` must be the LIVE accumulator the kernels add to -- an env (its ``stats_raw`` is taken) or that tensor itself;
    ``env.stats`` is a fresh sum on every access and would freeze the reducer on one snapshot, so it is refused.

    Every reduction works on a SNAPSHOT: the accumulator is copied on the caller's (step) stream, i.e. between two step
    launches, so all counters of the snapshot belong to the same step; the all-reduce of the copy then runs on a side
    stream while later launches keep adding to the live accumulator.  Snapshots and results are double-buffered, and
    :meth:`wait` orders the reader's stream after the reduction it returns.
    """

    def __init__(self, stats, every: int = 128, group=None, inline: bool = False):
        """``inline=True`` issues the all-reduce on the caller's stream instead of a side stream (it then costs its
        ~20-30 us latency once per ``every`` steps but never competes with the env kernel for SM slots)."""
        if hasattr(stats, "stats_raw"):
            stats = stats.stats_raw
        if not isinstance(stats, torch.Tensor) or stats.dtype != torch.int64:
            raise TypeError("StatsReducer needs the env or its live int64 `stats_raw` tensor")
        if stats.dim() == 1 and stats.is_cuda:
            raise ValueError("pass the env (or env.stats_raw, int64[16, 24]): env.stats is a detached copy and would never change")
        self.stats, self.every, self.group, self.inline = stats, int(every), group, bool(inline)
        self._buf = [torch.zeros_like(stats), torch.zeros_like(stats)]
        self._cur = 0                       # buffer of the most recent reduction
        self._steps = self._last = self.reductions = 0
        self._side = torch.cuda.Stream(device=stats.device) if stats.is_cuda else None
        self._done = [None, None]           # CUDA events: reduction into buffer i has finished
        self._work = None

    @property
    def global_stats(self) -> torch.Tensor:
        return self._buf[self._cur]

    def step(self, n: int = 1):
        """Call after ``n`` env steps (a single step, or a replayed CUDA graph of ``n``); launches a reduction whenever ``every``
        more steps have accumulated.  Returns True when one was launched."""
        self._steps += int(n)
        if self._steps - self._last >= self.every:
            self._last = self._steps
            self.reduce_async()
            return True
        return False

    def reduce_async(self):
        nxt = self._cur ^ 1
        out = self._buf[nxt]
        self.reductions += 1
        ready = dist.is_available() and dist.is_initialized()
        if self._side is None:                                    # CPU tensors (gloo): asynchronous work handle
            if self._work is not None:
                self._work.wait()
            out.copy_(self.stats)
            self._work = dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group, async_op=True) if ready else None
            self._cur = nxt
            return
        main = torch.cuda.current_stream(self.stats.device)
        if self._done[nxt] is not None:
            main.wait_event(self._done[nxt])                      # the reduction that last used this buffer (two back) is over
        out.copy_(self.stats, non_blocking=True)                  # the snapshot: on the step stream, between two launches
        if self.inline or not ready:
            if ready:
                dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
            self._done[nxt] = main.record_event()
        else:
            snap = main.record_event()
            with torch.cuda.stream(self._side):
                self._side.wait_event(snap)
                dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
                self._done[nxt] = self._side.record_event()
        self._cur = nxt

    def total(self) -> torch.Tensor:
        """Global statistics vector ``int64[24]`` of the last reduction (the device replicas summed)."""
        g = self.wait()
        return g.sum(dim=0) if g.dim() == 2 else g

    def wait(self) -> torch.Tensor:
        """The result of the most recent reduction; the caller's stream is ordered after it."""
        if self._work is not None:
            self._work.wait()
            self._work = None
        ev = self._done[self._cur]
        if ev is not None:
            torch.cuda.current_stream(self.stats.device).wait_event(ev)
        return self._buf[self._cur]
