"""Frame-pair data parallelism across the GPUs of one box.

A frame pair (disp_prev, disp_now, flow, transform, dt) -> (cloud, mask, labels, objects) depends on no other pair
(the reference's only cross-frame state, "previous = last frame's now", scene_flow_constructor.cpp:398, is an
input), so a batch is split into CONTIGUOUS blocks of pairs, one block per rank: inside a block disp_prev of pair
i+1 is disp_now of pair i, so only one boundary frame per rank is duplicated.  Pixel data never leaves its GPU; the
only collective is an all-gather of the fixed-size per-frame MovingObject records and frame stats (NCCL on GPUs,
gloo in the CPU tests), once per batch of steps -- the analogue of every constructor/clusterer pair publishing
~moving_objects to one tracker.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_range(n_pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """[first, last) of the contiguous block of frame pairs owned by `rank`; blocks differ in size by at most one."""
    if world <= 0 or not (0 <= rank < world) or n_pairs < 0:
        raise ValueError(f"bad shard request: n_pairs={n_pairs} rank={rank} world={world}")
    base, extra = divmod(n_pairs, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def shard_frames(n_pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """[first, last) of the disparity FRAMES a rank needs: its pairs' frames plus the one shared boundary frame."""
    first, last = shard_range(n_pairs, rank, world)
    return (first, last + 1) if last > first else (first, first)


def padded_block(n_pairs: int, world: int) -> int:
    """Per-rank record count of the gathered buffers (all-gather needs equal sizes): ceil(n_pairs / world)."""
    return -(-n_pairs // world)


def gather_object_lists(objs, stats, group=None):
    """All-gathers per-frame object records.  objs: (block, max_objs * record_bytes) uint8 tensor, stats: (block,
    stats_bytes) uint8 tensor, both padded to padded_block() rows on every rank.  Returns (world, block, ...) tensors
    on the same device.  One collective per tensor; no pixel data."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    # concatenated layout (world * block, ...): the form both the NCCL and the gloo backend accept
    g_objs = torch.empty((world * objs.shape[0],) + tuple(objs.shape[1:]), dtype=objs.dtype, device=objs.device)
    g_stats = torch.empty((world * stats.shape[0],) + tuple(stats.shape[1:]), dtype=stats.dtype, device=stats.device)
    dist.all_gather_into_tensor(g_objs, objs.contiguous(), group=group)
    dist.all_gather_into_tensor(g_stats, stats.contiguous(), group=group)
    return g_objs.view((world,) + tuple(objs.shape)), g_stats.view((world,) + tuple(stats.shape))


class ObjectListGatherer:
    """The collective of the data-parallel path with its buffers allocated once.

    The send buffer is a ring of `ring` step slots; slot k holds the records one step's clusterer call wrote:
    `slot(k)` = (objs (block, objs_bytes), stats (block, stats_bytes)) uint8 views the clusterer writes into through
    their device pointers.  ONE all-gather moves the first `slots` ring entries of every rank (launch(slots)), so a
    stream of steps pays one collective per ring, not one per step: NCCL's CTAs spin on their peers while they hold SM
    slots, and a per-step collective overlapping the next step's kernels made every rank pay the slowest rank's tail
    (round 1: -6 % at eight GPUs).  On CUDA the collective runs on a side stream: launch() orders it after everything
    already enqueued on the current stream, later work overlaps it, and wait() must be called (on the current stream)
    before a send slot is overwritten and before results are read.  On CPU tensors (gloo, tests) everything is
    synchronous.  `objs` / `stats` are slot 0 (the one-slot form the tests and small drivers use).
    """

    def __init__(self, block: int, objs_bytes: int, stats_bytes: int, device, group=None, ring: int = 1):
        import torch
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.block, self.objs_bytes, self.stats_bytes, self.ring = block, objs_bytes, stats_bytes, ring
        self.slot_bytes = block * (objs_bytes + stats_bytes)
        self._send = torch.zeros(ring * self.slot_bytes, dtype=torch.uint8, device=device)
        self._recv = torch.zeros(self.world * ring * self.slot_bytes, dtype=torch.uint8, device=device)
        self.objs, self.stats = self.slot(0)
        self._cuda = self._send.is_cuda
        if self._cuda:
            self._side = torch.cuda.Stream(device=device)
            self._ready = torch.cuda.Event()
            self._done = torch.cuda.Event()
        self._pending = False
        self._slots = ring
        self.launches = 0

    def slot(self, k: int):
        """(objs, stats) views of ring slot k of the send buffer."""
        base = k * self.slot_bytes
        split = base + self.block * self.objs_bytes
        return (self._send[base:split].view(self.block, self.objs_bytes),
                self._send[split:base + self.slot_bytes].view(self.block, self.stats_bytes))

    def launch(self, slots: int = None):
        """All-gathers ring slots [0, slots) (default: the whole ring) of every rank."""
        import torch
        import torch.distributed as dist
        slots = self.ring if slots is None else slots
        assert 1 <= slots <= self.ring
        self._slots = slots
        n = slots * self.slot_bytes
        send, recv = self._send[:n], self._recv[:self.world * n]
        self.launches += 1
        if not self._cuda:
            dist.all_gather_into_tensor(recv, send, group=self.group)
            return
        self._ready.record()
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._ready)
            dist.all_gather_into_tensor(recv, send, group=self.group)
            self._done.record()
        self._pending = True

    def wait(self):
        """Makes the current stream wait for the collective launched last (no host synchronisation)."""
        import torch
        if self._cuda and self._pending:
            torch.cuda.current_stream().wait_event(self._done)
            self._pending = False

    def result(self, slot: int = 0):
        """(world, block, objs_bytes) and (world, block, stats_bytes) views of ring slot `slot` of the last gather;
        call wait() first."""
        assert 0 <= slot < self._slots
        per = self._recv[:self.world * self._slots * self.slot_bytes].view(self.world, self._slots, self.slot_bytes)
        one = per[:, slot]
        split = self.block * self.objs_bytes
        return (one[:, :split].reshape(self.world, self.block, self.objs_bytes),
                one[:, split:].reshape(self.world, self.block, self.stats_bytes))


def merge_gathered(g_objs: np.ndarray, g_stats: np.ndarray, n_pairs: int, obj_dtype, stats_dtype) -> List[np.ndarray]:
    """Host side of the gather: per-pair object arrays in global pair order (drops the padding rows)."""
    world = g_objs.shape[0]
    out: List[np.ndarray] = []
    for r in range(world):
        first, last = shard_range(n_pairs, r, world)
        st = np.frombuffer(np.ascontiguousarray(g_stats[r]).tobytes(), dtype=stats_dtype)
        for k in range(last - first):
            rec = np.frombuffer(np.ascontiguousarray(g_objs[r, k]).tobytes(), dtype=obj_dtype)
            out.append(rec[:max(0, min(int(st["n_objects"][k]), len(rec)))].copy())
    return out
