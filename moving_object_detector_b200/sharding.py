"""Frame-pair data parallelism across the GPUs of one box.

A frame pair (disp_prev, disp_now, flow, transform, dt) -> (cloud, mask, labels, objects) depends on no other pair
(the reference's only cross-frame state, "previous = last frame's now", scene_flow_constructor.cpp:398, is an
input), so a batch is split into CONTIGUOUS blocks of pairs, one block per rank: inside a block disp_prev of pair
i+1 is disp_now of pair i, so only one boundary frame per rank is duplicated.  Pixel data never leaves its GPU; the
only collective is an all-gather of the fixed-size per-frame MovingObject records and frame stats (NCCL on GPUs,
gloo in the CPU tests) -- the analogue of every constructor/clusterer pair publishing ~moving_objects to one tracker.
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
