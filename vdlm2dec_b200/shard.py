"""Channel sharding across the GPUs of one box.

The path has no exchange step: channels (with the stream they are demodulated from) are
independent, so stream s lives on GPU `s mod N`, every GPU runs the same kernel on its own
streams, and the host merges the completed blocks.  torch.distributed is plumbing only
(barrier, max-over-ranks of the device time, gathering the blocks); no NCCL call sits on
the data path.
"""
from __future__ import annotations

import numpy as np


def shard_streams(total_streams: int, rank: int, world: int) -> list[int]:
    """Global stream indices owned by `rank`: s mod world == rank."""
    return list(range(rank, total_streams, world))


def shard_channels(total_streams: int, ch_per_stream: int, rank: int, world: int) -> list[int]:
    """Global channel numbers owned by `rank` (channels follow their stream)."""
    return [s * ch_per_stream + k for s in shard_streams(total_streams, rank, world) for k in range(ch_per_stream)]


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def merge_blocks(per_rank_blocks: list[np.ndarray]) -> np.ndarray:
    """Union of the ranks' completed blocks in the order one GPU would have produced them
    (oldest trigger first, then channel number)."""
    allb = np.concatenate([b for b in per_rank_blocks if len(b)]) if any(len(b) for b in per_rank_blocks) else per_rank_blocks[0]
    order = np.lexsort((allb["chn"], allb["sync_dump"]))
    return allb[order]


def gather_blocks(blocks: np.ndarray) -> np.ndarray | None:
    """all ranks -> rank 0 (returns None elsewhere); single process: identity."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return merge_blocks([blocks])
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(blocks, out, dst=0)
    return merge_blocks(out) if dist.get_rank() == 0 else None


def merge_frames(per_rank_frames: list[np.ndarray]) -> np.ndarray:
    """Union of the ranks' frames (row f1: what the reference hands to out()) in the order vdl2_drain_frames uses on
    one GPU: oldest trigger first, then channel, then position inside the block.  `block` is rank-local and is reset
    to -1: chn / Fr / ppm / sync_dump travel in the frame."""
    parts = [f for f in per_rank_frames if len(f)]
    if not parts:
        return per_rank_frames[0]
    allf = np.concatenate(parts).copy()
    allf["block"] = -1
    return allf[np.lexsort((allf["len"], allf["chn"], allf["sync_dump"]))]


def gather_frames(frames: np.ndarray) -> np.ndarray | None:
    """all ranks -> rank 0 (returns None elsewhere); single process: identity."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return merge_frames([frames])
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(frames, out, dst=0)
    return merge_frames(out) if dist.get_rank() == 0 else None
