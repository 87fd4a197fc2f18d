"""Sequence sharding for the multi-GPU front end: frames are independent and frame-to-frame
matching needs only (i-1, i), so every rank owns a contiguous range of the sequence plus a one-frame halo
(the frame before its range, re-extracted locally).  There is no collective in the data path; results are
concatenated by the caller (rank order = sequence order)."""
from __future__ import annotations


def shard_range(n_frames: int, rank: int, world: int):
    """Frames [start, stop) owned by `rank`, balanced to within one frame."""
    if not (0 <= rank < world) or n_frames < 0:
        raise ValueError("bad rank / world / n_frames")
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_with_halo(n_frames: int, rank: int, world: int):
    """(first frame to process, stop, number of leading halo frames whose results are dropped)."""
    start, stop = shard_range(n_frames, rank, world)
    halo = 1 if start > 0 and stop > start else 0
    return start - halo, stop, halo


def drop_halo(results: dict, halo: int) -> dict:
    """Per-frame result arrays of a shard processed with shard_with_halo -> the owned frames only."""
    return {k: v[halo:] for k, v in results.items()}
