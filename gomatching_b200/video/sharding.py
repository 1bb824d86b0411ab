"""Frame-wise sharding of a clip over the GPUs of one box.

``GoMatching.inference`` for frame t reads nothing from any other frame (gomatching/modeling/meta_arch/
gom_lstmatcher.py:268-351); only ID assignment is sequential (:373-402).  So frames are the independent unit:
frame t of a chunk goes to rank ``t mod W`` (round robin keeps arrival order close to consumption order on the
tracker rank).  Chunks follow the reference: 100 frames (eval.py:329; ``start_frame_id = batch_id * 100``,
gom_lstmatcher.py:368).
"""
from __future__ import annotations

from typing import List

CHUNK_FRAMES = 100


def frame_owner(frame_in_chunk: int, world_size: int) -> int:
    return frame_in_chunk % world_size


def frames_of_rank(n_frames: int, rank: int, world_size: int) -> List[int]:
    """Chunk-local indices of the frames ``rank`` spots."""
    return list(range(rank, n_frames, world_size))


def slots_per_rank(n_frames: int, world_size: int) -> int:
    """Record slots every rank contributes to the gather (equal on all ranks; the tail is padding)."""
    return (n_frames + world_size - 1) // world_size


def slot_of_frame(frame_in_chunk: int, world_size: int):
    """(rank, slot) of a frame's record in the gathered (world_size, slots, stride) buffer."""
    return frame_in_chunk % world_size, frame_in_chunk // world_size


def chunk_ranges(n_frames: int, chunk: int = CHUNK_FRAMES):
    return [(s, min(s + chunk, n_frames)) for s in range(0, n_frames, chunk)]
