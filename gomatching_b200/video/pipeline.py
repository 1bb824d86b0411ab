"""Sharded video inference: the reference's sequential loop split into a parallel half and a serial half.

Reference (one process, one frame at a time):
    GoMBatchPredictor.__call__      gomatching/text_track_visualizer.py:295-335
    GoMatching.batch_inference      gomatching/modeling/meta_arch/gom_lstmatcher.py:366-403
        for frame: inference([frame])  (:369-372, frame-independent)  ->  ID assignment (:373-402, sequential)

Here:
    spot(frame)            = GoMatching.inference([frame]) -- runs on rank ``t mod W``; this is where the B200
                             MSDeformAttn kernels run (12 calls per frame inside the frozen DeepSolo spotter)
    gather                 = per-chunk NCCL gather of fixed-stride records to the tracker rank
    associate(frames...)   = the reference's ID-assignment logic, UNCHANGED, on the tracker rank:
                             ``reference_association_step`` below drives the reference model's own
                             run_short_term_match / run_long_term_match in exactly the order batch_inference does.

Because every frame is still spotted with N=1 and the tracker sees the same tensors in the same order, track
IDs -- and therefore MOTA / IDF1 -- are identical to the single-process run (tests/test_video_sharding.py).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .gather import gather_records
from .records import RecordSchema
from .sharding import CHUNK_FRAMES, chunk_ranges, frames_of_rank, slots_per_rank

SpotFn = Callable[[object, int], Tuple[Dict[str, torch.Tensor], Tuple[int, int]]]
# spot(frame, global_frame_index) -> (fields of the frame's detections, (image_height, image_width))


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def spot_chunk(frames: Sequence[object], chunk_start: int, spot: SpotFn, schema: RecordSchema, device="cpu",
               group=None) -> torch.Tensor:
    """Spot this rank's frames of one chunk; returns the rank's (slots, stride) record block."""
    rank, world = _world(group)
    n = len(frames)
    block = schema.empty(slots_per_rank(n, world), device=device)
    for slot, t in enumerate(frames_of_rank(n, rank, world)):
        fields, image_size = spot(frames[t], chunk_start + t)
        schema.pack_into(block[slot], fields, chunk_start + t, image_size)
    return block


def run_clip(frames: Sequence[object], spot: SpotFn, associate: Callable[[List[dict], int, object], object],
             schema: Optional[RecordSchema] = None, chunk: int = CHUNK_FRAMES, tracker_rank: int = 0, device="cpu",
             group=None, state=None):
    """Whole-clip driver (the role of eval.py:329-345 + GoMBatchPredictor.__call__).

    associate(chunk_detections, chunk_start, state) -> state is called on the tracker rank only, once per chunk,
    with the chunk's frames in order: a list of {"fields": {...}, "frame_index": t, "image_size": (h, w)}.
    Returns the final state on the tracker rank and None elsewhere.
    """
    schema = schema or RecordSchema()
    rank, _ = _world(group)
    for start, end in chunk_ranges(len(frames), chunk):
        block = spot_chunk(frames[start:end], start, spot, schema, device=device, group=group)
        records = gather_records(block, end - start, dst=tracker_rank, group=group)
        if rank == tracker_rank:
            dets = []
            for row in records:
                fields, t, size = schema.unpack(row)
                dets.append({"fields": fields, "frame_index": t, "image_size": size})
            assert [d["frame_index"] for d in dets] == list(range(start, end)), "gather lost the frame order"
            state = associate(dets, start, state)
    return state if rank == tracker_rank else None


def reference_association_step(model, instances: list, real_frame_id: int, id_count: int):
    """ID assignment for the newest frame ``instances[real_frame_id]``, calling the REFERENCE model's own matchers.

    Control flow of GoMatching.batch_inference, gom_lstmatcher.py:373-402, restated (the matchers themselves --
    run_short_term_match :405-465, run_long_term_match :467-564 -- are the reference's code, untouched):
      frame 0      ids 1..n
      frame 1      short-term match against frame 0 with the running id counter
      frame t>=2   short-term match; if any detection stayed unmatched (-1) fall back to the long-term match over
                   the last TEST_LEN frames
      then drop reid features that left the window.
    """
    f = real_frame_id
    if f == 0:
        first = instances[0]
        first.track_ids = torch.arange(1, len(first) + 1, device=first.reid_features.device)
        id_count = len(first) + 1
    elif f == 1:
        instances[0:2], id_count = model.run_short_term_match(instances[0:2], id_count=id_count)
    else:
        instances[f - 1:f + 1], cur_id = model.run_short_term_match(instances[f - 1:f + 1])
        if -1 in cur_id:
            lo = max(0, f + 1 - model.test_len)
            instances[lo:f + 1], id_count = model.run_long_term_match(
                instances[lo:f + 1], k=min(model.test_len - 1, f), id_count=id_count, cur_id=cur_id)
    newest = instances[-1].track_ids
    assert len(newest) == len(torch.unique(newest))
    if f - model.test_len >= 0:
        instances[f - model.test_len].remove("reid_features")
    return instances, id_count
