"""Fixed-stride per-frame detection records: what travels from the spotting ranks to the tracker rank.

The reference keeps a frame's detections in a detectron2 ``Instances`` with the fields set at
gomatching/modeling/meta_arch/gom_lstmatcher.py:340-349 --
    reid_features (n,1024) f32   pred_boxes (n,4) f32   scores (n,) f32   pred_classes (n,) i64
    ctrl_points (n,50) f32       recs (n,25) i64        bd (n,25,4) f32
with n <= NUM_QUERIES (100 for ICDAR15, 300 for DSText; configs/*.yaml).  Ragged n cannot go through one
collective, so a frame becomes ONE fixed-stride byte record: a 32-byte header (n, frame index, image height,
image width) followed by, field after field, ``max_instances`` rows of that field (rows >= n are zero).  4.9 KB
per row -> <= 0.49 MB per frame at 100 queries.  Packing is a handful of torch copies on
whatever device the fields live on; nothing is converted, so unpack(pack(x)) == x bit for bit.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Sequence, Tuple

import torch

HEADER_BYTES = 32   # 4 x int64: n, frame index, image height, image width


@dataclass(frozen=True)
class Field:
    name: str
    shape: Tuple[int, ...]      # per-instance trailing shape
    dtype: torch.dtype

    @property
    def row_bytes(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n * torch.empty((), dtype=self.dtype).element_size()


# the fields GoMatching.inference attaches to a frame's Instances (gom_lstmatcher.py:340-349)
GOMATCHING_FIELDS: Tuple[Field, ...] = (
    Field("reid_features", (1024,), torch.float32),
    Field("pred_boxes", (4,), torch.float32),
    Field("scores", (), torch.float32),
    Field("pred_classes", (), torch.int64),
    Field("ctrl_points", (50,), torch.float32),
    Field("recs", (25,), torch.int64),
    Field("bd", (25, 4), torch.float32),
)


class RecordSchema:
    def __init__(self, fields: Sequence[Field] = GOMATCHING_FIELDS, max_instances: int = 100):
        self.fields = tuple(fields)
        self.max_instances = int(max_instances)
        off = HEADER_BYTES
        self.offsets: Dict[str, int] = {}
        for f in self.fields:
            assert f.row_bytes % 4 == 0, "fields are 4-byte multiples so every block stays aligned"
            self.offsets[f.name] = off
            off += f.row_bytes * self.max_instances
        self.stride = (off + 15) // 16 * 16          # bytes per frame record
        self.row_bytes = sum(f.row_bytes for f in self.fields)

    # ------------------------------------------------------------------------------------------
    def empty(self, n_records: int, device="cpu") -> torch.Tensor:
        return torch.zeros((n_records, self.stride), dtype=torch.uint8, device=device)

    def pack_into(self, out_row: torch.Tensor, fields: Dict[str, torch.Tensor], frame_index: int,
                  image_size: Tuple[int, int]) -> None:
        """Write one frame's fields into ``out_row`` (uint8, (stride,))."""
        n = int(next(iter(fields.values())).shape[0]) if fields else 0
        if n > self.max_instances:
            raise ValueError("frame has %d instances, record holds %d" % (n, self.max_instances))
        out_row.zero_()
        head = torch.tensor([n, frame_index, int(image_size[0]), int(image_size[1])], dtype=torch.int64)
        out_row[:HEADER_BYTES].copy_(head.view(torch.uint8).to(out_row.device))
        for f in self.fields:
            t = fields[f.name]
            if tuple(t.shape) != (n,) + f.shape or t.dtype != f.dtype:
                raise ValueError("field %s: expected %s %s, got %s %s" % (f.name, (n,) + f.shape, f.dtype,
                                                                          tuple(t.shape), t.dtype))
            if n:
                o = self.offsets[f.name]
                out_row[o:o + n * f.row_bytes].copy_(t.contiguous().view(-1).view(torch.uint8))

    def pack(self, fields: Dict[str, torch.Tensor], frame_index: int, image_size: Tuple[int, int],
             device=None) -> torch.Tensor:
        dev = device if device is not None else (next(iter(fields.values())).device if fields else "cpu")
        row = torch.empty((self.stride,), dtype=torch.uint8, device=dev)
        self.pack_into(row, fields, frame_index, image_size)
        return row

    def unpack(self, row: torch.Tensor):
        """-> (fields dict, frame_index, (height, width)); tensors are views copied out of the record."""
        head = row[:HEADER_BYTES].cpu().view(torch.int64)
        n, frame_index, h, w = (int(v) for v in head)
        if n < 0 or n > self.max_instances:
            raise ValueError("corrupt record: n=%d" % n)
        out = {}
        for f in self.fields:
            o = self.offsets[f.name]
            raw = row[o:o + n * f.row_bytes].clone()
            out[f.name] = raw.view(f.dtype).view((n,) + f.shape)
        return out, frame_index, (h, w)
