"""Frame-sharded video inference around the MSDeformAttn hot path."""
from .batcher import batch_frames, padded_size
from .resize import resize_frames_u8, shortest_edge_size
from .jpeg import decode_jpeg, jpeg_size, read_image_bgr
from .tracking import ClipTracker, round_plan
from .gather import gather_records
from .pipeline import reference_association_step, run_clip, spot_chunk
from .records import GOMATCHING_FIELDS, Field, RecordSchema
from .writers import frame_rows, result_xml_name, write_track_transcriptions, write_video_results
from .sharding import CHUNK_FRAMES, chunk_ranges, frame_owner, frames_of_rank, slot_of_frame, slots_per_rank

__all__ = ["decode_jpeg", "jpeg_size", "read_image_bgr", "batch_frames", "padded_size", "resize_frames_u8", "shortest_edge_size", "ClipTracker", "round_plan", "gather_records", "reference_association_step", "run_clip", "spot_chunk", "GOMATCHING_FIELDS", "Field",
           "RecordSchema", "CHUNK_FRAMES", "chunk_ranges", "frame_owner", "frames_of_rank", "slot_of_frame",
           "slots_per_rank", "frame_rows", "result_xml_name", "write_track_transcriptions", "write_video_results"]
