"""Clip driver for a reference-API video model: frame-sharded spotting, per-round record gather, and the reference's
sequential LST-Matcher association overlapped with the next round's spotting.

Takes the place of (reference, one process, one frame at a time):
    eval.py:329-345                       chunk loop, ``video_text_spotter(frames, instances, batch_id, id_count, ...)``
    GoMBatchPredictor.__call__            gomatching/text_track_visualizer.py:295-335   host preprocessing, last-batch
                                          ``_remove_short_track`` + ``batch_postprocess``
    GoMatching.batch_inference            gomatching/modeling/meta_arch/gom_lstmatcher.py:366-403

The model is used through the reference's own methods and nothing else:
    model.inference([frame], time_cost)            -> frame-local spotting (:268-351), N = 1 per forward
    model.run_short_term_match / run_long_term_match / test_len   (:405-564) via ``reference_association_step``
    model._remove_short_track, model.batch_postprocess            (:566-577, :353-364)
so the tracker code is the reference's, unchanged, and sees the same tensors in the same order as in the serial loop.

Schedule.  The clip is cut into ROUNDS.  In a round rank r spots ``weights[r]`` consecutive frames; the ranks' records
are gathered to the tracker rank (one collective per round), which appends them IN FRAME ORDER to a queue.  A worker
thread on the tracker rank -- its own CUDA stream, ``torch.no_grad`` -- pops frames and runs the association step, so
round k is associated while all ranks (the tracker rank included) spot round k+1.  ``weights`` lets the tracker rank
take fewer frames per round when the association is the longer pole (SURVEY.md s8e "scaling limiter").
"""
from __future__ import annotations

import queue
import sys
import threading
import time
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .association import FastAssociation
from .jpeg import decode_jpeg
from .pipeline import _world, reference_association_step
from .records import RecordSchema

__all__ = ["round_plan", "ClipTracker"]


def round_plan(n_frames: int, weights: Sequence[int]) -> List[List[Tuple[int, int, int]]]:
    """Frame -> (rank, slot) assignment.  Returns the rounds; a round is a list of (frame, rank, slot) in frame order.
    With all weights 1 this is ``frame t -> rank t mod W`` (sharding.frame_owner)."""
    if any(w < 0 for w in weights) or sum(weights) <= 0:
        raise ValueError("weights must be non-negative with a positive sum, got %r" % (list(weights),))
    per_round = sum(weights)
    owner: List[Tuple[int, int]] = []
    for r, w in enumerate(weights):
        owner.extend((r, s) for s in range(w))
    rounds = []
    for start in range(0, n_frames, per_round):
        rounds.append([(t, *owner[t - start]) for t in range(start, min(start + per_round, n_frames))])
    return rounds


_switch_lock = threading.Lock()
_switch_users = 0
_switch_saved: Optional[float] = None


def _switch_interval_acquire() -> None:
    """While any matcher thread lives, CPython's GIL switch interval is 0.5 ms (GOM_SWITCH_INTERVAL seconds; 0 = leave it);
    the previous value comes back with the last one's ``drain()``."""
    import os
    global _switch_users, _switch_saved
    si = float(os.environ.get("GOM_SWITCH_INTERVAL", "0.0005"))
    with _switch_lock:
        _switch_users += 1
        if _switch_users == 1 and si > 0 and sys.getswitchinterval() > si:
            _switch_saved = sys.getswitchinterval()
            sys.setswitchinterval(si)


def _switch_interval_release() -> None:
    global _switch_users, _switch_saved
    with _switch_lock:
        _switch_users = max(0, _switch_users - 1)
        if _switch_users == 0 and _switch_saved is not None:
            sys.setswitchinterval(_switch_saved)
            _switch_saved = None


class _JpegPrefetch:
    """Decode-ahead for JPEG inputs: the Huffman stage of ``decode_jpeg`` is host work (ctypes releases the GIL), so the
    frames a rank owns are decoded by a few worker threads, each on its own CUDA stream, while the rank spots earlier
    frames; the consumer waits on the frame's event."""

    def __init__(self, device, workers: int = 4, ahead: int = 8):
        from concurrent.futures import ThreadPoolExecutor
        self.device = device
        self.pool = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="jpeg-decode")
        self.local = threading.local()
        self.ahead = ahead

    def _decode(self, data):
        st = getattr(self.local, "stream", None)
        if st is None:
            st = self.local.stream = torch.cuda.Stream(device=self.device)
        with torch.cuda.device(self.device), torch.cuda.stream(st):
            img = decode_jpeg(data, self.device, bgr=True)
            ev = torch.cuda.Event()
            ev.record(st)
        return img, ev

    def submit(self, data):
        return self.pool.submit(self._decode, data)

    @staticmethod
    def take(fut):
        img, ev = fut.result()
        cur = torch.cuda.current_stream(img.device)
        cur.wait_event(ev)
        img.record_stream(cur)
        return img

    def shutdown(self):
        self.pool.shutdown(wait=False, cancel_futures=True)


class _Association(threading.Thread):
    """Tracker-rank worker: consumes gathered rounds in order and runs the reference's ID assignment."""

    def __init__(self, tracker: "ClipTracker"):
        super().__init__(daemon=True, name="lst-matcher")
        self.t = tracker
        self.q: "queue.Queue" = queue.Queue()
        self.error: Optional[BaseException] = None
        self.busy_s = 0.0
        # high priority: the matcher's many tiny kernels must not queue behind the replays of the frames in flight
        self.stream = torch.cuda.Stream(device=tracker.device, priority=-1) if tracker.device.type == "cuda" else None

    def run(self):
        try:
            with torch.no_grad():
                if self.stream is not None:
                    with torch.cuda.device(self.t.device), torch.cuda.stream(self.stream):
                        self._loop()
                else:
                    self._loop()
        except BaseException as e:  # surfaced by ClipTracker.finish()
            self.error = e

    def _loop(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            if isinstance(item, threading.Event):        # flush marker: everything queued before it is done
                item.set()
                continue
            records, frames, ready = item
            if ready is not None:
                self.stream.wait_event(ready)
            t0 = time.perf_counter()
            self.t._associate_round(records, frames)
            if self.stream is not None:
                self.stream.synchronize()
            self.busy_s += time.perf_counter() - t0


class ClipTracker:
    """One per process.  ``feed(frames)`` any number of times (a reference "chunk" or the whole clip), then
    ``finish()`` -> on the tracker rank the per-frame results ``[{"instances": Instances}, ...]`` with ``track_ids``
    (what GoMBatchPredictor returns after the last batch); None on the other ranks."""

    def __init__(self, model, schema: Optional[RecordSchema] = None, weights: Optional[Sequence[int]] = None,
                 tracker_rank: int = 0, overlap: bool = True, group=None, frame_size: Optional[Tuple[int, int]] = None,
                 use_batcher: Optional[bool] = None, input_format: str = "RGB", associate: bool = True,
                 host_results: bool = False, graph: Optional[bool] = None,
                 test_size: Optional[Tuple[int, int]] = None, fast_association: bool = True):
        self.model = model
        self.group = group
        self.rank, self.world = _world(group)
        self.weights = list(weights) if weights is not None else [1] * self.world
        if len(self.weights) != self.world:
            raise ValueError("need one weight per rank (%d), got %d" % (self.world, len(self.weights)))
        self.tracker_rank = tracker_rank
        self.device = torch.device(getattr(model, "device", "cpu"))
        self.use_batcher = self.device.type == "cuda" if use_batcher is None else use_batcher
        self.input_format = input_format
        mod = sys.modules[type(model).__module__]
        self._Instances, self._Boxes, self._ImageList = mod.Instances, mod.Boxes, mod.ImageList
        nq = int(model.cfg.MODEL.TRANSFORMER.NUM_QUERIES) if hasattr(model, "cfg") else 100
        self.schema = schema or RecordSchema(max_instances=nq)
        self.max_slots = max(self.weights)
        self.frame_size = frame_size
        # running state of the tracker rank (the arguments batch_inference threads through, eval.py:333-344)
        self.instances: list = []
        self.id_count = 0
        self.n_fed = 0
        # the reference's matchers with their bookkeeping on the host (video/association.py: identical IDs, ~3x faster);
        # False: GoMatching.run_short_term_match / run_long_term_match verbatim
        self._fast = FastAssociation(model) if fast_association and hasattr(model, "roi_heads") else None
        self.time_cost = {k: 0 for k in ("total_time", "pre_process", "backbone", "detector", "rescore", "tracker",
                                         "long_match", "short_match", "post_process")}
        self.spot_s = 0.0
        self.assoc_inline_s = 0.0
        # (MIN_SIZE_TEST, MAX_SIZE_TEST): resize every frame on the device the way the reference's predictor does on the
        # host (ResizeShortestEdge -> PIL bilinear, text_track_visualizer.py:318-319; video/resize.py is bit-identical);
        # None: frames are spotted at the size they come in
        self.test_size = test_size
        self.original_size: Optional[Tuple[int, int]] = None
        self.associate = associate            # False: records are gathered and dropped (spotting-only measurement)
        self.host_results = host_results      # True: every frame's track ids are copied to the host as they are assigned
        self.host_ids: list = []
        self._prefetch: Optional[_JpegPrefetch] = None
        self._worker: Optional[_Association] = None
        if overlap and self.rank == tracker_rank:
            self._worker = _Association(self)
            self._worker.start()
            # two Python threads now share the GIL (spotting + matcher); CPython hands it over only every 5 ms by default,
            # long enough for the replay queue of the spotting thread to run dry.  GOM_SWITCH_INTERVAL (seconds) overrides.
            _switch_interval_acquire()
        shared = model.__dict__.get("_msda_b200_spotter_graph")
        if shared is not None:
            shared.disable()                  # another ClipTracker of this model enabled it: re-point after the batcher below
        if self.use_batcher:
            self._install_batcher()
        # CUDA-graph replay of the static-shape part of the spotter (video/spotter_graph.py): needs the host-free
        # DeformableTransformer (install level "transformer") and the uint8 frame path
        self.spotter_graph = None
        tr = getattr(getattr(model, "detection_transformer", None), "transformer", None)
        if graph is None:
            graph = self.use_batcher and hasattr(tr, "_shape_tensors")
        if not graph and shared is not None and shared.graphs is not None:
            pass                              # stays disabled for this tracker's lifetime (eager spotting was asked for)
        if graph:
            if not self.use_batcher:
                raise ValueError("graph=True needs the uint8 frame path (use_batcher) on a CUDA device")
            from .spotter_graph import GraphedSpotter
            # one graphed spotter per model object: several ClipTrackers (concurrent clips) share its captures
            sg = model.__dict__.get("_msda_b200_spotter_graph")
            if sg is None:
                sg = GraphedSpotter(model, self._ImageList, input_format)
                model.__dict__["_msda_b200_spotter_graph"] = sg
            if sg.test_size != self.test_size:
                sg.test_size = self.test_size
                sg.reset()
            self.spotter_graph = sg
            sg.enable()

    # ------------------------------------------------------------------------------------------ spotting
    def _install_batcher(self):
        """Replace the model's host-side float conversion + eager normalise/pad (text_track_visualizer.py:315-321,
        gom_lstmatcher.py:159-170) by the frame-batcher kernel on the uint8 frame; values are bit-identical."""
        from .batcher import batch_frames
        from .resize import resize_frames_u8, shortest_edge_size

        cfg = self.model.cfg
        mean, std = list(cfg.MODEL.PIXEL_MEAN), list(cfg.MODEL.PIXEL_STD)
        flip = self.input_format == "RGB"
        ImageList = self._ImageList

        def preprocess_image(batched_inputs):
            frames = [x["image"] for x in batched_inputs]
            if any(f.dtype != torch.uint8 or f.dim() != 3 or f.shape[-1] != 3 for f in frames):
                raise ValueError("the B200 frame batcher takes uint8 (H, W, 3) frames as decoded")
            if len({tuple(f.shape) for f in frames}) != 1:
                raise ValueError("frames of one forward must share a size")
            stack = torch.stack(frames) if len(frames) > 1 else frames[0]
            if self.test_size is not None:
                h, w = frames[0].shape[:2]
                nh, nw = shortest_edge_size(h, w, self.test_size[0], self.test_size[1])
                stack = resize_frames_u8(stack, nh, nw)
            batch = batch_frames(stack, mean, std, flip_channels=flip)
            return ImageList(batch, [tuple(stack.shape[-3:-1])] * len(frames))

        self.model.preprocess_image = preprocess_image

    def _to_input(self, frame):
        """One frame as the predictor hands it to the model.  With the batcher: the uint8 HWC frame itself (BGR, as
        read); otherwise the reference's host conversion (optional RGB flip, float32 CHW)."""
        if self.use_batcher:
            if isinstance(frame, (bytes, bytearray, memoryview)):     # a JPEG file as read from disk: decoded on the device
                t = decode_jpeg(frame, self.device, bgr=True)
            elif hasattr(frame, "result"):                            # ... ahead of time by the prefetch threads
                t = _JpegPrefetch.take(frame)
            else:
                t = frame if isinstance(frame, torch.Tensor) else torch.from_numpy(frame)
            h, w = t.shape[:2]
            if self.original_size is None:
                self.original_size = (int(h), int(w))
            return {"image": t.to(self.device, non_blocking=True), "height": h, "width": w, "video_id": 0}
        if isinstance(frame, dict):
            return frame
        x = frame.numpy() if isinstance(frame, torch.Tensor) else frame
        if self.input_format == "RGB":
            x = x[:, :, ::-1]
        h, w = x.shape[:2]
        return {"image": torch.as_tensor(x.astype("float32").transpose(2, 0, 1)), "height": h, "width": w,
                "video_id": 0}

    def _spot(self, frame, prepared_input: bool = False) -> Tuple[Dict[str, torch.Tensor], Tuple[int, int]]:
        inst = self.model.inference([frame if prepared_input else self._to_input(frame)], self.time_cost)[0]
        fields = {k: (v.tensor if isinstance(v, self._Boxes) else v) for k, v in inst.get_fields().items()}
        return fields, tuple(int(s) for s in inst.image_size)

    # ------------------------------------------------------------------------------------------ association
    def _associate_round(self, records: torch.Tensor, frames: List[int]):
        if not self.associate:
            return
        for row, t in zip(records, frames):
            fields, frame_index, size = self.schema.unpack(row)
            assert frame_index == t == len(self.instances), "gather lost the frame order"
            inst = self._Instances(size)
            for k, v in fields.items():
                inst.set(k, self._Boxes(v) if k == "pred_boxes" else v)
            self.instances.append(inst)
            if self._fast is not None:
                self.instances, self.id_count = self._fast.step(self.instances, t, self.id_count)
            else:
                self.instances, self.id_count = reference_association_step(self.model, self.instances, t, self.id_count)
            if self.host_results:
                self.host_ids.append(self.instances[-1].track_ids.cpu())

    # ------------------------------------------------------------------------------------------ driver
    @torch.no_grad()
    def feed(self, frames: Sequence) -> None:
        """Spot and associate ``frames`` (numpy / torch uint8 HWC BGR as decoded, JPEG files as ``bytes`` -- decoded on the
        device, video/jpeg.py -- or, without the batcher, the reference's input dicts).  Every rank passes the same list; a rank only touches the frames it owns."""
        base = self.n_fed
        plan = round_plan(len(frames), self.weights)
        # JPEG inputs of this rank: decoded a few frames ahead by worker threads (host Huffman stage off the spotting thread)
        mine = [t for rnd in plan for t, r, _ in rnd if r == self.rank]
        pending = {}

        def top_up():
            pass
        if self.use_batcher and self.device.type == "cuda" and any(isinstance(frames[t], (bytes, bytearray, memoryview)) for t in mine):
            if self._prefetch is None:
                self._prefetch = _JpegPrefetch(self.device)
            order = [t for t in mine if isinstance(frames[t], (bytes, bytearray, memoryview))]
            nxt = [0]

            def top_up():
                while nxt[0] < len(order) and len(pending) < self._prefetch.ahead:
                    pending[order[nxt[0]]] = self._prefetch.submit(frames[order[nxt[0]]])
                    nxt[0] += 1
            top_up()
        # replay-ahead (video/spotter_graph.py): while frame t's eager tail runs, frame t + 1 is already on the GPU; the
        # graphed spotter is told which tensor inference() will ask for next
        sg = self.spotter_graph if (self.spotter_graph is not None and self.spotter_graph.enabled) else None
        nxt_of = {a: b for a, b in zip(mine, mine[1:])}
        prepared = {}

        def fetch(t):
            frame = frames[t]
            if t in pending:
                frame = pending.pop(t)
                top_up()
            return self._to_input(frame)

        for rnd in plan:
            t0 = time.perf_counter()
            block = self.schema.empty(self.max_slots, device=self.device)
            for t, r, s in rnd:
                if r == self.rank:
                    inp = prepared.pop(t) if t in prepared else fetch(t)
                    if sg is not None and isinstance(inp, dict) and torch.is_tensor(inp.get("image")):
                        ahead, a = [], t
                        for _ in range(sg.depth):
                            a = nxt_of.get(a)
                            if a is None:
                                break
                            if a not in prepared:
                                prepared[a] = fetch(a)
                            ahead.append(prepared[a]["image"])
                        sg.next_frames = ahead
                    fields, size = self._spot(inp, prepared_input=True)
                    self.schema.pack_into(block[s], fields, base + t, size)
            self.spot_s += time.perf_counter() - t0
            gathered = self._gather(block)
            if self.rank != self.tracker_rank:
                continue
            rows = torch.stack([gathered[r, s] for _, r, s in rnd]) if self.world > 1 else gathered[0, :len(rnd)]
            ids = [base + t for t, _, _ in rnd]
            if self._worker is not None:
                ready = None
                if self.device.type == "cuda":
                    ready = torch.cuda.Event()
                    ready.record()
                    rows.record_stream(self._worker.stream)
                self._worker.q.put((rows, ids, ready))
                if self._worker.error is not None:
                    raise self._worker.error
            else:
                t1 = time.perf_counter()
                self._associate_round(rows, ids)
                self.assoc_inline_s += time.perf_counter() - t1
        self.n_fed += len(frames)

    def _gather(self, block: torch.Tensor) -> Optional[torch.Tensor]:
        """(max_slots, stride) per rank -> (world, max_slots, stride) on the tracker rank."""
        if self.world == 1:
            return block.unsqueeze(0)
        out, bucket = None, None
        if self.rank == self.tracker_rank:
            out = torch.empty((self.world,) + tuple(block.shape), dtype=block.dtype, device=block.device)
            bucket = list(out.unbind(0))
        dist.gather(block, bucket, dst=self.tracker_rank, group=self.group)
        return out

    def flush(self) -> None:
        """Block until every frame fed so far has been associated; the worker stays alive (tracker rank; no-op
        elsewhere).  bench.py brackets its timed region with this."""
        if self._worker is not None:
            marker = threading.Event()
            self._worker.q.put(marker)
            while not marker.wait(0.05):
                if not self._worker.is_alive():
                    break
            if self._worker.error is not None:
                raise self._worker.error

    def close(self) -> None:
        """Give the model back its own forwards (undo the graph / batcher patches)."""
        if self._prefetch is not None:
            self._prefetch.shutdown()
            self._prefetch = None
        if self.spotter_graph is not None:
            self.spotter_graph.disable()
        self.model.__dict__.pop("preprocess_image", None)

    def drain(self) -> None:
        """Block until every fed frame has been associated (tracker rank; no-op elsewhere).  Ends the worker."""
        if self._worker is not None:
            self._worker.q.put(None)                     # in-order sentinel: everything queued before it is processed
            self._worker.join()
            err, self.assoc_worker_s = self._worker.error, self._worker.busy_s
            self._worker = None
            _switch_interval_release()
            if err is not None:
                raise err

    @torch.no_grad()
    def finish(self, image_size: Optional[Tuple[int, int]] = None):
        """Last-batch post-processing of GoMBatchPredictor.__call__ (text_track_visualizer.py:326-331)."""
        self.drain()
        if self.rank != self.tracker_rank:
            return None
        instances = self.instances
        if self.model.min_track_len > 0:
            instances = self.model._remove_short_track(instances)
        # GoMBatchPredictor scales the results back to the ORIGINAL frame size (text_track_visualizer.py:317, :329)
        size = image_size or self.original_size or (instances[0].image_size if instances else (0, 0))
        return self.model.batch_postprocess(instances, [size for _ in range(len(instances))])

    def association_seconds(self) -> float:
        return getattr(self, "assoc_worker_s", 0.0) + self.assoc_inline_s + (self._worker.busy_s if self._worker else 0.0)
