"""CUDA-graph replay of the frame-local, static-shape part of ``GoMatching.inference``.

At one frame per forward (the reference's video loop, gom_lstmatcher.py:369-372, and the only batch size that keeps its
numerics) the spotter is launch-bound: about 1500 kernels per 1280x720 frame (ResNet-50, 6 + 6 transformer layers, the
heads), 29 ms of Python and launch overhead around 8 ms of GPU work (profiles/r02_clip_probe.txt).  Everything from the
uint8 frame to the detection heads' raw outputs has static shapes for a given frame size and -- with the host-free
``DeformableTransformer`` of ``transformer_dropin`` -- never touches the host, so it is captured ONCE per frame size
and replayed:

    frame (uint8 HWC, static buffer) -> frame-batcher kernel -> backbone + positional encodings -> input_proj ->
    encoder x6 -> proposals / top-k -> decoder x6 -> class / text / point / boundary heads -> rescoring head

The data-dependent tail of ``inference`` (score threshold, NMS, the association head's FC on the kept queries,
``Instances``) stays eager and stays the reference's code: the model object is not re-implemented, its sub-modules'
``forward`` attributes are pointed at the replayed buffers for the duration of a call, so ``GoMatching.inference``
itself (:268-351) still drives the frame.  The captured kernels are the ones eager execution launches, in the same
order, so the outputs are bit-identical to the eager path (tests/test_clip_gpu.py).

Replay-ahead.  The eager tail of frame t (1.7 ms of reference Python with a dozen synchronisations) used to leave the GPU
idle before frame t + 1's replay (7.3 ms) was even enqueued.  There are therefore ``depth + 1`` captured buffer sets per
frame size, used round-robin, and the replays run on ``depth`` streams of their own: when ``inference`` asks for frame t,
the frames ``ClipTracker.feed`` announced through ``next_frames`` (t + 1, t + 2) are copied into free sets and replayed at
once, so the GPU works on them -- two frames overlapping, which also fills the gaps between a replay's many small
kernels -- while the calling thread runs frame t's tail on its own stream.  A set is rewritten only after the tail has
cloned its outputs (``free`` event); the consumer waits for the set's ``done`` event.  Same kernels, same order per frame:
results unchanged.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .batcher import batch_frames
from .resize import resize_frames_u8, shortest_edge_size

__all__ = ["GraphedSpotter"]


class _FrameGraph:
    def __init__(self, owner: "GraphedSpotter", hw: Tuple[int, int]):
        model = owner.model
        h, w = hw
        dev = owner.device
        self.frame = torch.zeros((h, w, 3), dtype=torch.uint8, device=dev)

        nh, nw = (h, w) if owner.test_size is None else shortest_edge_size(h, w, owner.test_size[0], owner.test_size[1])
        self.size = (nh, nw)

        def body():
            frame = self.frame if (nh, nw) == (h, w) else resize_frames_u8(self.frame, nh, nw)      # Pillow-exact, in the graph
            img = batch_frames(frame, owner.mean, owner.std, flip_channels=owner.flip)
            images = owner.ImageList(img, [(nh, nw)])
            features, pos = model.backbone(images)
            out = model.detection_transformer(features, pos, model.backbone)
            re = model.roi_heads.rescoring_head(out["query_features"]) if model.with_rescore else None
            return img, out, re

        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(3):                       # lazy initialisation (cuDNN plans, weight splits, shape caches) happens here
                body()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        from .. import _native
        calls0 = _native.calls
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.img, self.out, self.re = body()
        self.calls = _native.calls - calls0          # kernel-launching C-ABI calls of this library inside one replay
        self.done: Optional[torch.cuda.Event] = None   # recorded on the replay stream after the newest replay into this set
        self.free: Optional[torch.cuda.Event] = None   # recorded by the consumer once it has cloned this set's outputs
        # The captured launches hold raw pointers to weight-derived tensors that live OUTSIDE the graph's memory pool
        # (TF32 weight splits, merged query projections): keep them alive for as long as the graph exists.  Weights are
        # baked in at capture time -- GoMatching's spotter is frozen; after changing them call GraphedSpotter.reset().
        from .. import projections
        self.keepalive = [list(projections._split_cache.values()),
                          [m._qproj_cache for m in model.modules() if getattr(m, "_qproj_cache", None) is not None]]

    def replay(self, frame: torch.Tensor):
        self.frame.copy_(frame, non_blocking=True)
        self.graph.replay()

    def launch(self, frame: torch.Tensor, stream: "torch.cuda.Stream"):
        """Copy ``frame`` in and replay on ``stream``, after the frame exists (it was produced on the current stream) and
        after the previous consumer of this buffer set has cloned its outputs."""
        cur = torch.cuda.current_stream(frame.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        stream.wait_event(ready)
        if self.done is not None:                    # the previous replay into this set may have run on the other stream (and,
            stream.wait_event(self.done)             # if its frame was dropped, was never consumed): never two writers at once
        if self.free is not None:
            stream.wait_event(self.free)
            self.free = None
        with torch.cuda.stream(stream):
            self.replay(frame)
            self.done = torch.cuda.Event()
            self.done.record(stream)
        frame.record_stream(stream)


class GraphedSpotter:
    """Attach to a reference-API GoMatching model on a CUDA device; ``enable()`` re-points the model's static-shape
    sub-module forwards at graph replays, ``disable()`` restores eager execution."""

    def __init__(self, model, image_list_cls, input_format: str = "RGB"):
        self.model = model
        self.ImageList = image_list_cls
        self.device = torch.device(model.device)
        cfg = model.cfg
        self.mean, self.std = list(cfg.MODEL.PIXEL_MEAN), list(cfg.MODEL.PIXEL_STD)
        self.flip = input_format == "RGB"
        self.test_size = None                        # (MIN_SIZE_TEST, MAX_SIZE_TEST) or None: set by ClipTracker
        self.graphs: Dict[Tuple[int, int], list] = {}      # frame size -> depth + 1 buffer sets, used round-robin
        self.turn: Dict[Tuple[int, int], int] = {}
        self.depth = 2                                     # replays in flight beyond the frame being consumed
        self.replay_streams: list = []                     # one per frame in flight: consecutive frames overlap on the GPU
        self.next_frames: list = []                        # set by ClipTracker.feed: the frames inference() will ask for next
        self.inflight: list = []                           # [(frame tensor, buffer set)] already replaying, in order
        self.replay_ahead = True
        self.prefetched = 0
        self.current: Optional[_FrameGraph] = None
        self.enabled = False
        self.replays = 0
        self.replayed_calls = 0                      # C-ABI kernel launches executed through graph replays so far
        self.failed: Optional[str] = None
        self._eager_preprocess = None

    # -- the four call sites of GoMatching.inference (gom_lstmatcher.py:274-292) --------------------------------------
    def _preprocess_image(self, batched_inputs):
        frames = [x["image"] for x in batched_inputs]
        if len(frames) != 1 or frames[0].dtype != torch.uint8 or frames[0].dim() != 3:
            raise ValueError("the graphed spotter takes one uint8 (H, W, 3) frame per forward")
        f = frames[0]
        hw = (int(f.shape[0]), int(f.shape[1]))
        sets = self.graphs.get(hw)
        if sets is None:
            if len(self.graphs) >= 4:                # a clip has one frame size; do not hoard activations of old ones
                self.reset()
            try:
                with self.eager():
                    sets = [_FrameGraph(self, hw) for _ in range(self.depth + 1 if self.replay_ahead else 1)]
            except Exception as e:                   # not capturable in this configuration: stay eager, say why
                self.failed = "%s: %s" % (type(e).__name__, e)
                self.disable()
                torch.cuda.synchronize(self.device)
                return self.model.preprocess_image(batched_inputs)
            self.graphs[hw] = sets
            self.turn[hw] = 0
        if not self.replay_streams:
            self.replay_streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, self.depth))]
        if self.inflight and self.inflight[0][0] is f:
            g = self.inflight.pop(0)[1]              # announced by an earlier call: already replaying
            self.prefetched += 1
        else:
            self.inflight = []                       # out of order (another clip's frame): whatever flies is dropped
            g = self._launch(f, sets, hw)
        ahead, self.next_frames = self.next_frames, []
        if self.replay_ahead and len(sets) > 1:
            for nf in ahead[:self.depth]:            # the next frames go to the GPU before this frame's tail starts
                if any(nf is x for x, _ in self.inflight):
                    continue
                if len(self.inflight) >= self.depth or (int(nf.shape[0]), int(nf.shape[1])) != hw or nf.dtype != torch.uint8:
                    break
                self.inflight.append((nf, self._launch(nf, sets, hw)))
        torch.cuda.current_stream(self.device).wait_event(g.done)
        self.replays += 1
        self.replayed_calls += g.calls
        self.current = g
        return self.ImageList(g.img, [g.size])

    def _launch(self, f, sets, hw) -> _FrameGraph:
        g = sets[self.turn[hw] % len(sets)]
        g.launch(f, self.replay_streams[self.turn[hw] % len(self.replay_streams)])
        self.turn[hw] += 1
        return g

    def _consumed(self):
        """The tail holds clones of everything it needs from the current buffer set: the set may be rewritten."""
        g = self.current
        if g is not None:
            g.free = torch.cuda.Event()
            g.free.record(torch.cuda.current_stream(self.device))

    def _backbone(self, images):
        return None, None                            # consumed only by detection_transformer, which is replayed too

    def _detection_transformer(self, features, pos, backbone):
        # fresh tensors: inference() and detection() modify their inputs in place (:590-603)
        out = {k: (v.clone() if v is not None else None) for k, v in self.current.out.items()}
        if not self.model.with_rescore:
            self._consumed()
        return out

    def _rescoring_head(self, query_features):
        re = self.current.re.clone()
        self._consumed()                              # called after detection_transformer (gom_lstmatcher.py:286-288)
        return re

    def reset(self):
        """Forget the captured graphs (after the model's weights changed); the next frame re-captures."""
        for st in self.replay_streams:
            st.synchronize()
        self.graphs.clear()
        self.turn.clear()
        self.current = None
        self.inflight = []
        self.next_frames = []

    def enable(self):
        if self.enabled:
            return
        m = self.model
        self._eager_preprocess = m.__dict__.get("preprocess_image")
        m.preprocess_image = self._preprocess_image
        m.backbone.forward = self._backbone
        m.detection_transformer.forward = self._detection_transformer
        if m.with_rescore:
            m.roi_heads.rescoring_head.forward = self._rescoring_head
        self.enabled = True

    def disable(self):
        if not self.enabled:
            return
        m = self.model
        if self._eager_preprocess is not None:
            m.preprocess_image = self._eager_preprocess
        else:
            m.__dict__.pop("preprocess_image", None)
        for mod in (m.backbone, m.detection_transformer, m.roi_heads.rescoring_head if m.with_rescore else None):
            if mod is not None:
                mod.__dict__.pop("forward", None)
        self.enabled = False
        self.inflight = []
        self.next_frames = []

    class _Eager:
        def __init__(self, sp):
            self.sp = sp

        def __enter__(self):
            self.was = self.sp.enabled
            self.sp.disable()

        def __exit__(self, *a):
            if self.was:
                self.sp.enable()

    def eager(self):
        """Context manager: run the model's own forwards (used while a graph is being captured)."""
        return GraphedSpotter._Eager(self)
