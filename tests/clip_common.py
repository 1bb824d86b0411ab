"""Shared helpers for the clip tests: the reference's OWN GoMatching (DeepSolo spotter + LST-Matcher) driven through
the sharded loop.  The reference Python is imported from /root/reference in the build container and from the staged,
unmodified copy under baseline/_ref on the GPU box (tools/refhost/loader.py)."""
import os
import socket
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tools.refhost import loader as L  # noqa: E402


def have_reference():
    return L.reference_root() is not None


def small_cfg(device="cpu", enc=1, dec=1, **kw):
    return L.build_cfg(device=device, MODEL__TRANSFORMER__ENC_LAYERS=enc, MODEL__TRANSFORMER__DEC_LAYERS=dec, **kw)


def reference_loop(model, frames, chunk=100, image_size=None):
    """The reference's own serial loop: eval.py:333-344 + GoMBatchPredictor.__call__ (text_track_visualizer.py:313-331)
    + GoMatching.batch_inference (gom_lstmatcher.py:366-403), chunked by ``chunk`` like eval.py:329."""
    h, w = frames[0].shape[:2]
    instances, id_count = [], 0
    n_chunks = (len(frames) + chunk - 1) // chunk
    assert chunk == 100 or n_chunks == 1, "batch_inference hard-codes start_frame_id = batch_id * 100 (gom_lstmatcher.py:368)"
    with torch.no_grad():
        for b in range(n_chunks):
            inputs = L.frames_to_inputs(frames[b * chunk:(b + 1) * chunk])
            instances, id_count = model.batch_inference(inputs, b, id_count, instances, L.new_time_cost())
        if model.min_track_len > 0:
            instances = model._remove_short_track(instances)
        return model.batch_postprocess(instances, [(h, w)] * len(instances)), id_count


def summarize(results):
    """Per frame: (track_ids, boxes, scores) as numpy -- what identity is judged on."""
    out = []
    for r in results:
        i = r["instances"]
        out.append((i.track_ids.cpu().numpy().copy(), i.pred_boxes.tensor.cpu().numpy().copy(),
                    i.scores.cpu().numpy().copy(), i.recs.cpu().numpy().copy()))
    return out


def assert_identical(a, b, what=""):
    assert len(a) == len(b), what
    for t, (x, y) in enumerate(zip(a, b)):
        assert np.array_equal(x[0], y[0]), "%s: track ids differ at frame %d" % (what, t)
        for u, v in zip(x[1:], y[1:]):
            assert np.array_equal(u, v), "%s: detections differ at frame %d" % (what, t)


def mot_scores(summary, gt_summary):
    """MOTA / IDF1 of a run against a synthetic ground truth made from ``gt_summary`` (the serial reference run):
    three quarters of its tracks are the objects, so the score is non-trivial (the rest count as false positives)."""
    from gomatching_b200.video import mot_metrics as MM

    acc = MM.MOTAccumulator()
    for t, ((ids, boxes, _, _), (gids, gboxes, _, _)) in enumerate(zip(summary, gt_summary)):
        keep = gids % 4 != 0
        acc.update([int(i) for i in gids[keep]], [int(i) for i in ids], MM.iou_distance(gboxes[keep], boxes), frameid=t)
    return acc.summary()


def same_scores(a, b):
    return a.keys() == b.keys() and all((a[k] == b[k]) or (a[k] != a[k] and b[k] != b[k]) for k in a)


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


REFCUDA = os.path.join(ROOT, "oracle", "_ref", "libmsda_refcuda.so")


def use_reference_cuda_kernel():
    """adet._C.ms_deform_attn_forward -> the unmodified reference CUDA kernel (tools/refhost/loader.py)."""
    return L.use_reference_cuda_kernel(REFCUDA)
