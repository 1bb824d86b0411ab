"""video/association.py (ID assignment with the bookkeeping on the host) against the reference's own
``run_short_term_match`` / ``run_long_term_match`` (gom_lstmatcher.py:405-564) on the reference's real LST-Matcher head:
identical ``track_ids`` for every frame, across detection counts that change from frame to frame (including empty
frames), tracks that vanish and come back inside the window, and every switch the matchers read
(``NOT_MULT_THRESH``, ``WITH_IOU``, ``DECAY_TIME``, ``MAX_CENTER_DIST``, ``OVERLAP_THRESH``)."""
import sys

import numpy as np
import pytest
import torch

import clip_common as C

pytestmark = pytest.mark.skipif(not C.have_reference(), reason="reference tree not available")


def _clip_of_instances(model, seed, n_frames, max_objects, allow_empty):
    """Synthetic detections with persistent identities: every object has a base re-id feature and a slowly moving box; it
    is visible in a frame with probability 0.7, so tracks break and resume inside the long-term window."""
    mod = sys.modules[type(model).__module__]
    g = torch.Generator().manual_seed(seed)
    H, W = 240, 320
    feat_dim = model.roi_heads.asso_head.fc_dim if hasattr(model.roi_heads.asso_head, "fc_dim") else 1024
    base = torch.randn(max_objects, feat_dim, generator=g)
    centre = torch.rand(max_objects, 2, generator=g) * torch.tensor([W * 0.8, H * 0.8]) + torch.tensor([W * 0.1, H * 0.1])
    velocity = torch.randn(max_objects, 2, generator=g) * 2.0
    size = torch.rand(max_objects, 2, generator=g) * 30 + 10
    frames = []
    for t in range(n_frames):
        visible = torch.rand(max_objects, generator=g) < 0.7
        if not allow_empty and not visible.any():
            visible[0] = True
        if allow_empty and t in (3, 4):
            visible[:] = False                                   # two consecutive empty frames
        idx = torch.nonzero(visible).flatten()
        idx = idx[torch.randperm(len(idx), generator=g)]          # detection order is not identity order
        n = len(idx)
        c = centre[idx] + velocity[idx] * t
        boxes = torch.cat([c - size[idx] / 2, c + size[idx] / 2], dim=1)
        inst = mod.Instances((H, W))
        inst.reid_features = base[idx] + 0.3 * torch.randn(n, feat_dim, generator=g)
        inst.pred_boxes = mod.Boxes(boxes)
        inst.scores = torch.rand(n, generator=g)
        inst.pred_classes = torch.zeros(n, dtype=torch.long)
        inst.ctrl_points = torch.rand(n, 50, generator=g)
        inst.recs = torch.randint(0, 37, (n, 25), generator=g)
        inst.bd = torch.rand(n, 25, 4, generator=g)
        frames.append(inst)
    return frames


def _copy(inst, mod):
    out = mod.Instances(inst.image_size)
    for k, v in inst.get_fields().items():
        out.set(k, mod.Boxes(v.tensor.clone()) if isinstance(v, mod.Boxes) else v.clone())
    return out


@pytest.fixture(scope="module")
def model():
    torch.set_num_threads(2)
    return C.L.build_gomatching(C.small_cfg(enc=1, dec=1), seed=0)


@pytest.mark.parametrize("flags", [
    {},
    {"not_mult_thresh": True},
    {"with_iou": False},
    {"decay_time": 0.9},
    {"max_center_dist": 0.5},
    {"overlap_thresh": 0.02, "max_center_dist": 4.0, "decay_time": 0.8},
    {"test_len": 3},
])
@pytest.mark.parametrize("seed,allow_empty", [(0, False), (1, True), (2, False)])
def test_fast_association_assigns_the_reference_ids(model, flags, seed, allow_empty):
    from gomatching_b200.video.association import FastAssociation
    from gomatching_b200.video.pipeline import reference_association_step
    mod = sys.modules[type(model).__module__]
    saved = {k: getattr(model, k) for k in flags}
    try:
        for k, v in flags.items():
            setattr(model, k, v)
        frames = _clip_of_instances(model, seed, n_frames=14, max_objects=9, allow_empty=allow_empty)
        with torch.no_grad():
            ref, ref_count = [], 0
            for t, f in enumerate(frames):
                ref.append(_copy(f, mod))
                ref, ref_count = reference_association_step(model, ref, t, ref_count)
            fast, got, got_count = FastAssociation(model), [], 0
            for t, f in enumerate(frames):
                got.append(_copy(f, mod))
                got, got_count = fast.step(got, t, got_count)
        assert got_count == ref_count
        long_term_used = 0
        for t, (a, b) in enumerate(zip(ref, got)):
            assert a.track_ids.dtype == b.track_ids.dtype and a.track_ids.device == b.track_ids.device
            assert torch.equal(a.track_ids, b.track_ids), "frame %d: %s vs %s" % (t, a.track_ids.tolist(), b.track_ids.tolist())
            assert np.array_equal(fast.ids_host[t], b.track_ids.numpy())
            assert a.has("reid_features") == b.has("reid_features")
            if t >= 2 and len(a) and int(a.track_ids.max()) > (max(int(x.track_ids.max()) if len(x) else 0 for x in ref[:t]) if t else 0):
                long_term_used += 1
        assert long_term_used > 0, "the clip never reached the long-term matcher"
        ids = torch.cat([x.track_ids for x in ref])
        assert len(torch.unique(ids)) < len(ids), "no track was ever continued: the threshold branch was not exercised"
    finally:
        for k, v in saved.items():
            setattr(model, k, v)
