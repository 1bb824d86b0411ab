"""ID assignment for the tracker rank without the per-element device round trips.

``GoMatching.run_short_term_match`` / ``run_long_term_match`` (gom_lstmatcher.py:405-465, :467-564) assign track IDs with
Python loops over DEVICE tensors: ``trk_id not in cur_id`` for every detection of the window (:473), boolean-list
indexing of seven fields per window frame (:474-485), ``traj_score[i, j] > thresh`` and ``track_ids[i] = ...`` per matched
pair (:449-453, :551-559) -- about 2 600 ``.item()`` synchronisations and 12 ms per 1280x720 frame at 20 detections
(tools/tracker_profile.py), which makes the tracker, not the spotter, the limit of a clip.

This module restates the BOOKKEEPING of those two functions on the host and leaves every floating-point decision where
the reference computes it:

* the association scores come from the reference's own modules, called the same way (``roi_heads._forward_transformer``,
  ``_activate_asso``), the trajectory scores / last-box IoUs / distance gate are the same tensor expressions on the same
  device, and the Hungarian assignment is the same ``scipy.optimize.linear_sum_assignment`` on the same matrix;
* track IDs are mirrored in numpy arrays (this code assigns every one of them), so membership tests, keep masks,
  ``unique`` and the one-hot ``id_inds`` are integer work on the host -- exact by construction;
* the threshold test runs on the host copy of ``traj_score`` the Hungarian step already needs, in float32 like the
  device expression (``overlap_thresh * id_inds[:, j].sum()`` is an fp32 product of an fp32-rounded constant and an
  exact integer);
* list-of-index gathers over a contiguous range become slices made contiguous (same values, same layout for ``mm``).

Result: the same ``track_ids`` tensors, bit for bit (tests/test_clip_real_model.py, tests/test_clip_gpu.py compare
against the verbatim path on the reference's real model), at about a third of the time.  ``ClipTracker(fast_association=
False)`` keeps the verbatim path, which is also what the parity tests use as the checker.
"""
from __future__ import annotations

import sys
from typing import List, Optional, Tuple

import numpy as np
import torch
from scipy.optimize import linear_sum_assignment

__all__ = ["FastAssociation"]


class FastAssociation:
    """State: the host mirror of every frame's ``track_ids``.  ``step`` is a drop-in for
    ``pipeline.reference_association_step`` (control flow of ``GoMatching.batch_inference``, gom_lstmatcher.py:373-402)."""

    def __init__(self, model):
        self.model = model
        mod = sys.modules[type(model).__module__]
        self._Instances, self._Boxes, self._pairwise_iou = mod.Instances, mod.Boxes, mod.pairwise_iou
        self.ids_host: List[np.ndarray] = []

    # ------------------------------------------------------------------------------------------------------------
    def step(self, instances: list, real_frame_id: int, id_count: int):
        m, f = self.model, real_frame_id
        assert f == len(self.ids_host) == len(instances) - 1, "frames must be associated in order"
        newest = instances[f]
        dev = newest.reid_features.device
        if f == 0:                                                                     # :378-382
            n = len(newest)
            self.ids_host.append(np.arange(1, n + 1, dtype=np.int64))
            newest.track_ids = torch.arange(1, n + 1, device=dev)
            id_count = n + 1
        elif f == 1:                                                                   # :383-387
            ids, id_count, _ = self._short(instances[0], instances[1], self.ids_host[0], id_count)
            self.ids_host.append(ids)
            newest.track_ids = torch.tensor(ids, device=dev)
        else:                                                                          # :388-399
            ids, _, cur_id = self._short(instances[f - 1], instances[f], self.ids_host[f - 1], None)
            self.ids_host.append(ids)
            if (cur_id == -1).any():
                lo = max(0, f + 1 - m.test_len)
                k = min(m.test_len - 1, f)
                id_count = self._long(instances[lo:f + 1], self.ids_host[lo:f + 1], k, id_count, cur_id)
            newest.track_ids = torch.tensor(self.ids_host[f], device=dev)
        assert len(self.ids_host[f]) == len(np.unique(self.ids_host[f]))              # :400
        if f - m.test_len >= 0:                                                        # :401-402
            instances[f - m.test_len].remove("reid_features")
        return instances, id_count

    # ------------------------------------------------------------------------------------------------------------
    def _scores(self, instances, n_t, query_frame, short_term):
        """asso_output (n_k x N) and pred_boxes (N x 4): :409-416 / :490-497, the reference's modules."""
        rh = self.model.roi_heads
        reid_features = torch.cat([x.reid_features for x in instances], dim=0)[None]
        if short_term:
            asso_output, pred_boxes, _, _ = rh._forward_transformer(instances, reid_features, query_frame, short_term=True)
        else:
            asso_output, pred_boxes, _, _ = rh._forward_transformer(instances, reid_features, query_frame)
        asso_output = asso_output[-1].split(n_t, dim=1)
        asso_output = rh._activate_asso(asso_output)
        return torch.cat(asso_output, dim=1), pred_boxes

    def _match(self, asso_nonk, k_boxes, nonk_boxes, ids_host: np.ndarray, n_k: int, gate: bool):
        """Trajectory scores -> Hungarian -> thresholded IDs (:429-453 / :507-559).  Returns the int64 host array of the
        n_k query detections' IDs, -1 where nothing matched."""
        m = self.model
        dev = asso_nonk.device
        Np = int(ids_host.shape[0])
        unique_host = np.unique(ids_host)                                              # torch.unique: sorted
        id_inds_host = (unique_host[None, :] == ids_host[:, None]).astype(np.float32)   # Np x M
        id_inds = torch.from_numpy(id_inds_host).to(dev)
        traj_score = torch.mm(asso_nonk, id_inds)                                      # n_k x M
        if id_inds.numel() > 0:
            last_inds = (id_inds * torch.arange(Np, device=dev)[:, None]).max(dim=0)[1]
            last_boxes = nonk_boxes[last_inds]
            last_ious = self._pairwise_iou(self._Boxes(k_boxes), self._Boxes(last_boxes))
        else:
            last_ious = traj_score.new_zeros(traj_score.shape)
        if m.with_iou:
            traj_score = torch.max(traj_score, last_ious)
        if gate and m.max_center_dist > 0.:                                            # :534-547, long-term only
            k_ct = (k_boxes[:, :2] + k_boxes[:, 2:]) / 2
            k_s = ((k_boxes[:, 2:] - k_boxes[:, :2]) ** 2).sum(dim=1)
            nonk_ct = (nonk_boxes[:, :2] + nonk_boxes[:, 2:]) / 2
            dist = ((k_ct[:, None] - nonk_ct[None, :]) ** 2).sum(dim=2)
            norm_dist = dist / (k_s[:, None] + 1e-8)
            valid = norm_dist < m.max_center_dist
            valid_assn = torch.mm(valid.float(), id_inds).clamp_(max=1.).long().bool()
            traj_score = traj_score.masked_fill(~valid_assn, 0)                        # traj_score[~valid_assn] = 0
        neg = (-traj_score).cpu()                                                      # the one synchronisation
        match_i, match_j = linear_sum_assignment(neg)
        traj = -neg.numpy()                                                            # exact
        counts = id_inds_host.sum(axis=0, dtype=np.float32)                            # exact integers
        thr = np.float32(m.overlap_thresh)
        track_ids = np.full((n_k,), -1, dtype=np.int64)
        for i, j in zip(match_i, match_j):
            thresh = thr * counts[j] if not m.not_mult_thresh else thr
            if traj[i, j] > thresh:
                track_ids[i] = unique_host[j]
        return track_ids

    def _short(self, prev, cur, prev_ids: np.ndarray, id_count: Optional[int]):
        """run_short_term_match on (prev, cur): returns (cur's ids, id_count, unique ids of cur)."""
        n0, n1 = len(prev), len(cur)
        asso_output, pred_boxes = self._scores([prev, cur], [n0, n1], 1, True)
        asso_nonk = asso_output[:, :n0].contiguous()                                   # columns nonk_inds = 0 .. n0-1
        track_ids = self._match(asso_nonk, pred_boxes[n0:], pred_boxes[:n0], prev_ids, n1, gate=False)
        if id_count:                                                                   # :455-459
            for i in range(n1):
                if track_ids[i] < 0:
                    id_count = id_count + 1
                    track_ids[i] = id_count
        return track_ids, id_count, np.unique(track_ids)

    def _long(self, window: list, window_ids: List[np.ndarray], k: int, id_count: int, cur_id: np.ndarray) -> int:
        """run_long_term_match over the window; updates window_ids[k] (the newest frame) in place."""
        last = len(window) - 1
        assert k == last, "the query frame is the newest frame of the window"
        dev = window[0].reid_features.device
        subset, kept_ids = [], []
        reid_idx = None
        for idx, p in enumerate(window):                                               # :468-486
            if idx != last:
                keep = ~np.isin(window_ids[idx], cur_id)
                kept_ids.append(window_ids[idx][keep])
            else:
                keep = window_ids[idx] == -1
                reid_idx = keep
            inst = self._Instances(window[0].image_size)
            if keep.all():
                inst.reid_features, inst.pred_boxes = p.reid_features, p.pred_boxes
            else:
                sel = torch.from_numpy(np.nonzero(keep)[0]).to(dev)
                inst.reid_features = p.reid_features[sel]
                inst.pred_boxes = self._Boxes(p.pred_boxes.tensor[sel])
            subset.append(inst)
        n_t = [len(x) for x in subset]
        N, T = sum(n_t), len(n_t)
        asso_output, pred_boxes = self._scores(subset, n_t, k, False)
        n_k = n_t[k]
        Np = N - n_k
        ids_host = np.concatenate(kept_ids) if kept_ids else np.zeros((0,), dtype=np.int64)
        assert ids_host.shape[0] == Np
        asso_nonk = asso_output[:, :Np].contiguous()                                   # the query frame is the trailing block
        m = self.model
        if m.decay_time > 0:                                                           # :515-519
            dts = torch.cat([x.reid_features.new_full((len(x),), T - t - 2) for t, x in enumerate(subset) if t != k], dim=0)
            asso_nonk = asso_nonk * (m.decay_time ** dts[None, :])
        track_ids = self._match(asso_nonk, pred_boxes[Np:], pred_boxes[:Np], ids_host, n_k, gate=True)
        for i in range(n_k):                                                           # :557-560
            if track_ids[i] < 0:
                id_count = id_count + 1
                track_ids[i] = id_count
        window_ids[k][reid_idx] = track_ids                                            # :561
        return id_count
