"""CLEAR-MOT (MOTA / MOTP) and ID (IDF1) measures, restated minimally.

Definitions follow the py-motmetrics copy the reference vendors for its offline evaluation
(tools/Evaluation_Protocol_ArtVideo/motmetrics/): event generation ``MOTAccumulator.update`` (mot.py:136-325:
carry established matches forward, then Kuhn-Munkres on the rest, SWITCH when a match contradicts the previous
one), ``mota`` (metrics.py:543-548), the global min-cost ID assignment (metrics.py:629-663) and
``idf1 = 2 IDTP / (num_objects + num_predictions)`` (metrics.py:726-729).  The vendored package no longer runs
under numpy 2 / pandas 3, hence this restatement; tests/test_video_sharding.py checks it against the known
answers of the reference's own tests (motmetrics/tests/test_metrics.py:252-346).

Used to state the end-to-end acceptance criterion of the sharded video path: identical track IDs, hence
identical MOTA / IDF1, for 1, 2, 4, 8 ranks.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, Hashable, Sequence

import numpy as np
from scipy.optimize import linear_sum_assignment


def _assign(costs: np.ndarray):
    """Min-cost assignment that never pairs a NaN entry (motmetrics lap.py: expensive-edge substitution)."""
    if costs.size == 0:
        return []
    finite = np.isfinite(costs)
    if not finite.any():
        return []
    big = (np.abs(costs[finite]).max() + 1.0) * (max(costs.shape) + 1) * 2.0
    filled = np.where(finite, costs, big)
    r, c = linear_sum_assignment(filled)
    return [(i, j) for i, j in zip(r, c) if finite[i, j]]


class MOTAccumulator:
    def __init__(self, max_switch_time: float = float("inf")):
        self.max_switch_time = max_switch_time
        self.m: Dict[Hashable, Hashable] = {}          # object -> hypothesis of its last match
        self.last_occurrence: Dict[Hashable, int] = {}
        self.frames = 0
        self.matches = self.switches = self.fp = self.misses = 0
        self.dist_sum = 0.0
        self.n_objects = self.n_predictions = 0
        self._o_frames = defaultdict(int)              # frames each object / hypothesis appears in
        self._h_frames = defaultdict(int)
        self._pair_frames = defaultdict(int)           # frames with a finite distance for (o, h)
        self._next_frame = 0

    def update(self, oids: Sequence, hids: Sequence, dists, frameid=None):
        oids = list(oids)
        hids = list(hids)
        d = np.atleast_2d(np.asarray(dists, dtype=float)).reshape(len(oids), len(hids)).copy()
        if frameid is None:
            frameid = self._next_frame
        self._next_frame = frameid + 1
        self.frames += 1
        self.n_objects += len(oids)
        self.n_predictions += len(hids)
        for o in oids:
            self._o_frames[o] += 1
        for h in hids:
            self._h_frames[h] += 1
        for i, j in zip(*np.where(np.isfinite(d))):
            self._pair_frames[(oids[i], hids[j])] += 1
        o_done = np.zeros(len(oids), bool)
        h_done = np.zeros(len(hids), bool)
        if len(oids) and len(hids):
            # 1. carry established correspondences forward
            for i, o in enumerate(oids):
                if o not in self.m:
                    continue
                js = [j for j, h in enumerate(hids) if not h_done[j] and h == self.m[o]]
                if js and np.isfinite(d[i, js[0]]):
                    j = js[0]
                    o_done[i] = h_done[j] = True
                    self.matches += 1
                    self.dist_sum += d[i, j]
            # 2. Kuhn-Munkres on what is left
            d[o_done, :] = np.nan
            d[:, h_done] = np.nan
            for i, j in _assign(d):
                o, h = oids[i], hids[j]
                is_switch = (o in self.m and self.m[o] != h
                             and abs(frameid - self.last_occurrence[o]) <= self.max_switch_time)
                if is_switch:
                    self.switches += 1
                else:
                    self.matches += 1
                self.dist_sum += d[i, j]
                o_done[i] = h_done[j] = True
                self.m[o] = h
        self.misses += int((~o_done).sum())
        self.fp += int((~h_done).sum())
        for o in oids:
            self.last_occurrence[o] = frameid
        return frameid

    # ------------------------------------------------------------------------------------------
    def id_measures(self):
        """Global min-cost ID assignment (metrics.py:629-663) -> idtp, idfp, idfn."""
        oids = sorted(self._o_frames, key=repr)
        hids = sorted(self._h_frames, key=repr)
        no, nh = len(oids), len(hids)
        if no == 0 and nh == 0:
            return 0, 0, 0
        oi = {o: i for i, o in enumerate(oids)}
        hi = {h: i for i, h in enumerate(hids)}
        fpm = np.zeros((no + nh, no + nh))
        fnm = np.zeros((no + nh, no + nh))
        fpm[no:, :nh] = np.nan
        fnm[:no, nh:] = np.nan
        for o, c in self._o_frames.items():
            fnm[oi[o], :nh] = c
            fnm[oi[o], nh + oi[o]] = c
        for h, c in self._h_frames.items():
            fpm[:no, hi[h]] = c
            fpm[hi[h] + no, hi[h]] = c
        for (o, h), ex in self._pair_frames.items():
            fpm[oi[o], hi[h]] -= ex
            fnm[oi[o], hi[h]] -= ex
        pairs = _assign(fpm + fnm)
        idfp = float(np.nansum([fpm[i, j] for i, j in pairs]))
        idfn = float(np.nansum([fnm[i, j] for i, j in pairs]))
        idtp = self.n_objects - idfn
        return idtp, idfp, idfn

    def summary(self) -> dict:
        det = self.matches + self.switches
        idtp, idfp, idfn = self.id_measures()
        q = lambda a, b: float(a) / b if b else float("nan")
        return {
            "num_frames": self.frames, "num_matches": self.matches, "num_switches": self.switches,
            "num_false_positives": self.fp, "num_misses": self.misses, "num_detections": det,
            "num_objects": self.n_objects, "num_predictions": self.n_predictions,
            "mota": 1.0 - q(self.misses + self.switches + self.fp, self.n_objects),
            "motp": q(self.dist_sum, det),
            "idtp": idtp, "idfp": idfp, "idfn": idfn,
            "idp": q(idtp, idtp + idfp), "idr": q(idtp, idtp + idfn),
            "idf1": q(2 * idtp, self.n_objects + self.n_predictions),
        }


def iou_distance(gt_boxes: np.ndarray, hyp_boxes: np.ndarray, max_iou: float = 0.5) -> np.ndarray:
    """1 - IoU for xyxy boxes, NaN where the distance exceeds ``max_iou`` (motmetrics distances.iou_matrix)."""
    gt = np.asarray(gt_boxes, float).reshape(-1, 4)
    hy = np.asarray(hyp_boxes, float).reshape(-1, 4)
    out = np.full((len(gt), len(hy)), np.nan)
    for i, a in enumerate(gt):
        for j, b in enumerate(hy):
            iw = min(a[2], b[2]) - max(a[0], b[0])
            ih = min(a[3], b[3]) - max(a[1], b[1])
            inter = max(iw, 0.0) * max(ih, 0.0)
            union = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
            dist = 1.0 - (inter / union if union > 0 else 0.0)
            if dist <= max_iou:
                out[i, j] = dist
    return out
