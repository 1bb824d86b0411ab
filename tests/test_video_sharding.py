"""CPU tests of the frame-sharded video path: records, sharding, the gloo gather at
world_size 2, the association driver's control flow, and the MOTA / IDF1 restatement against the known answers
of the reference's own metric tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gomatching_b200 import video as V
from gomatching_b200.video import mot_metrics as MM


# ---------------------------------------------------------------------------------------------------
# records
# ---------------------------------------------------------------------------------------------------
def synth_fields(n, seed):
    g = torch.Generator().manual_seed(seed)
    return {
        "reid_features": torch.randn(n, 1024, generator=g),
        "pred_boxes": torch.rand(n, 4, generator=g) * 500,
        "scores": torch.rand(n, generator=g),
        "pred_classes": torch.zeros(n, dtype=torch.int64),
        "ctrl_points": torch.rand(n, 50, generator=g) * 700,
        "recs": torch.randint(0, 97, (n, 25), generator=g),
        "bd": torch.rand(n, 25, 4, generator=g),
    }


@pytest.mark.parametrize("n", [0, 1, 37, 100])
def test_record_roundtrip_is_bit_exact(n):
    schema = V.RecordSchema(max_instances=100)
    assert schema.row_bytes == (1024 + 4 + 1 + 50 + 100) * 4 + (1 + 25) * 8        # 4.9 KB per row
    f = synth_fields(n, 5 + n)
    row = schema.pack(f, frame_index=1234, image_size=(720, 1280))
    assert row.dtype == torch.uint8 and row.numel() == schema.stride and schema.stride % 16 == 0
    g, t, size = schema.unpack(row)
    assert t == 1234 and size == (720, 1280)
    for k in f:
        assert g[k].dtype == f[k].dtype and torch.equal(g[k], f[k]), k


def test_record_rejects_overflow_and_bad_shapes():
    schema = V.RecordSchema(max_instances=4)
    with pytest.raises(ValueError):
        schema.pack(synth_fields(5, 0), 0, (1, 1))
    f = synth_fields(2, 0)
    f["scores"] = f["scores"].double()
    with pytest.raises(ValueError):
        schema.pack(f, 0, (1, 1))


def test_sharding_is_a_partition_in_frame_order():
    for n in (0, 1, 7, 100):
        for w in (1, 2, 4, 8):
            owned = [V.frames_of_rank(n, r, w) for r in range(w)]
            assert sorted(sum(owned, [])) == list(range(n))
            assert all(len(o) <= V.slots_per_rank(n, w) for o in owned)
            for t in range(n):
                r, s = V.slot_of_frame(t, w)
                assert owned[r][s] == t and r == V.frame_owner(t, w)
    assert V.chunk_ranges(250) == [(0, 100), (100, 200), (200, 250)]


# ---------------------------------------------------------------------------------------------------
# a deterministic synthetic spotter + a sequential tracker (stand-ins for DeepSolo and the LST-Matcher)
# ---------------------------------------------------------------------------------------------------
N_OBJ = 9


def spot(frame, t):
    """Objects drift linearly; each has a fixed appearance vector; detections drop out pseudo-randomly."""
    g = torch.Generator().manual_seed(1000 + t)
    keep = torch.rand(N_OBJ, generator=g) > 0.2
    ids = torch.nonzero(keep).flatten()
    base = torch.Generator().manual_seed(7)
    app = torch.randn(N_OBJ, 1024, generator=base)
    pos0 = torch.rand(N_OBJ, 2, generator=base) * 400
    vel = torch.randn(N_OBJ, 2, generator=base) * 3
    c = pos0[ids] + vel[ids] * t
    boxes = torch.cat([c, c + 40], 1)
    n = len(ids)
    fields = {
        "reid_features": app[ids] + 0.05 * torch.randn(n, 1024, generator=g),
        "pred_boxes": boxes, "scores": torch.rand(n, generator=g), "pred_classes": torch.zeros(n, dtype=torch.int64),
        "ctrl_points": torch.rand(n, 50, generator=g), "recs": torch.randint(0, 97, (n, 25), generator=g),
        "bd": torch.rand(n, 25, 4, generator=g),
    }
    return fields, (720, 1280)


def associate(dets, chunk_start, state):
    """Greedy appearance tracker: strictly sequential, depends on every earlier frame -- like the real one."""
    state = state or {"next_id": 1, "gallery": {}, "tracks": [], "boxes": []}
    for d in dets:
        feats = d["fields"]["reid_features"]
        ids = []
        for f in feats:
            best, best_id = 0.5, None
            for tid, g in state["gallery"].items():
                s = float(torch.nn.functional.cosine_similarity(f, g, dim=0))
                if s > best and tid not in ids:
                    best, best_id = s, tid
            if best_id is None:
                best_id = state["next_id"]
                state["next_id"] += 1
            state["gallery"][best_id] = f
            ids.append(best_id)
        state["tracks"].append(ids)
        state["boxes"].append(d["fields"]["pred_boxes"].clone())
    return state


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, chunk, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        state = V.run_clip(list(range(n_frames)), spot, associate, V.RecordSchema(max_instances=16), chunk=chunk)
        if rank == 0:
            out.put((state["tracks"], [b.numpy() for b in state["boxes"]]))
        else:
            assert state is None
    finally:
        dist.destroy_process_group()


def run_world(world, n_frames, chunk):
    if world == 1:
        state = V.run_clip(list(range(n_frames)), spot, associate, V.RecordSchema(max_instances=16), chunk=chunk)
        return state["tracks"], [b.numpy() for b in state["boxes"]]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, _free_port_holder["p"], n_frames, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    return res


_free_port_holder = {"p": None}


@pytest.mark.parametrize("world", [2])
def test_sharded_clip_gives_identical_track_ids_and_metrics(world):
    n_frames, chunk = 23, 10                       # chunk tails: 10 + 10 + 3 frames, odd counts per rank
    _free_port_holder["p"] = _free_port()
    ref_tracks, ref_boxes = run_world(1, n_frames, chunk)
    got_tracks, got_boxes = run_world(world, n_frames, chunk)
    assert got_tracks == ref_tracks                # identical track IDs frame by frame
    for a, b in zip(ref_boxes, got_boxes):
        assert np.array_equal(a, b)
    # hence identical MOTA / IDF1 against a synthetic ground truth (the undropped objects, true identities)
    def score(tracks, boxes):
        acc = MM.MOTAccumulator()
        base = torch.Generator().manual_seed(7)
        torch.randn(N_OBJ, 1024, generator=base)
        pos0 = torch.rand(N_OBJ, 2, generator=base) * 400
        vel = torch.randn(N_OBJ, 2, generator=base) * 3
        for t, (ids, bx) in enumerate(zip(tracks, boxes)):
            c = (pos0 + vel * t).numpy()
            gt = np.concatenate([c, c + 40], 1)
            acc.update(list(range(N_OBJ)), ids, MM.iou_distance(gt, bx), frameid=t)
        return acc.summary()
    a, b = score(ref_tracks, ref_boxes), score(got_tracks, got_boxes)
    assert a == b
    assert 0.5 < a["mota"] <= 1.0 and 0.5 < a["idf1"] <= 1.0


# ---------------------------------------------------------------------------------------------------
# association driver: same control flow as GoMatching.batch_inference (gom_lstmatcher.py:373-402)
# ---------------------------------------------------------------------------------------------------
class FakeInstances:
    def __init__(self, n):
        self.reid_features = torch.zeros(n, 4)
        self.track_ids = None
        self.removed = []

    def __len__(self):
        return self.reid_features.shape[0]

    def remove(self, name):
        self.removed.append(name)


class FakeModel:
    test_len = 3

    def __init__(self, unmatched_at=()):
        self.calls = []
        self.unmatched_at = set(unmatched_at)
        self.frame = 0

    def run_short_term_match(self, inst, id_count=None):
        self.calls.append(("short", len(inst), id_count))
        n = len(inst[-1])
        if id_count is not None:
            inst[-1].track_ids = torch.arange(id_count, id_count + n)
            return inst, id_count + n
        cur = torch.arange(100 * self.frame, 100 * self.frame + n)
        if self.frame in self.unmatched_at:
            cur[0] = -1
        inst[-1].track_ids = cur.clone()
        return inst, cur

    def run_long_term_match(self, inst, k, id_count, cur_id):
        self.calls.append(("long", len(inst), k, id_count))
        inst[-1].track_ids = torch.where(cur_id < 0, torch.tensor(id_count), cur_id)
        return inst, id_count + 1


def test_reference_association_control_flow():
    model = FakeModel(unmatched_at={3})
    instances, id_count = [], 0
    for f in range(6):
        instances.append(FakeInstances(3))
        model.frame = f
        instances, id_count = V.reference_association_step(model, instances, f, id_count)
    assert instances[0].track_ids.tolist() == [1, 2, 3]
    assert model.calls == [
        ("short", 2, 4),                 # frame 1: with the running counter (gom_lstmatcher.py:380-383)
        ("short", 2, None),              # frame 2
        ("short", 2, None), ("long", 3, 2, 7),   # frame 3: a -1 -> long-term over the last test_len frames (:390-399)
        ("short", 2, None),
        ("short", 2, None),
    ]
    # reid features leave the window after test_len frames (:401-402)
    assert instances[0].removed == ["reid_features"] and instances[2].removed == ["reid_features"]
    assert instances[3].removed == []


# ---------------------------------------------------------------------------------------------------
# MOTA / IDF1 restatement vs the known answers of the reference's vendored motmetrics tests
# ---------------------------------------------------------------------------------------------------
def test_mota_motp_known_answer():
    """tools/Evaluation_Protocol_ArtVideo/motmetrics/tests/test_metrics.py:252-284"""
    nan = np.nan
    acc = MM.MOTAccumulator()
    acc.update([], [1, 2], [], frameid=0)
    acc.update([1, 2], [], [], frameid=1)
    acc.update([1, 2], [1, 2], [[1, 0.5], [0.3, 1]], frameid=2)
    acc.update([1, 2], [1, 2], [[0.2, nan], [nan, 0.1]], frameid=3)
    acc.update([1, 2], [1, 2], [[5, 1], [1, 5]], frameid=4)
    acc.update([], [], [], frameid=5)
    m = acc.summary()
    assert (m["num_matches"], m["num_false_positives"], m["num_misses"], m["num_switches"]) == (4, 2, 2, 2)
    assert (m["num_detections"], m["num_objects"], m["num_predictions"], m["num_frames"]) == (6, 8, 8, 6)
    assert m["mota"] == pytest.approx(1.0 - (2 + 2 + 2) / 8)
    assert m["motp"] == pytest.approx(11.1 / 6)


def test_switch_counting_known_answer():
    """test_metrics.py:287-325 (test_ids): 7 matches, 3 switches, MOTA 0.7, MOTP 1.6/10"""
    nan = np.nan
    acc = MM.MOTAccumulator()
    acc.update([], [], [], frameid=0)
    acc.update([1, 2], [1, 2], [[1, 0], [0, 1]], frameid=1)
    acc.update([1, 2], [1, 2], [[0.4, nan], [nan, 0.4]], frameid=2)
    acc.update([1, 2], [1, 2], [[0, 1], [1, 0]], frameid=3)
    acc.update([1, 2], [2, 3], [[1, 0], [0.4, 0.7]], frameid=4)
    acc.update([1, 3], [2, 3], [[1, 0], [0.4, 0.7]], frameid=5)
    acc.update([], [], [], frameid=6)
    m = acc.summary()
    assert (m["num_matches"], m["num_switches"], m["num_false_positives"], m["num_misses"]) == (7, 3, 0, 0)
    assert m["mota"] == pytest.approx(1.0 - 3 / 10) and m["motp"] == pytest.approx(1.6 / 10)


def test_figure3_known_answer_and_idf1():
    """test_metrics.py:328-346: MOTA 0.2; IDF1 by hand: IDTP 4, 20 objects, 4 predictions -> 8/24"""
    acc = MM.MOTAccumulator()
    for _ in range(4):
        acc.update([1, 2, 3, 4], [], [])
    for _ in range(4):
        acc.update([4], [4], [0])
    m = acc.summary()
    assert m["mota"] == pytest.approx(0.2)
    assert m["idtp"] == 4 and m["idf1"] == pytest.approx(2 * 4 / (20 + 4))
    # an identity swap halfway: MOTA sees 2 switches, IDF1 sees half of each track mismatched
    acc = MM.MOTAccumulator()
    for t in range(4):
        acc.update([1, 2], [1, 2], [[0.1, np.nan], [np.nan, 0.1]])
    for t in range(4):
        acc.update([1, 2], [2, 1], [[0.1, np.nan], [np.nan, 0.1]])
    m = acc.summary()
    assert m["num_switches"] == 2 and m["mota"] == pytest.approx(1 - 2 / 16)
    assert m["idtp"] == 8 and m["idf1"] == pytest.approx(0.5)


def test_gil_switch_interval_is_lowered_only_while_a_matcher_thread_lives():
    import sys
    from gomatching_b200.video import tracking
    before = sys.getswitchinterval()
    tracking._switch_interval_acquire()
    tracking._switch_interval_acquire()
    assert sys.getswitchinterval() <= min(before, 0.0005) + 1e-12
    tracking._switch_interval_release()
    assert sys.getswitchinterval() <= min(before, 0.0005) + 1e-12          # one user left
    tracking._switch_interval_release()
    assert abs(sys.getswitchinterval() - before) < 1e-12
