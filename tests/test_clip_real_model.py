"""CPU tests of the sharded clip loop with the REFERENCE's real model: DeepSolo spotter (adet) + GoMatching +
LSTMatcher, default-initialised and seeded, detectron2 replaced by tools/refhost/d2_standins.py.

On the CPU the operator is the reference's own ``ms_deform_attn_core_pytorch`` (the B200 operator has no CPU path);
what is under test here is the HOST logic of row a12 / e: ``ClipTracker`` (sharding rounds, records, gather, the
association thread) must reproduce ``GoMatching.batch_inference`` exactly -- identical track-ID tensors, hence
identical MOTA / IDF1 -- for 1, 2, 4 and 8 ranks (gloo), any round weighting, with and without overlap.
The GPU counterparts (B200 operator vs the reference CUDA kernel, NCCL) are in tests/test_clip_gpu.py.
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import clip_common as C
from gomatching_b200.video.tracking import ClipTracker, round_plan

pytestmark = pytest.mark.skipif(not C.have_reference(), reason="reference tree not available (neither /root/reference nor baseline/_ref)")

H, W, N_FRAMES = 128, 192, 14
THREADS = 2


def test_round_plan_is_a_partition_in_frame_order():
    for n in (0, 1, 7, 23):
        for weights in ([1], [1, 1], [1, 2], [0, 1, 1], [1, 2, 2, 2], [1] * 8):
            rounds = round_plan(n, weights)
            flat = [x for r in rounds for x in r]
            assert [t for t, _, _ in flat] == list(range(n))
            for rnd in rounds:
                assert len(rnd) <= sum(weights)
                assert len({(r, s) for _, r, s in rnd}) == len(rnd)
                assert all(s < weights[r] for _, r, s in rnd)
    assert [(t, r) for t, r, _ in round_plan(5, [1, 1])[0]] == [(0, 0), (1, 1)]      # == frame t -> rank t mod W
    with pytest.raises(ValueError):
        round_plan(3, [0, 0])


def test_standins_follow_detectron2_semantics():
    from tools.refhost import d2_standins as D2

    a = D2.Boxes(torch.tensor([[0., 0., 2., 2.], [5., 5., 5., 5.]]))
    b = D2.Boxes(torch.tensor([[1., 1., 3., 3.], [5., 5., 5., 5.]]))
    iou = D2.pairwise_iou(a, b)
    assert torch.equal(iou, torch.tensor([[1. / 7., 0.], [0., 0.]]))           # degenerate boxes: 0, never NaN
    inst = D2.Instances((4, 6))
    inst.scores = torch.arange(3.)
    with pytest.raises(AssertionError):
        inst.boxes = torch.zeros(2, 4)                                          # length-checked
    inst.boxes = D2.Boxes(torch.zeros(3, 4))
    sub = inst[torch.tensor([True, False, True])]
    assert len(sub) == 2 and torch.equal(sub.scores, torch.tensor([0., 2.])) and len(sub.boxes) == 2
    assert len(D2.Boxes([])) == 0 and D2.Boxes([]).tensor.shape == (0, 4)
    il = D2.ImageList.from_tensors([torch.ones(3, 4, 5), torch.ones(3, 2, 7)])
    assert il.tensor.shape == (2, 3, 4, 7) and il.image_sizes == [(4, 5), (2, 7)] and il.tensor[1, :, 2:].sum() == 0


@pytest.fixture(scope="module")
def model_and_clip():
    torch.set_num_threads(THREADS)
    cfg = C.small_cfg()
    model = C.L.build_gomatching(cfg, seed=0)
    frames = C.L.synthetic_clip(N_FRAMES, H, W, seed=1)
    ref, id_count = C.reference_loop(model, frames)
    return model, frames, C.summarize(ref), id_count


def test_reference_model_runs_and_tracks(model_and_clip):
    model, frames, ref, id_count = model_and_clip
    assert sum(p.numel() for p in model.parameters()) > 60e6          # the real thing: R50 + DeepSolo + LST-Matcher
    assert len(ref) == N_FRAMES
    n = [len(r[0]) for r in ref]
    assert min(n) >= 20, n                                             # a non-trivial clip (~100 / frame)
    assert id_count > max(n)                                           # new identities appeared after frame 0
    ids0 = set(ref[0][0].tolist())
    assert len(ids0 & set(ref[-1][0].tolist())) >= 5                   # and tracks persist across the clip


@pytest.mark.parametrize("overlap", [False, True])
def test_clip_tracker_equals_reference_batch_inference(model_and_clip, overlap):
    model, frames, ref, id_count = model_and_clip
    ct = ClipTracker(model, overlap=overlap)
    ct.feed(frames[:5])
    ct.feed(frames[5:])                                                # chunk boundaries do not matter
    got = C.summarize(ct.finish())
    C.assert_identical(ref, got, "ClipTracker(overlap=%s)" % overlap)
    assert ct.id_count == id_count
    a, b = C.mot_scores(ref, ref), C.mot_scores(got, ref)
    assert C.same_scores(a, b) and 0 < a["mota"] < 1 and 0 < a["idf1"] < 1, a


def _worker(rank, world, port, weights, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(THREADS)          # same thread count as the serial run: CPU reductions are order-sensitive
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = C.L.build_gomatching(C.small_cfg(), seed=0)
        frames = C.L.synthetic_clip(N_FRAMES, H, W, seed=1)
        ct = ClipTracker(model, weights=weights, overlap=True)
        ct.feed(frames[:9])
        ct.feed(frames[9:])
        res = ct.finish()
        if rank == 0:
            q.put((C.summarize(res), ct.id_count))
        else:
            assert res is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,weights", [(2, None), (2, [1, 3]), (4, None), (4, [0, 1, 2, 1]), (8, None)])
def test_sharded_ranks_give_identical_track_ids_and_metrics(model_and_clip, world, weights):
    """SURVEY s8e determinism check: W in {2, 4, 8} bit-identical to W = 1, reference matchers on the tracker rank."""
    _, _, ref, id_count = model_and_clip
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = C.free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, weights, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, got_count = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    C.assert_identical(ref, got, "world %d weights %s" % (world, weights))
    assert got_count == id_count
    a, b = C.mot_scores(ref, ref), C.mot_scores(got, ref)
    assert C.same_scores(a, b) and np.isfinite(a["mota"]) and np.isfinite(a["idf1"])


def test_gomatching_pp_config_runs_through_the_same_loop():
    """BASELINE.json configs[4] model family: GoMatching++ (configs/GoMatching_PP_DSText.yaml -- 300 queries, no rescoring
    head, the SHA_FFN_CRSATTN matcher head).  Same ClipTracker, identical to the reference's batch_inference."""
    torch.set_num_threads(THREADS)
    cfg = C.L.build_cfg(config="GoMatching_PP_DSText", device="cpu", MODEL__TRANSFORMER__ENC_LAYERS=1,
                        MODEL__TRANSFORMER__DEC_LAYERS=1)
    model = C.L.build_gomatching(cfg, seed=0)
    assert type(model.roi_heads).__name__ == "SHA_FFN_CRSATTN" and not model.with_rescore
    frames = C.L.synthetic_clip(9, H, W, seed=1)
    C.L.calibrate_detections(model, C.L.frames_to_inputs(frames[:1])[0], 60)      # the class head's prior is 0.01: nothing passes 0.5
    ref, id_count = C.reference_loop(model, frames)
    assert min(len(r["instances"]) for r in ref) >= 10
    ct = ClipTracker(model, overlap=True)
    assert ct.schema.max_instances == 300
    ct.feed(frames)
    C.assert_identical(C.summarize(ref), C.summarize(ct.finish()), "GoMatching++")
    assert ct.id_count == id_count
