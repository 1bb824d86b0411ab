"""GPU acceptance tests of the video path with the REFERENCE's real model (north_star: "end-to-end MOTA/IDF1 on a
synthetic clip must be identical").

The clip: 24 seeded 1280x720 frames; the model: the reference's GoMatching (ResNet-50 + 6+6-layer DeepSolo +
LSTMatcher) at its own default initialisation (with WITH_RESR the rescoring head passes 30-100 detections per frame).
  reference arm   the reference's serial loop (GoMatching.batch_inference), its own modules, and the UNMODIFIED
                  reference CUDA kernel (oracle/_ref/libmsda_refcuda.so) behind adet._C.ms_deform_attn_forward
  B200 arm        ClipTracker (frame batcher kernel, sharded rounds, record gather, association thread) around the
                  same model object API, with the B200 operator installed at level "op" (kernel swap only: must be
                  IDENTICAL, bit for bit, down to every track ID) or "layers" (the whole drop-in stack: fused glue,
                  tensor-core projections, fused add+LayerNorm -- fp32-rounding-level differences in the features,
                  checked within tolerance; identity flips at near-ties are counted and bounded)
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import clip_common as C
from gomatching_b200.video.tracking import ClipTracker

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not C.have_reference(), reason="reference Python not staged (baseline/_ref)")]

H, W, N_FRAMES = 720, 1280, 24


@pytest.fixture(scope="module")
def reference_run():
    if not os.path.exists(C.REFCUDA):
        pytest.skip("oracle/_ref/libmsda_refcuda.so not built")
    cfg = C.L.build_cfg(device="cuda")
    frames = C.L.synthetic_clip(N_FRAMES, H, W, seed=1)
    model = C.L.build_gomatching(cfg, seed=0)
    assert C.use_reference_cuda_kernel()
    ref, id_count = C.reference_loop(model, frames)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    C.L.restore_reference_classes()
    return cfg, frames, sd, C.summarize(ref), id_count


def test_operator_swap_leaves_every_track_id_identical(reference_run):
    cfg, frames, sd, ref, id_count = reference_run
    n = [len(r[0]) for r in ref]
    assert min(n) >= 10 and id_count > max(n), (n, id_count)           # a non-trivial clip
    model = C.L.build_gomatching(cfg, seed=0, b200="op", state_dict=sd)
    for overlap in (False, True):
        ct = ClipTracker(model, overlap=overlap)                        # uint8 frames -> frame batcher kernel
        ct.feed(frames[:10])
        ct.feed(frames[10:])
        got = C.summarize(ct.finish())
        C.assert_identical(ref, got, "B200 operator, overlap=%s" % overlap)
        assert ct.id_count == id_count
        assert C.same_scores(C.mot_scores(ref, ref), C.mot_scores(got, ref))        # identical MOTA / IDF1


def _spotter_outputs(model, frame):
    with torch.no_grad():
        images = model.preprocess_image([frame])
        features, pos = model.backbone(images)
        out = model.detection_transformer(features, pos, model.backbone)
        out["re_pred_logits"] = model.roi_heads.rescoring_head(out["query_features"])
    return {k: v.double() for k, v in out.items() if v is not None}


@pytest.mark.parametrize("level", ["module", "layers", "heads"])
def test_dropin_module_and_layers_stay_within_tolerance(reference_run, level):
    """Levels "module" / "layers" / "heads" (fused glue, tensor-core projections, fused
    add+LayerNorm; "heads": every remaining nn.Linear of the spotter on the tensor-core GEMM) change the summation order of the dense projections, so the spotter's outputs move at fp32 rounding level
    (bar: 1e-4 of max|ref|, the operator's fp32 bar).  Discrete decisions downstream (NMS order, Hungarian) can then
    flip at near-ties -- which a default-initialised model has in abundance (near-duplicate boxes) -- so identity of
    track IDs is asserted for level "op" only; here the raw outputs are checked, and the clip-level effect is
    bounded and printed."""
    cfg, frames, sd, ref, id_count = reference_run
    ref_model = C.L.build_gomatching(cfg, seed=0, state_dict=sd)
    assert C.use_reference_cuda_kernel()
    inputs = C.L.frames_to_inputs(frames[:2])
    want = [_spotter_outputs(ref_model, f) for f in inputs]
    del ref_model
    model = C.L.build_gomatching(cfg, seed=0, b200=level, state_dict=sd)
    if level == "heads":
        from gomatching_b200 import TensorCoreLinear
        assert sum(isinstance(m, TensorCoreLinear) for m in model.detection_transformer.modules()) >= 10
        assert not any(isinstance(m, TensorCoreLinear) for m in model.roi_heads.asso_head.modules()), "the tracker runs verbatim"
    for f, w in zip(inputs, want):
        got = _spotter_outputs(model, f)
        for k in w:
            err = float((w[k] - got[k]).abs().max() / w[k].abs().max())
            assert err <= 1e-4, "%s: %s moved by %.3g of max|ref|" % (level, k, err)
    ct = ClipTracker(model, overlap=True)
    ct.feed(frames)
    got = C.summarize(ct.finish())
    assert len(got) == len(ref)
    n_ref, n_got = [len(a[0]) for a in ref], [len(b[0]) for b in got]
    same = sum(int(np.array_equal(a[0], b[0])) for a, b in zip(ref, got))
    sa, sb = C.mot_scores(ref, ref), C.mot_scores(got, ref)
    print("level %s: %d of %d frames with identical track ids; detections per frame %s vs %s; MOTA %.4f vs %.4f, "
          "IDF1 %.4f vs %.4f" % (level, same, len(ref), n_ref[:8], n_got[:8], sa["mota"], sb["mota"], sa["idf1"], sb["idf1"]))
    assert max(abs(a - b) for a, b in zip(n_ref, n_got)) <= 5
    assert np.isfinite(sb["mota"]) and np.isfinite(sb["idf1"])


def _nccl_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cfg = C.L.build_cfg(device="cuda:%d" % rank)
        model = C.L.build_gomatching(cfg, seed=0, b200="layers")
        frames = C.L.synthetic_clip(N_FRAMES, H, W, seed=1)
        ct = ClipTracker(model, overlap=True, weights=[1] + [2] * (world - 1))
        ct.feed(frames)
        res = ct.finish()
        if rank == 0:
            q.put((C.summarize(res), ct.id_count))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_nccl_sharded_clip_identical_to_one_gpu():
    """W GPUs (NCCL record gather) vs one GPU, full drop-in stack on both sides: identical track IDs."""
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    cfg = C.L.build_cfg(device="cuda:0")
    model = C.L.build_gomatching(cfg, seed=0, b200="layers")
    frames = C.L.synthetic_clip(N_FRAMES, H, W, seed=1)
    ct = ClipTracker(model, overlap=False)
    ct.feed(frames)
    one = C.summarize(ct.finish())
    one_count = ct.id_count
    del model, ct
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = C.free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, got_count = q.get()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    C.assert_identical(one, got, "NCCL world %d" % world)
    assert got_count == one_count


def test_host_free_transformer_and_graph_replay_are_bit_identical_to_eager_layers():
    """Install level "transformer" only removes host round trips, and the CUDA-graph replay launches the kernels eager
    execution launches: both must reproduce level "layers" bit for bit -- every detection, every track ID."""
    cfg = C.L.build_cfg(device="cuda")
    frames = C.L.synthetic_clip(12, H, W, seed=1)
    runs = {}
    sd = None
    for name, level, graph in (("layers", "layers", False), ("transformer", "transformer", False), ("graph", "transformer", True),
                               ("heads", "heads", False), ("heads_graph", "heads", True)):
        model = C.L.build_gomatching(cfg, seed=0, b200=level, state_dict=sd)
        if sd is None:
            sd = {k: v.clone() for k, v in model.state_dict().items()}
        ct = ClipTracker(model, overlap=False, graph=graph)
        ct.feed(frames)
        runs[name] = (C.summarize(ct.finish()), ct.id_count)
        if graph:
            assert ct.spotter_graph.failed is None and ct.spotter_graph.replays == len(frames)
        ct.close()
        del model, ct
    C.assert_identical(runs["layers"][0], runs["transformer"][0], "level transformer vs layers")
    C.assert_identical(runs["layers"][0], runs["graph"][0], "graph replay vs eager")
    C.assert_identical(runs["heads"][0], runs["heads_graph"][0], "level heads: graph replay vs eager")
    assert runs["layers"][1] == runs["transformer"][1] == runs["graph"][1]
    assert runs["heads"][1] == runs["heads_graph"][1]


def test_device_resize_matches_the_reference_predictor_path():
    """The reference resizes every frame on the host (ResizeShortestEdge -> PIL bilinear) before the model sees it
    (text_track_visualizer.py:318-322).  ClipTracker(test_size=...) does it on the device, inside the replayed graph:
    the detections must be identical to feeding host-resized frames, and results are reported at the original size."""
    from PIL import Image
    from gomatching_b200.video.resize import shortest_edge_size
    cfg = C.small_cfg(device="cuda", enc=2, dec=2)
    frames = C.L.synthetic_clip(6, 180, 320, seed=5)
    nh, nw = shortest_edge_size(180, 320, 250, 3000)
    host_resized = [np.asarray(Image.fromarray(f).resize((nw, nh), Image.BILINEAR)) for f in frames]
    model = C.L.build_gomatching(cfg, seed=0, b200="transformer")
    runs = []
    for fr, ts, graph in ((host_resized, None, False), (frames, (250, 3000), False), (frames, (250, 3000), True)):
        ct = ClipTracker(model, overlap=False, graph=graph, test_size=ts)
        ct.feed(fr)
        runs.append(C.summarize(ct.finish(image_size=(180, 320))))
        ct.close()
    C.assert_identical(runs[0], runs[1], "device resize (eager) vs host PIL resize")
    C.assert_identical(runs[0], runs[2], "device resize (graph) vs host PIL resize")


def test_jpeg_files_as_clip_input_give_the_results_of_the_reference_frame_read():
    """``ClipTracker.feed`` takes JPEG files as bytes and decodes them on the device (video/jpeg.py): the clip must come out
    exactly as when every frame is read the reference's way -- Pillow on the host (read_image, eval.py:327), BGR array."""
    import io
    from PIL import Image
    cfg = C.L.build_cfg(device="cuda")
    model = C.L.build_gomatching(cfg, seed=0, b200="transformer")
    frames = C.L.synthetic_clip(6, H, W, seed=5)
    files, arrays = [], []
    for f in frames:                                               # f: BGR uint8 HWC as read_image returns it
        buf = io.BytesIO()
        Image.fromarray(np.ascontiguousarray(f[:, :, ::-1])).save(buf, "JPEG", quality=90)
        files.append(buf.getvalue())
        arrays.append(np.ascontiguousarray(np.asarray(Image.open(io.BytesIO(files[-1])).convert("RGB"))[:, :, ::-1]))
    runs = []
    for inputs in (arrays, files):
        ct = ClipTracker(model, overlap=False)
        ct.feed(inputs)
        runs.append((C.summarize(ct.finish()), ct.id_count))
        ct.close()
    C.assert_identical(runs[0][0], runs[1][0], "JPEG bytes vs host-decoded arrays")
    assert runs[0][1] == runs[1][1]
