"""GPU: where a graphed frame's time goes (replay GPU time, eager tail, record packing, association).
    python tools/clip_stages.py [--frames 24]"""
import argparse, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import clip_common as C
from gomatching_b200.video.tracking import ClipTracker

ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=24); a = ap.parse_args()
cfg = C.L.build_cfg(device="cuda")
model = C.L.build_gomatching(cfg, seed=0, b200="heads")
C.L.calibrate_detections(model, C.L.frames_to_inputs(C.L.synthetic_clip(1, 720, 1280, seed=1))[0], 40)
frames = [torch.from_numpy(f).cuda() for f in C.L.synthetic_clip(a.frames, 720, 1280, seed=1)]
ct = ClipTracker(model, overlap=False, graph=True)
ct.feed(frames[:4])                                   # capture + warm
sg = ct.spotter_graph
g = next(iter(sg.graphs.values()))[0]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10): g.graph.replay()
e1.record(); torch.cuda.synchronize()
print("graph replay GPU time: %.2f ms/frame" % (e0.elapsed_time(e1) / 10))
T = {"inference": 0.0, "pack": 0.0, "assoc": 0.0, "unpack": 0.0}
sch = ct.schema
for t, f in enumerate(frames[4:], start=4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    inst = model.inference([ct._to_input(f)], ct.time_cost)[0]
    torch.cuda.synchronize(); t1 = time.perf_counter()
    fields = {k: (v.tensor if isinstance(v, ct._Boxes) else v) for k, v in inst.get_fields().items()}
    row = sch.pack(fields, t, tuple(inst.image_size))
    torch.cuda.synchronize(); t2 = time.perf_counter()
    ct._associate_round(row[None], [t])
    torch.cuda.synchronize(); t3 = time.perf_counter()
    T["inference"] += t1 - t0; T["pack"] += t2 - t1; T["assoc"] += t3 - t2
n = len(frames) - 4
print({k: "%.2f ms/frame" % (v / n * 1e3) for k, v in T.items()}, "detections last frame:", len(ct.instances[-1]))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for t, f in enumerate(frames[4:12], start=len(ct.instances)):
    inst = model.inference([ct._to_input(f)], ct.time_cost)[0]
    fields = {k: (v.tensor if isinstance(v, ct._Boxes) else v) for k, v in inst.get_fields().items()}
    row = sch.pack(fields, t, tuple(inst.image_size))
    ct._associate_round(row[None], [t])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
