"""GPU: cProfile of the reference tracker's association step at a realistic detection count (calibrated threshold).
    python tools/tracker_profile.py"""
import os, sys, time, cProfile, pstats
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import clip_common as C
from gomatching_b200.video.tracking import ClipTracker
cfg = C.L.build_cfg(device="cuda")
model = C.L.build_gomatching(cfg, seed=0, b200="transformer")
C.L.calibrate_detections(model, C.L.frames_to_inputs(C.L.synthetic_clip(1, 720, 1280, seed=1))[0], 40)
frames = [torch.from_numpy(f).cuda() for f in C.L.synthetic_clip(8, 720, 1280, seed=11)]
ct = ClipTracker(model, overlap=False, graph=True)
rows = []
for t in range(40):
    inst = model.inference([ct._to_input(frames[t % 8])], ct.time_cost)[0]
    fields = {k: (v.tensor if isinstance(v, ct._Boxes) else v) for k, v in inst.get_fields().items()}
    rows.append(ct.schema.pack(fields, t, tuple(inst.image_size)))
torch.cuda.synchronize()
for t in range(16):
    ct._associate_round(rows[t][None], [t])
torch.cuda.synchronize()
pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
for t in range(16, 40):
    ct._associate_round(rows[t][None], [t])
torch.cuda.synchronize()
pr.disable(); dt = time.perf_counter() - t0
print("association: %.2f ms/frame, detections/frame %.1f" % (dt / 24 * 1e3, sum(len(x) for x in ct.instances[16:]) / 24))
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
