"""One graphed 1280x720 frame of the clip workload inside a cudaProfilerStart/Stop range, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/clip_launches.py
(kernel nodes of the replayed CUDA graph are profiled one by one; the tracker's eager kernels of that frame follow)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import clip_common as C
from gomatching_b200.video.tracking import ClipTracker

eager = "--eager" in sys.argv
cfg = C.L.build_cfg(device="cuda")
model = C.L.build_gomatching(cfg, seed=0, b200=os.environ.get("CLIP_LEVEL", "heads"))
frames = [torch.from_numpy(f).cuda() for f in C.L.synthetic_clip(6, 720, 1280, seed=1)]
C.L.calibrate_detections(model, C.L.frames_to_inputs(C.L.synthetic_clip(1, 720, 1280, seed=1))[0], 40)
ct = ClipTracker(model, overlap=False, graph=not eager)
ct.feed(frames[:5])
torch.cuda.synchronize()
if "--graph-only" in sys.argv:                      # the replayed spotter graph alone (no eager tail, no tracker)
    g = next(iter(ct.spotter_graph.graphs.values()))[0]
    torch.cuda.cudart().cudaProfilerStart()
    g.replay(frames[5])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
else:
    torch.cuda.cudart().cudaProfilerStart()
    ct.feed(frames[5:6])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done; detections:", len(ct.instances[-1]))
