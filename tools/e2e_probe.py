import os, sys, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import clip_common as C
from gomatching_b200.video.tracking import ClipTracker
cfg = C.L.build_cfg(device="cuda")
model = C.L.build_gomatching(cfg, seed=0, b200="transformer")
C.L.calibrate_detections(model, C.L.frames_to_inputs(C.L.synthetic_clip(1, 720, 1280, seed=1))[0], 40)
clip = C.L.synthetic_clip(8, 720, 1280, seed=11)
host = [torch.from_numpy(f).pin_memory() for f in clip]
dev = [f.cuda() for f in host]
def run(pool, host_results, assoc=True, n=80):
    ct = ClipTracker(model, overlap=True, host_results=host_results, associate=assoc)
    for i in range(12): ct.feed([pool[i % 8]])
    ct.flush(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): ct.feed([pool[i % 8]])
    t_spot = time.perf_counter() - t0
    ct.flush(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ct.drain(); ct.close()
    return dt / n * 1e3, t_spot / n * 1e3
for name, pool, hr, assoc in (("dev frames", dev, False, True), ("dev + host ids", dev, True, True), ("host frames", host, False, True),
                              ("host frames + host ids", host, True, True), ("dev spot only", dev, False, False), ("host spot only", host, False, False)):
    a, b = run(pool, hr, assoc)
    print("%-24s %.2f ms/frame total, %.2f ms/frame until the last feed returned" % (name, a, b))
