"""GPU probe of the clip workload: reference model (reference CUDA kernel) vs B200 drop-in on the same weights --
identity of detections / track ids and a coarse per-stage timing.  Development tool, not a bench.
    python tools/clip_probe.py [--frames 12] [--height 720] [--width 1280] [--enc 6] [--dec 6]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import clip_common as C  # noqa: E402
from gomatching_b200.video.tracking import ClipTracker  # noqa: E402


def stage_timer(model):
    acc = {}

    def hook(name, mod):
        def pre(m, a):
            torch.cuda.synchronize(); acc.setdefault(name, [0.0, 0]); acc[name].append(time.perf_counter())

        def post(m, a, o):
            torch.cuda.synchronize(); t0 = acc[name].pop(); acc[name][0] += time.perf_counter() - t0; acc[name][1] += 1
        return mod.register_forward_pre_hook(pre), mod.register_forward_hook(post)

    dt = model.detection_transformer
    hs = []
    for name, mod in (("backbone", model.backbone), ("input_proj", dt.input_proj[0]), ("encoder", dt.transformer.encoder),
                      ("decoder", dt.transformer.decoder), ("detection_transformer", dt),
                      ("roi_heads(asso_head)", model.roi_heads)):
        hs.extend(hook(name, mod))
    return acc, hs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--enc", type=int, default=6)
    ap.add_argument("--dec", type=int, default=6)
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--level", default="transformer")
    ap.add_argument("--graph", type=int, default=1)
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = False
    cfg = C.small_cfg(device="cuda", enc=a.enc, dec=a.dec)
    frames = C.L.synthetic_clip(a.frames, a.height, a.width, seed=1)
    ref_model = C.L.build_gomatching(cfg, seed=0)
    sd = {k: v.clone() for k, v in ref_model.state_dict().items()}
    if not a.skip_ref:
        assert C.use_reference_cuda_kernel()
        t0 = time.perf_counter()
        ref, ref_count = C.reference_loop(ref_model, frames)
        torch.cuda.synchronize()
        print("reference loop (reference CUDA kernel, eager layers): %.1f ms/frame" % ((time.perf_counter() - t0) / a.frames * 1e3))
        t0 = time.perf_counter()
        ref, ref_count = C.reference_loop(ref_model, frames)
        torch.cuda.synchronize()
        print("reference loop second pass: %.1f ms/frame" % ((time.perf_counter() - t0) / a.frames * 1e3))
        ref = C.summarize(ref)
    del ref_model
    model = C.L.build_gomatching(cfg, seed=0, b200=a.level, state_dict=sd)
    for ov in (False, True):
        for rep in range(2):
            ct = ClipTracker(model, overlap=ov, graph=bool(a.graph))
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ct.feed(frames)
            res = ct.finish()
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        got = C.summarize(res)
        print("B200 ClipTracker level=%s graph=%s overlap=%s: %.1f ms/frame (spot %.1f, association %.1f ms/frame) %s" % (
            a.level, a.graph, ov, dt / a.frames * 1e3, ct.spot_s / a.frames * 1e3, ct.association_seconds() / a.frames * 1e3,
            ("graph replays %d failed=%s" % (ct.spotter_graph.replays, ct.spotter_graph.failed)) if ct.spotter_graph else ""))
        ct.close()
        if not a.skip_ref:
            same_ids = all(np.array_equal(x[0], y[0]) for x, y in zip(ref, got)) and len(ref) == len(got)
            md = max(float(np.abs(x[1] - y[1]).max()) if x[1].shape == y[1].shape and x[1].size else 0.0 for x, y in zip(ref, got))
            print("   track ids identical to the reference loop: %s; n/frame %s; max |box diff| %.3g px; id_count %d vs %d" % (
                same_ids, [len(x[0]) for x in got][:6], md, ct.id_count, ref_count))
    acc, hs = stage_timer(model)
    ct = ClipTracker(model, overlap=False, graph=False)
    ct.feed(frames)
    ct.finish()
    for h in hs:
        h.remove()
    for k, v in acc.items():
        print("   stage %-24s %.2f ms/frame (%d calls)" % (k, v[0] / a.frames * 1e3, v[1]))


if __name__ == "__main__":
    main()
