#!/usr/bin/env python
"""Run the encoder-layer drop-in a few times at the 720p shape -- the command ncu wraps for a per-kernel breakdown.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x/layer_launches.csv \
        python tools/layer_prof.py --frames 4
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gomatching_b200 as g  # noqa: E402
from gomatching_b200 import synthetic as syn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
shapes_l = syn.level_shapes(720, 1280, 4)
shapes = torch.as_tensor(shapes_l, dtype=torch.long, device=dev)
lsi = syn.level_start_index(shapes_l).to(dev)
S = int(shapes.prod(1).sum())
torch.manual_seed(2)
layer = g.DeformableTransformerEncoderLayer(256, 1024, 0.1, "relu", 4, 8, 4).to(dev).eval()
ref = syn.encoder_reference_points(shapes_l, 1).expand(a.frames, -1, -1, -1).contiguous().to(dev)
src = torch.randn(a.frames, S, 256, device=dev)
pos = torch.randn(a.frames, S, 256, device=dev) * 0.1
with torch.no_grad():
    for _ in range(a.iters):
        layer(src, pos, ref, shapes, lsi, None)
torch.cuda.synchronize()
print("done")
