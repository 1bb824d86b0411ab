#!/usr/bin/env python
"""Time the frame batcher against the reference's eager steps on the device (run on the GPU box).

    python tools/batcher_bench.py [--frames 8] [--height 720] [--width 1280]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gomatching_b200.video import batch_frames, padded_size  # noqa: E402

MEAN, STD = [123.675, 116.280, 103.530], [58.395, 57.120, 57.375]


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    a = ap.parse_args()
    n, h, w = a.frames, a.height, a.width
    hp, wp = padded_size(h, w, 32)
    sets = [torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda") for _ in range(6)]
    outs = [torch.empty(n, 3, hp, wp, device="cuda") for _ in range(6)]     # 6 x (n x 11 MB) > L2 at n = 8
    mean = torch.tensor(MEAN, device="cuda").view(3, 1, 1)
    std = torch.tensor(STD, device="cuda").view(3, 1, 1)
    i = [0]

    def ours():
        i[0] += 1
        batch_frames(sets[i[0] % 6], MEAN, STD, flip_channels=True, size_divisibility=32, out=outs[i[0] % 6])

    def eager():   # the device part of the reference only (its float32 CHW conversion runs on the host before the copy)
        i[0] += 1
        for f in sets[i[0] % 6]:
            t = f.flip(-1).permute(2, 0, 1).float()
            t = (t - mean) / std
            torch.nn.functional.pad(t, (0, wp - w, 0, hp - h))

    t_ours, t_eager = timeit(ours), timeit(eager)
    byts = n * (h * w * 3 + 3 * hp * wp * 4)
    print("frame batcher %d x %dx%d -> (%d,3,%d,%d): %.1f us = %.0f GB/s algorithmic | eager device ops %.1f us | "
          "H2D bytes per frame %.2f MB (uint8) vs %.2f MB (reference float32)" % (
              n, h, w, n, hp, wp, t_ours, byts / t_ours / 1e3, t_eager, h * w * 3 / 1e6, h * w * 12 / 1e6))


if __name__ == "__main__":
    main()
