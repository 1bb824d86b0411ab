"""Tile-shape sweep of the persistent 3xTF32 GEMM against the tile-per-CTA kernels (CUDA-graph replay of 20 calls per
shape).  python tools/gemm_sweep.py [--rows 19160 153280]"""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gomatching_b200.projections import linear_3xtf32

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, nargs="*", default=[19160, 153280])
ap.add_argument("--shapes", type=str, default="256x256,384x256,1024x256r,256x1024,768x256")
a = ap.parse_args()


def timed(M, N, K, relu, env):
    for k in ("MSDA_GEMM_PERSISTENT", "MSDA_GEMM_BN", "MSDA_GEMM_SPLIT_GROUPS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    x = [torch.randn(M, K, device="cuda") for _ in range(4)]
    w = torch.randn(N, K, device="cuda") * 0.05
    b = torch.randn(N, device="cuda")
    y = [torch.zeros(M, N, device="cuda") for _ in range(4)]
    linear_3xtf32(x[0], w, b, out=y[0], relu=relu)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for i in range(20):
                linear_3xtf32(x[i % 4], w, b, out=y[i % 4], relu=relu)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 40 * 1e3
    ref = torch.addmm(b.double(), x[3].double(), w.double().t())
    if relu:
        ref = ref.clamp(min=0)
    err = float((y[3].double() - ref).abs().max())
    return us, err


for M in a.rows:
    for sh in a.shapes.split(","):
        relu = sh.endswith("r")
        N, K = (int(v) for v in sh.rstrip("r").split("x"))
        res = []
        us, err = timed(M, N, K, relu, {"MSDA_GEMM_PERSISTENT": "0"})
        res.append(("tile-per-CTA", us, err))
        us, err = timed(M, N, K, relu, {})
        res.append(("auto", us, err))
        for g in (1, 2):
            us, err = timed(M, N, K, relu, {"MSDA_GEMM_SPLIT_GROUPS": str(g)})
            res.append(("auto, %d splitter group(s)" % g, us, err))
        for bn in (32, 64, 96, 128, 192, 256):
            if N % bn:
                continue
            us, err = timed(M, N, K, relu, {"MSDA_GEMM_BN": str(bn)})
            res.append(("128 x %d" % bn, us, err))
        tensor_us = 3 * 2.0 * M * N * K / 1.1e15 * 1e6
        print("M=%6d N=%4d K=%4d relu=%d (3 TF32 passes at 1.1 PF/s: %.1f us):" % (M, N, K, relu, tensor_us), flush=True)
        for name, us, err in res:
            print("    %-28s %7.1f us   max err %.2e" % (name, us, err), flush=True)
