"""Time the 3xTF32 tcgen05 GEMM at the shapes one DeepSolo frame uses (CUDA-graph replay of 20 calls per shape, so the
launch overhead of the Python wrapper is out of the number).  python tools/gemm_bench.py [--rows 19160]"""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gomatching_b200.projections import linear_3xtf32

ap = argparse.ArgumentParser(); ap.add_argument("--rows", type=int, nargs="*", default=[19160, 2500, 153280]); a = ap.parse_args()
for M in a.rows:
    for N, K, relu in ((256, 256, False), (384, 256, False), (1024, 256, True), (256, 1024, False), (768, 256, False)):
        x = [torch.randn(M, K, device="cuda") for _ in range(4)]
        w = torch.randn(N, K, device="cuda") * 0.05
        b = torch.randn(N, device="cuda")
        y = [torch.empty(M, N, device="cuda") for _ in range(4)]
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
        flop = 2.0 * M * N * K
        err = float((y[0].double() - torch.addmm(b.double(), x[0].double(), w.double().t()).clamp(min=0 if relu else -1e30)).abs().max())
        print("M=%6d N=%4d K=%4d relu=%d : %7.1f us  %6.1f TFLOP/s fp32-equivalent (%5.0f TF/s of TF32 MMA)  %5.0f GB/s  max err %.2e" % (
            M, N, K, relu, us, flop / us / 1e6, 3 * flop / us / 1e6, (M * K + M * N + 2 * N * K) * 4 / us / 1e3, err))
