#!/usr/bin/env python
"""Bring-up check of the 3xTF32 tcgen05 projection GEMM against float64 (run on the GPU box under `timeout`)."""
import sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gomatching_b200.projections import linear_3xtf32

torch.manual_seed(0)
dev = torch.device("cuda", 0)
for (M, N, K, mask) in [(128, 32, 16, False), (128, 256, 256, False), (300, 256, 256, True), (19160, 384, 256, False),
                        (19160 * 8, 256, 256, True)]:
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    rz = (torch.rand(M, device=dev) < 0.1) if mask else None
    y = linear_3xtf32(x, w, b, rz)
    torch.cuda.synchronize()
    ref = (x.double() @ w.double().t() + b.double())
    if rz is not None:
        ref = ref.masked_fill(rz[:, None], 0.0)
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    y32 = torch.nn.functional.linear(x, w, b)
    err32 = float((y32.double() - ref.masked_fill(rz[:, None], 0.0) if rz is not None else y32.double() - ref).abs().max() / ref.abs().max()) if rz is None else float('nan')
    print("M=%d N=%d K=%d mask=%s  max-norm err vs fp64: 3xTF32 %.3e  (torch fp32 %.3e)" % (M, N, K, mask, err, err32), flush=True)
    assert err < 1e-5, err
# timing
M, N, K = 19160 * 8, 256, 256
x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / 16; b = torch.randn(N, device=dev)
outs = [torch.empty(M, N, device=dev) for _ in range(2)]
xs = [torch.randn(M, K, device=dev) for _ in range(2)]
for name, fn in (("3xTF32 tcgen05", lambda i: linear_3xtf32(xs[i % 2], w, b, out=outs[i % 2])),
                 ("torch fp32 (cuBLAS)", lambda i: torch.nn.functional.linear(xs[i % 2], w, b))):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(20):
        fn(i)
    e.record(); torch.cuda.synchronize()
    us = a.elapsed_time(e) / 20 * 1e3
    print("%-22s M=%d N=%d K=%d: %.1f us  %.1f GB/s algorithmic  %.1f TFLOP/s" % (name, M, N, K, us, (M * K + M * N) * 4 / us / 1e3, 2.0 * M * N * K / us / 1e6))
