#!/usr/bin/env python
"""Launch the MSDeformAttn forward a few times with a pinned tuning -- the command ncu wraps.

    ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 3 -c 2 -o gpurun_out/prof \
        python tools/prof_one.py --kind encoder --frames 1 --tuning mode=1,tile_q=32,variant=4,ctas_per_sm=4
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gomatching_b200 as g  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="encoder")
ap.add_argument("--frames", type=int, default=1)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--dist", default="local")
ap.add_argument("--fused", type=int, default=0)
ap.add_argument("--tuning", default="")
ap.add_argument("--launches", type=int, default=5)
ap.add_argument("--ref", type=int, default=0, help="also launch the reference CUDA kernel once")
a = ap.parse_args()
dev = torch.device("cuda", 0)
tn = {k: int(v) for k, v in (kv.split("=") for kv in a.tuning.split(","))} if a.tuning else None
sets = [bench.device_workload(a.kind, a.frames, 300 + i, a.dist, dev) for i in range(4)]
for w in sets:
    w["loc"], w["attn"] = g.locations_softmax(w["shapes"], w["ref"], w["offsets"], w["logits"], 8)
    if a.dtype == "bf16":
        w["value"] = w["value"].to(torch.bfloat16)
torch.cuda.synchronize()
for i in range(a.launches):
    w = sets[i % len(sets)]
    if a.fused:
        g.ms_deform_attn_forward_fused(w["value"], w["shapes"], w["lsi"], w["ref"], w["offsets"], w["logits"], tuning=tn)
    else:
        g.ms_deform_attn_forward(w["value"], w["shapes"], w["lsi"], w["loc"], w["attn"], 64, tuning=tn)
torch.cuda.synchronize()
if a.ref:
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmsda_refcuda.so"))
    fn = lib.refcuda_msda_forward_f32_nomemset
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 7 + [ctypes.c_void_p] * 2
    w = sets[0]
    o = torch.empty(a.frames, w["Lq"], 256, device=dev)
    fn(w["value"].data_ptr(), w["shapes"].data_ptr(), w["lsi"].data_ptr(), w["loc"].data_ptr(), w["attn"].data_ptr(),
       a.frames, w["S"], 8, 32, 4, w["Lq"], 4, o.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
print("done")
