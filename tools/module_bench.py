#!/usr/bin/env python
"""Module-level timing: MSDeformAttn.forward (4 projections + glue + sampler) at the 720p encoder / decoder shapes.

    python tools/module_bench.py [--frames 8]
Rows: this repo with tcgen05 3xTF32 projections, this repo with F.linear (cuBLAS fp32), and the reference's eager
structure (separate projections, eager softmax / location glue, core kernel) for the same module.
"""
import argparse, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gomatching_b200 as g  # noqa: E402
from gomatching_b200 import synthetic as syn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=8)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = False
F = a.frames
shapes_l = syn.level_shapes(720, 1280, 4)
shapes = torch.as_tensor(shapes_l, dtype=torch.long, device=dev)
lsi = syn.level_start_index(shapes_l).to(dev)
S = int(shapes.prod(1).sum())
res = {}
for kind in ("encoder", "decoder"):
    torch.manual_seed(1)
    mod = g.MSDeformAttn(256, 4, 8, 4).to(dev).eval()
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.01)
        mod.attention_weights.weight.normal_(0, 0.05)
    if kind == "encoder":
        ref = syn.encoder_reference_points(shapes_l, 1).expand(F, -1, -1, -1).contiguous().to(dev)
    else:
        ref = syn.decoder_reference_points(torch.Generator().manual_seed(3), F, 100, 25, 4).to(dev)
    Lq = ref.shape[1]
    sets = [(torch.randn(F, Lq, 256, device=dev), torch.randn(F, S, 256, device=dev)) for _ in range(3)]

    def run(i):
        q, src = sets[i % len(sets)]
        return mod(q, ref, src, shapes, lsi, None)

    def timeit():
        with torch.no_grad():
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.iters):
                run(i)
            e1.record()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters * 1e3

    mod.tensor_core_projections, mod.use_fused, mod.merge_query_projections = True, True, True
    t_tc = timeit()
    with torch.no_grad():
        out_tc = run(0).clone()
    mod.tensor_core_projections = False
    t_cublas = timeit()
    with torch.no_grad():
        out_cb = run(0).clone()
    mod.use_fused, mod.merge_query_projections = False, False
    t_eager = timeit()
    err = float((out_tc - out_cb).abs().max() / out_cb.abs().max())
    res[kind] = {"tcgen05_3xtf32_us": t_tc, "cublas_fp32_fused_sampler_us": t_cublas, "reference_structure_us": t_eager,
                 "frames": F, "Lq": Lq, "max_norm_diff_tc_vs_cublas": err}
    print("%s F=%d Lq=%d: module forward  tcgen05 3xTF32 %.1f us | cuBLAS fp32 + fused sampler %.1f us | reference structure "
          "(cuBLAS fp32, eager glue, core kernel) %.1f us | tc vs cuBLAS max-norm diff %.2e" % (kind, F, Lq, t_tc, t_cublas, t_eager, err))

# ---- encoder layer (deformable_transformer.py:218-278): self-attention + residual/LayerNorm + feed-forward block ----
torch.manual_seed(2)
layer = g.DeformableTransformerEncoderLayer(256, 1024, 0.1, "relu", 4, 8, 4).to(dev).eval()
with torch.no_grad():
    layer.self_attn.sampling_offsets.weight.normal_(0, 0.01)
    layer.self_attn.attention_weights.weight.normal_(0, 0.05)
FL = min(F, 4)          # the 1024-wide hidden activation is 628 MB per frame
ref = syn.encoder_reference_points(shapes_l, 1).expand(FL, -1, -1, -1).contiguous().to(dev)
lsets = [(torch.randn(FL, S, 256, device=dev), torch.randn(FL, S, 256, device=dev) * 0.1) for _ in range(2)]


def lrun(i):
    src, pos = lsets[i % len(lsets)]
    return layer(src, pos, ref, shapes, lsi, None)


def ltime():
    with torch.no_grad():
        for i in range(2):
            lrun(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            lrun(i)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1e3


layer.tensor_core_ffn = layer.fused_add_norm = True
layer.self_attn.tensor_core_projections, layer.self_attn.use_fused, layer.self_attn.merge_query_projections = True, True, True
t_tc = ltime()
with torch.no_grad():
    o_tc = lrun(0).clone()
layer.tensor_core_ffn = layer.fused_add_norm = False
layer.self_attn.tensor_core_projections = False
t_cb = ltime()
with torch.no_grad():
    o_cb = lrun(0).clone()
layer.self_attn.use_fused, layer.self_attn.merge_query_projections = False, False
t_ref = ltime()
err = float((o_tc - o_cb).abs().max() / o_cb.abs().max())
res["encoder_layer"] = {"tcgen05_3xtf32_us": t_tc, "cublas_fp32_fused_sampler_us": t_cb, "reference_structure_us": t_ref,
                        "frames": FL, "max_norm_diff_tc_vs_cublas": err}
print("encoder LAYER F=%d (attention + LayerNorm + 256-1024-256 feed-forward): tcgen05 3xTF32 %.1f us | cuBLAS fp32 %.1f us | "
      "reference structure %.1f us | tc vs cuBLAS max-norm diff %.2e" % (FL, t_tc, t_cb, t_ref, err))
print(json.dumps(res))
