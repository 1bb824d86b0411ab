"""Generate the golden fixtures in tests/golden/ by running the REFERENCE's own code.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What runs is the reference, unmodified, imported from where it lies:
  * third_party/adet/layers/ms_deform_attn.py   ms_deform_attn_core_pytorch (:40-60), MSDeformAttn (:63-156)
  * third_party/adet/layers/deformable_transformer.py  DeformableTransformer encoder/decoder (:22-497)
The package __init__ files pull in detectron2 (absent here), so `adet`, `adet.layers`, ... are
registered as bare namespace modules first, and `adet._C` (the CUDA extension, which has no CPU
implementation: csrc/DeformAttn/ms_deform_attn.h:38) is a stub whose ms_deform_attn_forward routes to
the reference's own ms_deform_attn_core_pytorch and records its arguments.

Fixtures (all seeded, all small):
  core_cases.npz     operator-boundary cases: inputs + reference fp32 and fp64 outputs
  module_cases.npz   MSDeformAttn.forward cases (ref dim 2 and 4, padding mask): state dict, inputs, output
  network_setC.npz   (value, loc, attn, out) captured from the reference DeformableTransformer at its own
                     default initialisation: encoder layer 0 and decoder layer 0 ("Set C": samples sit
                     on pixel centres, the worst case for index rounding)
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("GOM_REFERENCE", "/root/reference")
ADET = os.path.join(REF, "third_party", "adet")
HERE = os.path.dirname(os.path.abspath(__file__))

CAPTURE = []


def import_reference():
    def ns(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m

    ns("adet", ADET)
    ns("adet.layers", os.path.join(ADET, "layers"))
    ns("adet.utils", os.path.join(ADET, "utils"))
    ns("adet.modeling", os.path.join(ADET, "modeling"))
    ns("adet.modeling.model", os.path.join(ADET, "modeling", "model"))
    stub = types.ModuleType("adet._C")
    sys.modules["adet._C"] = stub
    sys.modules["adet"]._C = stub
    msda = importlib.import_module("adet.layers.ms_deform_attn")

    def fwd(value, shapes, lsi, loc, attn, im2col_step):
        out = msda.ms_deform_attn_core_pytorch(value, shapes.tolist(), loc, attn)
        CAPTURE.append(dict(value=value.detach().clone(), shapes=shapes.clone(), lsi=lsi.clone(),
                            loc=loc.detach().clone(), attn=attn.detach().clone(), out=out.detach().clone()))
        return out

    stub.ms_deform_attn_forward = fwd
    dt = importlib.import_module("adet.layers.deformable_transformer")
    return msda, dt


def lsi_of(shapes):
    s = torch.as_tensor(shapes, dtype=torch.long)
    return torch.cat((s.new_zeros((1,)), s.prod(1).cumsum(0)[:-1]))


def core_case(msda, g, shapes, N, M, D, Lq, P, kind):
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    value = torch.randn(N, S, M, D, generator=g)
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    if kind == "uniform":
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.2 - 0.1
    elif kind == "edges":
        # exactly on pixel centres / pixel borders / one pixel outside, per level
        loc = torch.empty(N, Lq, M, L, P, 2)
        for l, (h, w) in enumerate(shapes):
            kx = torch.randint(-2, 2 * w + 3, (N, Lq, M, P), generator=g).float() * 0.5
            ky = torch.randint(-2, 2 * h + 3, (N, Lq, M, P), generator=g).float() * 0.5
            loc[:, :, :, l, :, 0] = kx / w
            loc[:, :, :, l, :, 1] = ky / h
    elif kind == "wide":
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * 3.0 - 1.0
    else:
        raise ValueError(kind)
    out32 = msda.ms_deform_attn_core_pytorch(value, shapes, loc, attn)
    out64 = msda.ms_deform_attn_core_pytorch(value.double(), shapes, loc.double(), attn.double())
    return dict(value=value, loc=loc, attn=attn, shapes=torch.as_tensor(shapes), lsi=lsi_of(shapes),
                out_f32=out32, out_f64=out64)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    msda, dt = import_reference()
    g = torch.Generator().manual_seed(1234)

    # ---------------- operator-boundary cases ----------------
    cases = {
        "uniform_d32": core_case(msda, g, [(8, 12), (4, 6), (2, 3), (1, 2)], 2, 8, 32, 23, 4, "uniform"),
        "edges_d32": core_case(msda, g, [(8, 12), (4, 6), (2, 3), (1, 2)], 1, 8, 32, 40, 4, "edges"),
        "wide_d32": core_case(msda, g, [(5, 7), (3, 4)], 3, 4, 32, 19, 2, "wide"),
        "nonpow2_d12": core_case(msda, g, [(6, 5), (3, 3), (2, 1)], 2, 3, 12, 11, 3, "uniform"),
        "single_d8": core_case(msda, g, [(9, 9)], 1, 2, 8, 5, 1, "uniform"),
        "d64_p8": core_case(msda, g, [(7, 6), (4, 3)], 1, 2, 64, 9, 8, "uniform"),
    }
    flat = {}
    for name, c in cases.items():
        for k, v in c.items():
            flat[f"{name}/{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "core_cases.npz"), **flat)

    # ---------------- module-level cases ----------------
    flat = {}
    for name, ref_dim, with_mask, d_model, heads, levels, points, shapes, N, Lq in [
        ("ref2_mask", 2, True, 256, 8, 4, 4, [(8, 12), (4, 6), (2, 3), (1, 2)], 2, 21),
        ("ref4_nomask", 4, False, 128, 4, 4, 4, [(8, 12), (4, 6), (2, 3), (1, 2)], 1, 33),
        ("small_heads", 2, False, 64, 4, 2, 2, [(6, 5), (3, 3)], 2, 7),
    ]:
        mod = msda.MSDeformAttn(d_model, levels, heads, points)
        # non-trivial weights everywhere (default init zeroes the offset/attention weights)
        with torch.no_grad():
            mod.sampling_offsets.weight.copy_(torch.randn(mod.sampling_offsets.weight.shape, generator=g) * 0.05)
            mod.attention_weights.weight.copy_(torch.randn(mod.attention_weights.weight.shape, generator=g) * 0.2)
            mod.attention_weights.bias.copy_(torch.randn(mod.attention_weights.bias.shape, generator=g) * 0.2)
            mod.value_proj.bias.copy_(torch.randn(mod.value_proj.bias.shape, generator=g) * 0.1)
            mod.output_proj.bias.copy_(torch.randn(mod.output_proj.bias.shape, generator=g) * 0.1)
        S = sum(h * w for h, w in shapes)
        query = torch.randn(N, Lq, d_model, generator=g)
        src = torch.randn(N, S, d_model, generator=g)
        ref = torch.rand(N, Lq, levels, ref_dim, generator=g)
        if ref_dim == 4:
            ref[..., 2:] = ref[..., 2:] * 0.3 + 0.05
        mask = (torch.rand(N, S, generator=g) < 0.15) if with_mask else None
        sh = torch.as_tensor(shapes, dtype=torch.long)
        CAPTURE.clear()
        with torch.no_grad():
            out = mod(query, ref, src, sh, lsi_of(shapes), mask)
        cap = CAPTURE[-1]
        for k, v in mod.state_dict().items():
            flat[f"{name}/sd/{k}"] = v.numpy()
        flat[f"{name}/query"] = query.numpy()
        flat[f"{name}/src"] = src.numpy()
        flat[f"{name}/ref"] = ref.numpy()
        flat[f"{name}/mask"] = mask.numpy() if mask is not None else np.zeros((0,), dtype=bool)
        flat[f"{name}/shapes"] = sh.numpy()
        flat[f"{name}/lsi"] = lsi_of(shapes).numpy()
        flat[f"{name}/cfg"] = np.asarray([d_model, levels, heads, points], dtype=np.int64)
        flat[f"{name}/out"] = out.numpy()
        flat[f"{name}/loc"] = cap["loc"].numpy()
        flat[f"{name}/attn"] = cap["attn"].numpy()
        flat[f"{name}/core_out"] = cap["out"].numpy()
    np.savez_compressed(os.path.join(HERE, "module_cases.npz"), **flat)

    # ---------------- Set C: default-init network capture ----------------
    torch.manual_seed(7)
    shapes = [(9, 14), (5, 7), (3, 4), (2, 2)]
    nq, npts = 4, 25
    tr = dt.DeformableTransformer(d_model=256, nhead=8, num_encoder_layers=1, num_decoder_layers=1,
                                  dim_feedforward=64, dropout=0.0, num_feature_levels=4, dec_n_points=4,
                                  enc_n_points=4, num_proposals=nq, num_points=npts).eval()
    bs = 1
    srcs = [torch.randn(bs, 256, h, w, generator=g) for h, w in shapes]
    masks = [torch.zeros(bs, h, w, dtype=torch.bool) for h, w in shapes]
    pos = [torch.randn(bs, 256, h, w, generator=g) * 0.1 for h, w in shapes]
    src_flat = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
    pos_flat = torch.cat([p.flatten(2).transpose(1, 2) + tr.level_embed[l].view(1, 1, -1) for l, p in enumerate(pos)], 1)
    mask_flat = torch.cat([m.flatten(1) for m in masks], 1)
    sh = torch.as_tensor(shapes, dtype=torch.long)
    lsi = lsi_of(shapes)
    valid_ratios = torch.stack([tr.get_valid_ratio(m) for m in masks], 1)
    CAPTURE.clear()
    with torch.no_grad():
        memory = tr.encoder(src_flat, sh, lsi, valid_ratios, pos_flat, mask_flat)
        bez = torch.rand(bs, nq, 8, generator=g) * 0.8 + 0.1
        refpts = tr.init_points_from_bezier_proposals(bez)            # (bs, nq, npts, 2)
        tgt = torch.randn(bs, nq, npts, 256, generator=g)
        tr.decoder(tgt, refpts, memory, sh, lsi, valid_ratios, query_pos=None, src_padding_mask=mask_flat)
    assert len(CAPTURE) == 2
    flat = {}
    for tag, cap in zip(["enc0", "dec0"], CAPTURE):
        for k, v in cap.items():
            flat[f"{tag}/{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "network_setC.npz"), **flat)
    for f in ["core_cases.npz", "module_cases.npz", "network_setC.npz"]:
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
