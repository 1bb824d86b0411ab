"""Golden fixtures for gomatching_b200/encoder_layer.py, produced by the REFERENCE's own encoder layer.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_encoder_layer.py

`DeformableTransformerEncoderLayer` (third_party/adet/layers/deformable_transformer.py:218-278) is imported unmodified
through the same namespace stubs as make_golden.py (its MSDeformAttn core routed to the reference's
ms_deform_attn_core_pytorch) and run on the CPU in eval mode with seeded, non-trivial weights.  Two cases:
with / without padding mask and positional embedding.  Output: tests/golden/encoder_layer_cases.npz
(state dict, inputs, the layer output and the intermediate after the attention block).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    msda, dt = mg.import_reference()
    g = torch.Generator().manual_seed(99)
    flat = {}
    for name, d_model, d_ffn, heads, shapes, N, with_mask, with_pos in [
        ("mask_pos", 256, 512, 8, [(8, 12), (4, 6), (2, 3), (1, 2)], 2, True, True),
        ("plain", 128, 256, 4, [(6, 5), (3, 3), (2, 2), (1, 1)], 1, False, False),
    ]:
        layer = dt.DeformableTransformerEncoderLayer(d_model, d_ffn, 0.1, "relu", len(shapes), heads, 4).eval()
        with torch.no_grad():
            for p in layer.parameters():      # default init zeroes the offset/attention weights: make everything non-trivial
                if p.dim() > 1:
                    p.copy_(torch.randn(p.shape, generator=g) * (0.05 if p.shape[0] != p.shape[1] else 0.08))
                else:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.1 + (1.0 if p.shape[0] == d_model and p.mean() > 0.5 else 0.0))
        S = sum(h * w for h, w in shapes)
        src = torch.randn(N, S, d_model, generator=g)
        pos = torch.randn(N, S, d_model, generator=g) * 0.1 if with_pos else None
        sh = torch.as_tensor(shapes, dtype=torch.long)
        lsi = mg.lsi_of(shapes)
        valid = torch.ones(N, len(shapes), 2)
        ref = dt.DeformableTransformerEncoder.get_reference_points(shapes, valid, "cpu")
        mask = (torch.rand(N, S, generator=g) < 0.1) if with_mask else None
        with torch.no_grad():
            out = layer(src, pos, ref, sh, lsi, mask)
            attn_block = layer.norm1(src + layer.self_attn(layer.with_pos_embed(src, pos), ref, src, sh, lsi, mask))
            ffn_only = layer.forward_ffn(src)
        for k, v in layer.state_dict().items():
            flat[f"{name}/sd/{k}"] = v.numpy()
        flat[f"{name}/cfg"] = np.asarray([d_model, d_ffn, heads, len(shapes), 4], dtype=np.int64)
        flat[f"{name}/src"] = src.numpy()
        flat[f"{name}/pos"] = pos.numpy() if pos is not None else np.zeros((0,), dtype=np.float32)
        flat[f"{name}/ref"] = ref.numpy()
        flat[f"{name}/mask"] = mask.numpy() if mask is not None else np.zeros((0,), dtype=bool)
        flat[f"{name}/shapes"] = sh.numpy()
        flat[f"{name}/lsi"] = lsi.numpy()
        flat[f"{name}/out"] = out.numpy()
        flat[f"{name}/attn_block"] = attn_block.numpy()
        flat[f"{name}/ffn_only"] = ffn_only.numpy()
    path = os.path.join(HERE, "encoder_layer_cases.npz")
    np.savez_compressed(path, **flat)
    print(path, os.path.getsize(path) // 1024, "KiB")

    # ---- decoder layer (deformable_transformer.py:326-427), same recipe ----
    flat = {}
    for name, d_model, d_ffn, heads, shapes, N, nq, npts, with_mask, ref4 in [
        ("dec_mask", 128, 256, 4, [(8, 12), (4, 6), (2, 3), (1, 2)], 2, 5, 7, True, False),
        ("dec_shared_ref", 128, 128, 4, [(6, 5), (3, 3), (2, 2), (1, 1)], 1, 3, 4, False, True),
    ]:
        layer = dt.DeformableCompositeTransformerDecoderLayer(d_model, d_ffn, 0.1, "relu", len(shapes), heads, 4).eval()
        with torch.no_grad():
            for p in layer.parameters():
                if p.dim() > 1:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.06)
                else:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            for nrm in (layer.norm_intra, layer.norm_inter, layer.norm_cross, layer.norm3):
                nrm.weight.add_(1.0)
        S = sum(h * w for h, w in shapes)
        tgt = torch.randn(N, nq, npts, d_model, generator=g)
        qpos = torch.randn(N, nq, npts, d_model, generator=g) * 0.1
        src = torch.randn(N, S, d_model, generator=g)
        sh = torch.as_tensor(shapes, dtype=torch.long)
        lsi = mg.lsi_of(shapes)
        # (bs, n_q, L, 2): one point per proposal, repeated over its points by the layer; else (bs, n_q, n_pts, L, 2)
        ref = torch.rand(N, nq, len(shapes), 2, generator=g) if ref4 else torch.rand(N, nq, npts, len(shapes), 2, generator=g)
        mask = (torch.rand(N, S, generator=g) < 0.1) if with_mask else None
        with torch.no_grad():
            out = layer(tgt, qpos, ref, src, sh, lsi, mask)
        for k, v in layer.state_dict().items():
            flat[f"{name}/sd/{k}"] = v.numpy()
        flat[f"{name}/cfg"] = np.asarray([d_model, d_ffn, heads, len(shapes), 4], dtype=np.int64)
        flat[f"{name}/tgt"] = tgt.numpy()
        flat[f"{name}/qpos"] = qpos.numpy()
        flat[f"{name}/src"] = src.numpy()
        flat[f"{name}/ref"] = ref.numpy()
        flat[f"{name}/mask"] = mask.numpy() if mask is not None else np.zeros((0,), dtype=bool)
        flat[f"{name}/shapes"] = sh.numpy()
        flat[f"{name}/lsi"] = lsi.numpy()
        flat[f"{name}/out"] = out.numpy()
    path = os.path.join(HERE, "decoder_layer_cases.npz")
    np.savez_compressed(path, **flat)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
