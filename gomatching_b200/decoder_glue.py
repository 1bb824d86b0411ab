"""One-launch versions of the elementwise glue around every layer of DeepSolo's point-query decoder
(csrc/decoder_glue.cu): ``gen_point_pos_embed`` (third_party/adet/modeling/model/utils.py:24-37) and the reference-point
refinement ``(tmp + inverse_sigmoid(ref)).sigmoid()`` (deformable_transformer.py:483-486).  Bit-identical to the eager
code; CUDA fp32 only (no CPU path -- callers keep the reference's functions for anything else)."""
from __future__ import annotations

import torch

from . import _native

_dim_t_cache: "dict[tuple, torch.Tensor]" = {}


def supported(*tensors: torch.Tensor) -> bool:
    return all(t.is_cuda and t.dtype == torch.float32 for t in tensors)


def _dim_t(d_model: int, temp, device) -> torch.Tensor:
    """``temp ** (2 * (arange(d_model / 2) // 2) / (d_model / 2))`` computed by torch itself (utils.py:27-29), once."""
    key = (int(d_model), float(temp), str(device))
    hit = _dim_t_cache.get(key)
    if hit is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("decoder_glue: run one forward before capturing a CUDA graph")
        dim = d_model // 2
        t = torch.arange(dim, dtype=torch.float32, device=device)
        hit = _dim_t_cache[key] = (temp ** (2 * torch.div(t, 2, rounding_mode='trunc') / dim)).contiguous()
    return hit


def point_pos_embed(reference_points: torch.Tensor, valid_ratios: "torch.Tensor | None", d_model: int, temp) -> torch.Tensor:
    """``gen_point_pos_embed((reference_points[:, :, :, None] * valid_ratios[:, None, None])[:, :, :, 0, :], d_model, temp)``.
    reference_points (bs, nq, n_pts, 2); valid_ratios (bs, L, 2) or None; returns (bs, nq, n_pts, d_model)."""
    if not supported(reference_points) or reference_points.shape[-1] != 2:
        raise RuntimeError("point_pos_embed: CUDA fp32 (.., 2) reference points only")
    ref = reference_points.contiguous()
    bs = ref.shape[0]
    points = ref.numel() // 2
    vr = None
    if valid_ratios is not None:
        vr = valid_ratios.contiguous()
        if not supported(vr) or vr.shape[0] != bs or vr.shape[-1] != 2:
            raise RuntimeError("point_pos_embed: valid_ratios must be CUDA fp32 (bs, L, 2)")
    dim_t = _dim_t(d_model, temp, ref.device)
    out = torch.empty(ref.shape[:-1] + (d_model,), dtype=torch.float32, device=ref.device)
    with torch.cuda.device(ref.device):
        _native.check(_native.lib().msda_b200_point_pos_embed_f32(
            ref.data_ptr(), vr.data_ptr() if vr is not None else None, vr.stride(0) if vr is not None else 0,
            dim_t.data_ptr(), points, points // bs, d_model // 2, out.data_ptr(),
            torch.cuda.current_stream(ref.device).cuda_stream), "msda_b200_point_pos_embed_f32")
    return out


def refine_points(tmp: torch.Tensor, reference_points: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """``(tmp + inverse_sigmoid(reference_points)).sigmoid()`` (adet/utils/misc.py:115-119 for the inverse)."""
    if not supported(tmp, reference_points) or tmp.shape != reference_points.shape:
        raise RuntimeError("refine_points: CUDA fp32 tensors of one shape only")
    t, r = tmp.contiguous(), reference_points.contiguous()
    out = torch.empty_like(r)
    with torch.cuda.device(r.device):
        _native.check(_native.lib().msda_b200_refine_points_f32(
            t.data_ptr(), r.data_ptr(), r.numel(), eps, out.data_ptr(),
            torch.cuda.current_stream(r.device).cuda_stream), "msda_b200_refine_points_f32")
    return out
