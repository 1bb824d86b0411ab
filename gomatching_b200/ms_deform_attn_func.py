"""Operator layer: drop-in for the reference's `adet._C` entry points and `_MSDeformAttnFunction`.

Reference interfaces mirrored (all under /root/reference/third_party/adet/layers/):
  * ``_C.ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
    im2col_step)``  -- csrc/vision.cpp:52-53, csrc/DeformAttn/ms_deform_attn.h:20-39,
    csrc/DeformAttn/ms_deform_attn_cuda.cu:20-80
  * ``_C.ms_deform_attn_backward(..., grad_output, im2col_step)`` -- csrc/vision.cpp:54-55,
    ms_deform_attn_cuda.cu:83-153
  * ``_MSDeformAttnFunction`` -- ms_deform_attn.py:20-37 (same positional arguments, same returned grads)

Differences, none of which change results: the output is allocated uninitialised (the kernels write
every element once; the reference memsets, ms_deform_attn_cuda.cu:54), all N batch items go in one launch
(``im2col_step`` only keeps its divisibility check), bf16 value is accepted (the reference dispatches
fp32/fp64 only, ms_deform_attn_cuda.cu:64).  CPU tensors raise, like the reference
(ms_deform_attn.h:38 "Not implemented on the CPU") -- there is no CPU or PyTorch fallback here.
"""
from __future__ import annotations

import torch
from torch.autograd.function import once_differentiable

from . import _native


def _require(cond: bool, msg: str) -> None:
    if not cond:
        raise RuntimeError(msg)


def _check_common(value, spatial_shapes, level_start_index, sampling_loc, attn_weight):
    # ms_deform_attn.h:28-38 / ms_deform_attn_cuda.cu:28-38
    _require(value.is_cuda, "Not implemented on the CPU")
    _require(value.is_contiguous(), "value tensor has to be contiguous")
    _require(spatial_shapes.is_contiguous(), "spatial_shapes tensor has to be contiguous")
    _require(level_start_index.is_contiguous(), "level_start_index tensor has to be contiguous")
    _require(sampling_loc.is_contiguous(), "sampling_loc tensor has to be contiguous")
    _require(attn_weight.is_contiguous(), "attn_weight tensor has to be contiguous")
    _require(spatial_shapes.is_cuda, "spatial_shapes must be a CUDA tensor")
    _require(level_start_index.is_cuda, "level_start_index must be a CUDA tensor")
    _require(sampling_loc.is_cuda, "sampling_loc must be a CUDA tensor")
    _require(attn_weight.is_cuda, "attn_weight must be a CUDA tensor")
    _require(spatial_shapes.dtype == torch.int64 and level_start_index.dtype == torch.int64,
             "spatial_shapes / level_start_index must be int64")
    _require(value.dim() == 4 and sampling_loc.dim() == 6 and attn_weight.dim() == 5, "bad tensor ranks")
    N, S, M, D = value.shape
    N2, Lq, M2, L, P, two = sampling_loc.shape
    _require((N2, M2, two) == (N, M, 2), "sampling_loc shape does not match value")
    _require(tuple(attn_weight.shape) == (N, Lq, M, L, P), "attn_weight shape does not match sampling_loc")
    _require(spatial_shapes.shape == (L, 2) and level_start_index.shape == (L,), "bad spatial_shapes / level_start_index")
    return N, S, M, D, L, Lq, P


def _check_im2col_step(N: int, im2col_step: int) -> None:
    step = min(N, int(im2col_step))                        # ms_deform_attn_cuda.cu:50
    _require(step > 0 and N % step == 0, "batch(%d) must divide im2col_step(%d)" % (N, step))   # :52


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step=64,
                           tuning=None, spatial_shapes_list=None):
    """Same call as ``adet._C.ms_deform_attn_forward``; returns (N, Lq, M*D).

    ``spatial_shapes_list`` (optional, [(H_l, W_l), ...] Python ints -- the list the caller built the tensor from,
    deformable_transformer.py:157-169) saves the one device->host read per new pyramid that the TMA window kernel's
    tensor maps otherwise cost; results never depend on it (the kernels validate it against the device tensor)."""
    N, S, M, D, L, Lq, P = _check_common(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    _check_im2col_step(N, im2col_step)
    lib = _native.lib()
    if value.dtype == torch.float64:        # the reference dispatches double too (ms_deform_attn_cuda.cu:64)
        loc64 = sampling_loc.double().contiguous()
        attn64 = attn_weight.double().contiguous()
        out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        with torch.cuda.device(value.device):
            rc = lib.msda_b200_forward_f64(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                           loc64.data_ptr(), attn64.data_ptr(), N, S, M, D, L, Lq, P, out.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream)
        _native.check(rc, "ms_deform_attn_forward")
        return out
    if value.dtype == torch.float32:
        fn = lib.msda_b200_forward_f32_ex
    elif value.dtype == torch.bfloat16:
        fn = lib.msda_b200_forward_bf16_ex
    else:
        raise RuntimeError("msda_b200: value dtype %s has no kernel (float32, float64, bfloat16 are implemented)" % value.dtype)
    loc = sampling_loc if sampling_loc.dtype == torch.float32 else sampling_loc.float()
    attn = attn_weight if attn_weight.dtype == torch.float32 else attn_weight.float()
    out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device):
        tuning = _native.auto_window_tuning(value, spatial_shapes, level_start_index, Lq, tuning, spatial_shapes_list)
        stream = torch.cuda.current_stream().cuda_stream
        rc = fn(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), loc.data_ptr(),
                attn.data_ptr(), N, S, M, D, L, Lq, P, out.data_ptr(), stream, _native.make_tuning(tuning))
    _native.check(rc, "ms_deform_attn_forward")
    return out


def _row_pitch(t, row_len):
    """Floats between consecutive (b,q) rows if ``t`` is a dense-row view (e.g. a column slice of a wider projection
    output), else None."""
    flat_ok = t.stride(-1) == 1
    inner = 1
    for d in range(t.dim() - 1, 1, -1):                    # dims after (N, Lq) must be dense
        flat_ok = flat_ok and (t.shape[d] == 1 or t.stride(d) == inner)
        inner *= t.shape[d]
    if not flat_ok or inner != row_len:
        return None
    pitch = t.stride(1) if t.shape[1] > 1 else row_len
    if pitch < row_len or (t.shape[0] > 1 and t.stride(0) != pitch * t.shape[1]):
        return None
    return pitch


def ms_deform_attn_forward_fused(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                                 attention_logits, n_points=None, tuning=None, spatial_shapes_list=None):
    """softmax(logits) + offsets->locations + sampling + weighted reduction in one kernel.

    value (N,S,M,D) fp32|bf16; reference_points (N,Lq,L,2|4) fp32; sampling_offsets (N,Lq,M,L,P,2) fp32 (raw
    projection output); attention_logits (N,Lq,M,L*P) fp32 (pre-softmax).  Equivalent to
    ms_deform_attn.py:137-152 without materialising sampling_locations / attention_weights.
    """
    _require(value.is_cuda, "Not implemented on the CPU")
    for t, name in ((value, "value"), (reference_points, "reference_points"), (spatial_shapes, "spatial_shapes"),
                    (level_start_index, "level_start_index")):
        _require(t.is_cuda and t.is_contiguous(), "%s tensor has to be a contiguous CUDA tensor" % name)
    _require(sampling_offsets.is_cuda and attention_logits.is_cuda, "offsets / logits must be CUDA tensors")
    N, S, M, D = value.shape
    N2, Lq, M2, L, P, two = sampling_offsets.shape
    attention_logits = attention_logits.reshape(N, Lq, M, L * P) if attention_logits.dim() != 4 else attention_logits
    # dense, or row-pitched views of one merged projection output (no copy); anything else is made contiguous
    off_pitch = _row_pitch(sampling_offsets, M * L * P * 2)
    lg_pitch = _row_pitch(attention_logits, M * L * P)
    pitched = (off_pitch is not None and lg_pitch is not None and (off_pitch != M * L * P * 2 or lg_pitch != M * L * P)
               and D == 32 and L == 4 and P == 4 and off_pitch % 2 == 0 and sampling_offsets.data_ptr() % 8 == 0)
    if not pitched:
        sampling_offsets = sampling_offsets.contiguous()
        attention_logits = attention_logits.contiguous()
    _require((N2, M2, two) == (N, M, 2), "sampling_offsets shape does not match value")
    ref_dim = reference_points.shape[-1]
    if ref_dim not in (2, 4):
        raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(ref_dim))
    _require(tuple(reference_points.shape) == (N, Lq, L, ref_dim), "bad reference_points shape")
    _require(attention_logits.numel() == N * Lq * M * L * P, "bad attention_logits shape")
    _require(reference_points.dtype == torch.float32 and sampling_offsets.dtype == torch.float32
             and attention_logits.dtype == torch.float32, "reference_points / offsets / logits must be float32")
    lib = _native.lib()
    if value.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("msda_b200: value dtype %s has no kernel" % value.dtype)
    f32 = value.dtype == torch.float32
    out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device):
        tuning = _native.auto_window_tuning(value, spatial_shapes, level_start_index, Lq, tuning, spatial_shapes_list)
        stream = torch.cuda.current_stream().cuda_stream
        log = _native.event_log
        if log is not None:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        if pitched:
            fn = lib.msda_b200_forward_fused_pitched_f32 if f32 else lib.msda_b200_forward_fused_pitched_bf16
            rc = fn(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                    reference_points.data_ptr(), ref_dim, sampling_offsets.data_ptr(), off_pitch,
                    attention_logits.data_ptr(), lg_pitch, N, S, M, D, L, Lq, P, out.data_ptr(), stream,
                    _native.make_tuning(tuning))
        else:
            fn = lib.msda_b200_forward_fused_f32 if f32 else lib.msda_b200_forward_fused_bf16
            rc = fn(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                    reference_points.data_ptr(), ref_dim, sampling_offsets.data_ptr(), attention_logits.data_ptr(), N, S,
                    M, D, L, Lq, P, out.data_ptr(), stream, _native.make_tuning(tuning))
        if log is not None:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            log.append((("fused", N, S, Lq, str(value.dtype)), ev0, ev1))
    _native.check(rc, "ms_deform_attn_forward_fused")
    return out


def pair_value_bf16(value, spatial_shapes, level_start_index):
    """value (N, S, M, D=32) float32 | bfloat16  ->  the neighbour-paired bf16 layout (N, M, S, 2, D) of the bf16 operator
    mode (include/msda_b200.h): pair p = {pixel p, its right-hand neighbour in the same image row (zeros at the row end)}."""
    _require(value.is_cuda, "Not implemented on the CPU")
    _require(value.dim() == 4 and value.is_contiguous(), "value must be a contiguous (N, S, M, D) tensor")
    _require(value.dtype in (torch.float32, torch.bfloat16), "pair_value_bf16 takes float32 or bfloat16 value")
    _require(spatial_shapes.is_cuda and level_start_index.is_cuda and spatial_shapes.dtype == torch.int64
             and level_start_index.dtype == torch.int64, "spatial_shapes / level_start_index must be CUDA int64 tensors")
    N, S, M, D = value.shape
    L = int(spatial_shapes.shape[0])
    paired = torch.empty((N, M, S, 2, D), dtype=torch.bfloat16, device=value.device)
    with torch.cuda.device(value.device):
        rc = _native.lib().msda_b200_pair_value_bf16(value.data_ptr(), int(value.dtype == torch.bfloat16),
                                                     spatial_shapes.contiguous().data_ptr(),
                                                     level_start_index.contiguous().data_ptr(), N, S, M, D, L, paired.data_ptr(),
                                                     torch.cuda.current_stream().cuda_stream)
    _native.check(rc, "pair_value_bf16")
    return paired


def _check_paired(paired, spatial_shapes, level_start_index):
    _require(paired.is_cuda, "Not implemented on the CPU")
    _require(paired.dim() == 5 and paired.shape[3] == 2 and paired.dtype == torch.bfloat16 and paired.is_contiguous(),
             "paired value must be a contiguous bfloat16 (N, M, S, 2, D) tensor (pair_value_bf16)")
    _require(spatial_shapes.is_cuda and level_start_index.is_cuda and spatial_shapes.is_contiguous()
             and level_start_index.is_contiguous() and spatial_shapes.dtype == torch.int64
             and level_start_index.dtype == torch.int64, "spatial_shapes / level_start_index must be contiguous CUDA int64")
    N, M, S, _, D = paired.shape
    return N, S, M, D, int(spatial_shapes.shape[0])


def ms_deform_attn_forward_paired(paired, spatial_shapes, level_start_index, sampling_loc, attn_weight):
    """Core operator on the paired bf16 layout: same arguments as ``ms_deform_attn_forward`` with ``value`` replaced by
    ``pair_value_bf16(value, ...)``.  Returns (N, Lq, M*D) bfloat16; 2e-2 bar vs the fp32 reference."""
    N, S, M, D, L = _check_paired(paired, spatial_shapes, level_start_index)
    _require(sampling_loc.is_cuda and attn_weight.is_cuda and sampling_loc.is_contiguous() and attn_weight.is_contiguous(),
             "sampling_loc / attn_weight must be contiguous CUDA tensors")
    _require(sampling_loc.dim() == 6 and tuple(sampling_loc.shape[:1] + sampling_loc.shape[2:4] + sampling_loc.shape[5:]) == (N, M, L, 2),
             "sampling_loc shape does not match value")
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    _require(tuple(attn_weight.shape) == (N, Lq, M, L, P), "attn_weight shape does not match sampling_loc")
    loc = sampling_loc if sampling_loc.dtype == torch.float32 else sampling_loc.float()
    attn = attn_weight if attn_weight.dtype == torch.float32 else attn_weight.float()
    out = torch.empty((N, Lq, M * D), dtype=torch.bfloat16, device=paired.device)
    with torch.cuda.device(paired.device):
        rc = _native.lib().msda_b200_forward_paired_bf16(paired.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                                         loc.data_ptr(), attn.data_ptr(), N, S, M, D, L, Lq, P, out.data_ptr(),
                                                         torch.cuda.current_stream().cuda_stream)
    _native.check(rc, "ms_deform_attn_forward_paired")
    return out


def ms_deform_attn_forward_fused_paired(paired, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                                        attention_logits):
    """``ms_deform_attn_forward_fused`` on the paired bf16 layout."""
    N, S, M, D, L = _check_paired(paired, spatial_shapes, level_start_index)
    _require(sampling_offsets.dim() == 6 and sampling_offsets.shape[0] == N and sampling_offsets.shape[2] == M
             and sampling_offsets.shape[3] == L and sampling_offsets.shape[5] == 2, "sampling_offsets shape does not match value")
    Lq, P = sampling_offsets.shape[1], sampling_offsets.shape[4]
    ref_dim = reference_points.shape[-1]
    if ref_dim not in (2, 4):
        raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(ref_dim))
    _require(tuple(reference_points.shape) == (N, Lq, L, ref_dim), "bad reference_points shape")
    _require(attention_logits.numel() == N * Lq * M * L * P, "bad attention_logits shape")
    ref = reference_points.float().contiguous()
    off = sampling_offsets.float().contiguous()
    lg = attention_logits.float().contiguous()
    _require(ref.is_cuda and off.is_cuda and lg.is_cuda, "reference_points / offsets / logits must be CUDA tensors")
    out = torch.empty((N, Lq, M * D), dtype=torch.bfloat16, device=paired.device)
    with torch.cuda.device(paired.device):
        log = _native.event_log
        if log is not None:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        rc = _native.lib().msda_b200_forward_fused_paired_bf16(
            paired.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), ref.data_ptr(), ref_dim, off.data_ptr(),
            lg.data_ptr(), N, S, M, D, L, Lq, P, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        if log is not None:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            log.append((("fused_paired", N, S, Lq, "bf16_paired"), ev0, ev1))
    _native.check(rc, "ms_deform_attn_forward_fused_paired")
    return out


def fused_supported(value_dtype, D: int, L: int, P: int) -> bool:
    """Shapes the fused kernel is instantiated for (csrc/msda_forward.cu tiled_supported)."""
    if value_dtype not in (torch.float32, torch.bfloat16) or L > 16:
        return False
    if D == 32:
        return L * P in (8, 16, 32) and (L * P) % (D * value_dtype.itemsize // 16) == 0
    if D == 64:
        return L * P == 16
    return False


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step=64):
    """Same call as ``adet._C.ms_deform_attn_backward``; returns [grad_value, grad_sampling_loc, grad_attn_weight]."""
    N, S, M, D, L, Lq, P = _check_common(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    _require(grad_output.is_cuda, "grad_output must be a CUDA tensor")
    _check_im2col_step(N, im2col_step)
    _require(value.dtype == torch.float32, "msda_b200 backward is implemented for float32")
    # the kernel reads fp32 everywhere: other dtypes are converted on the way in and the grads converted back
    loc_dtype, attn_dtype = sampling_loc.dtype, attn_weight.dtype
    sampling_loc = sampling_loc.float().contiguous()
    attn_weight = attn_weight.float().contiguous()
    grad_output = grad_output.float().contiguous()
    _require(tuple(grad_output.shape) == (N, Lq, M * D), "grad_output must be (N, Lq, M*D)")
    grad_value = torch.zeros_like(value)                   # accumulated with atomics, like ms_deform_attn_cuda.cu:118
    grad_loc = torch.empty_like(sampling_loc)
    grad_attn = torch.empty_like(attn_weight)
    with torch.cuda.device(value.device):
        stream = torch.cuda.current_stream().cuda_stream
        rc = _native.lib().msda_b200_backward_f32(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
            attn_weight.data_ptr(), grad_output.data_ptr(), N, S, M, D, L, Lq, P, grad_value.data_ptr(),
            grad_loc.data_ptr(), grad_attn.data_ptr(), stream)
    _native.check(rc, "ms_deform_attn_backward")
    return [grad_value, grad_loc.to(loc_dtype), grad_attn.to(attn_dtype)]


class MSDeformAttnFunction(torch.autograd.Function):
    """ms_deform_attn.py:20-37, same positional arguments."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        ctx.im2col_step = im2col_step
        output = ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                        attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        grad_value, grad_loc, grad_attn = ms_deform_attn_backward(value, shapes, lsi, loc, attn, grad_output,
                                                                  ctx.im2col_step)
        return grad_value, None, None, grad_loc, grad_attn, None


_MSDeformAttnFunction = MSDeformAttnFunction   # the reference's private name (ms_deform_attn.py:20)


def sample_index(sampling_loc, spatial_shapes, level_start_index, n_heads: int, d_head: int):
    """Dump the sampling indices the kernels derive (same device function): int64 tensor (N,Lq,M,L,P,6) viewed as
    [h_low, w_low, in_range, corner_mask, level_offset_lo, level_offset_hi] int32 -> returned as a dict of tensors."""
    _require(sampling_loc.is_cuda and sampling_loc.is_contiguous() and sampling_loc.dtype == torch.float32,
             "sampling_loc must be a contiguous float32 CUDA tensor")
    N, Lq, M, L, P, _ = sampling_loc.shape
    _require(M == n_heads, "n_heads mismatch")
    rec = torch.empty((N, Lq, M, L, P, 3), dtype=torch.int64, device=sampling_loc.device)   # 24-byte records
    with torch.cuda.device(sampling_loc.device):
        stream = torch.cuda.current_stream().cuda_stream
        rc = _native.lib().msda_b200_sample_index_f32(sampling_loc.data_ptr(), spatial_shapes.data_ptr(),
                                                      level_start_index.data_ptr(), N, Lq, M, d_head, L, P,
                                                      rec.data_ptr(), stream)
    _native.check(rc, "sample_index")
    i32 = rec.view(torch.int32)                            # (..., 6)
    return {"h_low": i32[..., 0], "w_low": i32[..., 1], "in_range": i32[..., 2], "corner_mask": i32[..., 3],
            "level_offset": rec[..., 2]}


def locations_softmax(spatial_shapes, reference_points, sampling_offsets, attention_logits, lanes_per_unit: int = 8):
    """The fused kernel's glue arithmetic on its own: returns (sampling_locations, attention_weights)."""
    N, Lq, M, L, P, _ = sampling_offsets.shape
    ref_dim = reference_points.shape[-1]
    if ref_dim not in (2, 4):
        raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(ref_dim))
    loc = torch.empty_like(sampling_offsets)
    attn = torch.empty((N, Lq, M, L, P), dtype=torch.float32, device=sampling_offsets.device)
    with torch.cuda.device(sampling_offsets.device):
        stream = torch.cuda.current_stream().cuda_stream
        rc = _native.lib().msda_b200_locations_softmax_f32(
            spatial_shapes.data_ptr(), reference_points.data_ptr(), ref_dim, sampling_offsets.data_ptr(),
            attention_logits.data_ptr(), N, M, L, Lq, P, lanes_per_unit, loc.data_ptr(), attn.data_ptr(), stream)
    _native.check(rc, "locations_softmax")
    return loc, attn
