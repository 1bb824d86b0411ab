"""Seeded synthetic workloads of the MSDeformAttn hot path, shared by bench.py and tests.

Shapes follow the reference's callers:
  * level pyramid of DeepSolo-R50: strides 8/16/32 by repeated ceil(x/2), 4th level a 3x3 stride-2 conv
    (third_party/adet/modeling/model/detection_transformer_wobackbone.py:82-88)
  * encoder reference points = pixel-centre grid / valid_ratio (third_party/adet/layers/
    deformable_transformer.py:288-300); decoder reference points = 25 points sampled on cubic Bezier
    centre lines of the proposals (:99-106)
  * offsets initialised on a k-pixel compass rose (third_party/adet/layers/ms_deform_attn.py:101-109)

Everything is generated on the CPU with a seeded torch.Generator and then moved, so a (seed, shape) pair
names the same bytes on every machine.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import torch


def level_shapes(height: int, width: int, n_levels: int = 4) -> List[Tuple[int, int]]:
    """(H_l, W_l) of the 4 DeepSolo feature levels for an input of height x width."""
    def half(x):
        return (x + 1) // 2
    h, w = height, width
    for _ in range(3):                      # stride 8 = three halvings
        h, w = half(h), half(w)
    shapes = [(h, w)]
    for _ in range(n_levels - 1):
        h, w = half(h), half(w)
        shapes.append((h, w))
    return shapes


def level_start_index(shapes: Sequence[Tuple[int, int]]) -> torch.Tensor:
    s = torch.as_tensor(shapes, dtype=torch.long)
    return torch.cat((s.new_zeros((1,)), s.prod(1).cumsum(0)[:-1]))


def encoder_reference_points(shapes: Sequence[Tuple[int, int]], n: int = 1) -> torch.Tensor:
    """(n, S, L, 2) pixel-centre grid, valid_ratio 1 -- deformable_transformer.py:288-300."""
    pts = []
    for (h, w) in shapes:
        ys, xs = torch.meshgrid(torch.linspace(0.5, h - 0.5, h), torch.linspace(0.5, w - 0.5, w), indexing="ij")
        pts.append(torch.stack((xs.reshape(-1) / w, ys.reshape(-1) / h), -1))
    ref = torch.cat(pts, 0)                                 # (S, 2)
    return ref[None, :, None, :].expand(n, -1, len(shapes), -1).contiguous()


def decoder_reference_points(g: torch.Generator, n: int, n_proposals: int, n_points: int, n_levels: int) -> torch.Tensor:
    """(n, n_proposals*n_points, L, 2): points on random cubic Bezier centre lines (deformable_transformer.py:99-106)."""
    ctrl = torch.rand(n, n_proposals, 4, 2, generator=g) * 0.2
    ctrl = ctrl + torch.rand(n, n_proposals, 1, 2, generator=g) * 0.8          # a short curve somewhere in the frame
    t = torch.linspace(0, 1, n_points)
    bern = torch.stack(((1 - t) ** 3, 3 * t * (1 - t) ** 2, 3 * t ** 2 * (1 - t), t ** 3), -1)   # (n_points, 4)
    pts = torch.einsum("pk,nqkc->nqpc", bern, ctrl).clamp(0, 1)               # (n, q, p, 2)
    return pts.reshape(n, n_proposals * n_points, 1, 2).expand(-1, -1, n_levels, -1).contiguous()


def compass_offsets(n_heads: int, n_levels: int, n_points: int) -> torch.Tensor:
    """(M, L, P, 2) default sampling_offsets bias in pixels -- ms_deform_attn.py:101-109."""
    thetas = torch.arange(n_heads, dtype=torch.float32) * (2.0 * math.pi / n_heads)
    grid = torch.stack([thetas.cos(), thetas.sin()], -1)
    grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(n_heads, 1, 1, 2).repeat(1, n_levels, n_points, 1)
    for i in range(n_points):
        grid[:, :, i, :] *= i + 1
    return grid


@dataclass
class Workload:
    """Operator-boundary tensors of one MSDeformAttn call (CPU, fp32)."""
    name: str
    shapes: torch.Tensor          # (L,2) int64
    lsi: torch.Tensor             # (L,)  int64
    value: torch.Tensor           # (N,S,M,D)
    ref: torch.Tensor             # (N,Lq,L,2)
    offsets: torch.Tensor         # (N,Lq,M,L,P,2)  pixels (pre-normalisation)
    logits: torch.Tensor          # (N,Lq,M,L*P)
    loc: torch.Tensor             # (N,Lq,M,L,P,2) = ref + offsets/(W,H)
    attn: torch.Tensor            # (N,Lq,M,L,P)   = softmax(logits)

    @property
    def dims(self):
        N, S, M, D = self.value.shape
        _, Lq, _, L, P, _ = self.loc.shape
        return N, S, M, D, L, Lq, P

    def algorithmic_bytes(self, value_bytes: int = 4, out_bytes: int = 4) -> int:
        """SURVEY.md s8(d) B_alg: every tensor of the operator boundary touched exactly once."""
        N, S, M, D, L, Lq, P = self.dims
        v = min(N * S * M * D, 4 * N * Lq * M * L * P * D) * value_bytes
        return v + 8 * N * Lq * M * L * P + 4 * N * Lq * M * L * P + out_bytes * N * Lq * M * D

    def gather_bytes(self, value_bytes: int = 4) -> int:
        N, S, M, D, L, Lq, P = self.dims
        return 4 * N * Lq * M * L * P * D * value_bytes


def make_workload(kind: str, height: int = 720, width: int = 1280, n: int = 1, seed: int = 0, dist: str = "local",
                  n_heads: int = 8, d_head: int = 32, n_levels: int = 4, n_points: int = 4, n_proposals: int = 100,
                  n_ctrl_points: int = 25, sigma_px: float = 2.0) -> Workload:
    """kind: 'encoder' (Lq = S) | 'decoder' (Lq = n_proposals * n_ctrl_points).
    dist: 'local' (compass rose + N(0, sigma_px) pixels, Distribution A) | 'uniform' (loc ~ U(-0.1, 1.1), B)."""
    g = torch.Generator().manual_seed(seed)
    shapes_l = level_shapes(height, width, n_levels)
    shapes = torch.as_tensor(shapes_l, dtype=torch.long)
    lsi = level_start_index(shapes_l)
    S = int(shapes.prod(1).sum())
    M, D, L, P = n_heads, d_head, n_levels, n_points
    value = torch.randn(n, S, M, D, generator=g)
    if kind == "encoder":
        ref = encoder_reference_points(shapes_l, n)
    elif kind == "decoder":
        ref = decoder_reference_points(g, n, n_proposals, n_ctrl_points, L)
    else:
        raise ValueError(kind)
    Lq = ref.shape[1]
    logits = torch.randn(n, Lq, M, L * P, generator=g)
    wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).float()                # (L,2) = (W,H)
    if dist == "local":
        offsets = compass_offsets(M, L, P)[None, None] + torch.randn(n, Lq, M, L, P, 2, generator=g) * sigma_px
    elif dist == "uniform":
        target = torch.rand(n, Lq, M, L, P, 2, generator=g) * 1.2 - 0.1
        offsets = (target - ref[:, :, None, :, None, :]) * wh[None, None, None, :, None, :]
    else:
        raise ValueError(dist)
    offsets = offsets.contiguous()
    loc = ref[:, :, None, :, None, :] + offsets / wh[None, None, None, :, None, :]   # ms_deform_attn.py:143-144
    attn = torch.softmax(logits, -1).view(n, Lq, M, L, P)
    return Workload("%s_%dx%d_%s" % (kind, width, height, dist), shapes, lsi, value, ref.contiguous(), offsets,
                    logits.contiguous(), loc.contiguous(), attn.contiguous())
