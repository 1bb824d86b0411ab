"""Golden gradients of the MSDeformAttn core, produced by the REFERENCE's own CPU path
(third_party/adet/layers/ms_deform_attn.py:40-60 ms_deform_attn_core_pytorch) under torch autograd in
float64.  The reference's CUDA backward (ms_deform_im2col_cuda.cuh:301-920) has no CPU implementation and no
tests; upstream Deformable-DETR validates it exactly this way (gradcheck against the PyTorch path).

Run in the build container only:  python tests/golden/make_golden_backward.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, lsi_of  # noqa: E402


def main():
    torch.manual_seed(0)
    msda, _ = import_reference()
    g = torch.Generator().manual_seed(4321)
    flat = {}
    for name, shapes, N, M, D, Lq, P in [
        ("d32", [(8, 12), (4, 6), (2, 3), (1, 2)], 2, 8, 32, 19, 4),
        ("d64", [(7, 6), (4, 3)], 1, 2, 64, 9, 8),
        ("d12_generic", [(6, 5), (3, 3), (2, 1)], 2, 3, 12, 11, 3),
    ]:
        L = len(shapes)
        S = sum(h * w for h, w in shapes)
        value = torch.randn(N, S, M, D, generator=g, dtype=torch.float64, requires_grad=True)
        loc = (torch.rand(N, Lq, M, L, P, 2, generator=g, dtype=torch.float64) * 1.2 - 0.1).requires_grad_(True)
        attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g, dtype=torch.float64), -1).view(
            N, Lq, M, L, P).detach().requires_grad_(True)
        grad_out = torch.randn(N, Lq, M * D, generator=g, dtype=torch.float64)
        out = msda.ms_deform_attn_core_pytorch(value, shapes, loc, attn)
        out.backward(grad_out)
        for k, v in dict(value=value, loc=loc, attn=attn, grad_out=grad_out, out=out, grad_value=value.grad,
                         grad_loc=loc.grad, grad_attn=attn.grad).items():
            flat[f"{name}/{k}"] = v.detach().numpy().astype(np.float32 if k in ("value", "loc", "attn", "grad_out") else np.float64)
        flat[f"{name}/shapes"] = np.asarray(shapes, dtype=np.int64)
        flat[f"{name}/lsi"] = lsi_of(shapes).numpy()
    np.savez_compressed(os.path.join(HERE, "backward_cases.npz"), **flat)
    print("backward_cases.npz", os.path.getsize(os.path.join(HERE, "backward_cases.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
