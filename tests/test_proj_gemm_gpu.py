"""GPU tests of the 3xTF32 tcgen05 projection GEMM (csrc/proj_gemm.cu) that replaces the nn.Linear calls bordering the
sampler (third_party/adet/layers/ms_deform_attn.py:133-153).  Reference: float64 matmul of the same fp32 inputs;
bar: max|y - ref| / max|ref| <= 1e-5 (torch's own fp32 GEMM sits at ~5e-7, a single-pass TF32 GEMM at ~1e-3)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, w, b, rz):
    y = x.double() @ w.double().t()
    if b is not None:
        y = y + b.double()
    if rz is not None:
        y = y.masked_fill(rz.reshape(-1, 1).bool(), 0.0)
    return y


@pytest.mark.parametrize("M,N,K", [(1, 32, 16), (128, 32, 16), (127, 256, 256), (129, 256, 256), (300, 256, 256),
                                   (2500, 384, 256), (19160, 256, 256), (19160, 384, 256), (19160, 128, 256),
                                   (80000, 256, 256), (80000, 384, 256), (4096, 512, 64)])
def test_linear_matches_float64(M, N, K):
    from gomatching_b200.projections import linear_3xtf32
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    x = torch.randn(M, K, generator=g, device="cuda") * 3.0
    w = torch.randn(N, K, generator=g, device="cuda") / K ** 0.5
    b = torch.randn(N, generator=g, device="cuda")
    y = linear_3xtf32(x, w, b)
    ref = _ref(x, w, b, None)
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    assert err <= 1e-5, err


@pytest.mark.parametrize("env", [{"MSDA_GEMM_PERSISTENT": "0"}, {"MSDA_GEMM_BN": "32"}, {"MSDA_GEMM_BN": "64"},
                                 {"MSDA_GEMM_BN": "128"}, {"MSDA_GEMM_BN": "256"}])
@pytest.mark.parametrize("M,N,K,relu", [(19160, 256, 256, False), (40000, 1024, 256, True), (2500, 256, 1024, False)])
def test_every_tile_shape_and_the_tile_per_cta_kernels(monkeypatch, env, M, N, K, relu):
    """The persistent kernel's tile width is a heuristic (and MSDA_GEMM_PERSISTENT=0 selects the round-1 kernels that the
    A/B numbers in DESIGN.md are measured against): every choice must give the same fp32-grade result, including many
    tiles per CTA (both TMEM accumulator buffers and every ring slot reused many times)."""
    from gomatching_b200.projections import linear_3xtf32
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g, device="cuda") * 2.0
    w = torch.randn(N, K, generator=g, device="cuda") / K ** 0.5
    b = torch.randn(N, generator=g, device="cuda")
    rz = None if relu else torch.rand(M, generator=g, device="cuda") < 0.1
    y = linear_3xtf32(x, w, b, rz, relu=relu)
    ref = _ref(x, w, b, rz)
    if relu:
        ref = ref.clamp(min=0)
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    assert err <= 1e-5, err
    if rz is not None:
        assert bool((y[rz] == 0).all())


def test_row_zero_bias_none_and_pitched_views():
    from gomatching_b200.projections import linear_3xtf32
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 1000, 256, 256
    xw = torch.randn(M, K + 64, generator=g, device="cuda")
    x = xw[:, :K]                                         # row pitch 320 floats
    w = torch.randn(N, K, generator=g, device="cuda") / 16
    rz = torch.rand(M, generator=g, device="cuda") < 0.25
    out_wide = torch.full((M, N + 128), 7.0, device="cuda")
    y = linear_3xtf32(x, w, None, rz, out=out_wide[:, :N])
    ref = _ref(x, w, None, rz)
    assert float((y.double() - ref).abs().max() / ref.abs().max()) <= 1e-5
    assert bool((y[rz] == 0).all())
    assert bool((out_wide[:, N:] == 7.0).all()), "columns outside the output view were touched"


def test_batched_shape_and_weight_cache_invalidation():
    from gomatching_b200.projections import linear_3xtf32
    x = torch.randn(2, 300, 256, device="cuda")
    lin = torch.nn.Linear(256, 256).cuda()
    y = linear_3xtf32(x, lin.weight, lin.bias)
    assert y.shape == (2, 300, 256)
    ref = _ref(x.reshape(-1, 256), lin.weight.detach(), lin.bias.detach(), None).reshape(2, 300, 256)
    assert float((y.double() - ref).abs().max() / ref.abs().max()) <= 1e-5
    with torch.no_grad():
        lin.weight.mul_(2.0)                              # in-place update bumps the version: the split must be redone
    y2 = linear_3xtf32(x, lin.weight, lin.bias)
    ref2 = _ref(x.reshape(-1, 256), lin.weight.detach(), lin.bias.detach(), None).reshape(2, 300, 256)
    assert float((y2.double() - ref2).abs().max() / ref2.abs().max()) <= 1e-5


def test_special_values_and_errors():
    from gomatching_b200.projections import linear_3xtf32
    from gomatching_b200._native import MSDAError
    x = torch.zeros(64, 32, device="cuda")
    x[0, 0] = float("inf")
    x[1, 1] = float("nan")
    x[2, 2] = 3.0e38
    w = torch.eye(32, device="cuda")
    y = linear_3xtf32(x, w)
    assert torch.isinf(y[0, 0]) or torch.isnan(y[0, 0])
    assert torch.isnan(y[1, 1])
    assert bool(torch.isfinite(y[3:]).all()) and float(y[3:].abs().max()) == 0.0
    with pytest.raises(RuntimeError):
        linear_3xtf32(torch.zeros(4, 32), w.cpu())        # CPU tensors: no CPU path
    with pytest.raises(MSDAError):
        linear_3xtf32(torch.zeros(4, 24, device="cuda"), torch.zeros(32, 24, device="cuda"))   # K % 16 != 0


def test_cuda_graph_capture():
    from gomatching_b200.projections import linear_3xtf32
    x = torch.randn(512, 256, device="cuda")
    w = torch.randn(256, 256, device="cuda") / 16
    out = torch.empty(512, 256, device="cuda")
    linear_3xtf32(x, w, out=out)                          # warm-up: weight split cached outside the capture
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        linear_3xtf32(x, w, out=out)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    ref = _ref(x, w, None, None)
    assert float((out.double() - ref).abs().max() / ref.abs().max()) <= 1e-5
