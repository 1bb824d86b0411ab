"""The bf16 operator mode on the neighbour-paired value layout (include/msda_b200.h, csrc/msda_forward_paired.cu).

Bars: BASELINE.json north_star -- bf16 outputs within 2e-2 (max|a-b| / max|b|) of the fp32 reference on the same
inputs; sampling indices the same bits as every other kernel (shared phase-1 code; checked here through edge-position
inputs whose result would move visibly if a corner or a zero-padding case were wrong).  Beyond the bar: the kernel must
equal a float64 evaluation of the SAME bf16-rounded values to fp32-accumulation accuracy (1e-5), which separates
"storage rounding" from "kernel error"."""
import numpy as np
import pytest
import torch

from conftest import Golden, rel_err
from oracle import msda_oracle as O

pytestmark = pytest.mark.gpu


def dev(x):
    t = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
    return t.cuda().contiguous()


def test_pair_value_layout_is_exact():
    import gomatching_b200 as g
    torch.manual_seed(0)
    shapes = torch.tensor([[5, 7], [3, 4], [2, 2], [1, 3]])
    lsi = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    for dtype in (torch.float32, torch.bfloat16):
        value = torch.randn(2, S, 3, 32).to(dtype).cuda()
        paired = g.pair_value_bf16(value, shapes.cuda(), lsi.cuda())
        assert paired.shape == (2, 3, S, 2, 32) and paired.dtype == torch.bfloat16
        want = torch.zeros(2, 3, S, 2, 32, dtype=torch.bfloat16)
        vb = value.to(torch.bfloat16).cpu()                       # round-to-nearest-even, like the kernel
        for l, (h, w) in enumerate(shapes.tolist()):
            for y in range(h):
                for x in range(w):
                    p = int(lsi[l]) + y * w + x
                    want[:, :, p, 0] = vb[:, p]
                    if x + 1 < w:
                        want[:, :, p, 1] = vb[:, p + 1]
        assert torch.equal(paired.cpu(), want)


def _float64_on_bf16_values(c):
    """The reference's arithmetic in float64 on bf16-rounded values (storage rounding isolated from kernel error)."""
    vb = torch.from_numpy(c["value"]).to(torch.bfloat16).float().numpy()
    return O.forward_f64(vb, c["shapes"], c["lsi"], c["loc"], c["attn"])


@pytest.mark.parametrize("name", ["uniform_d32", "edges_d32", "wide_d32"])
def test_paired_core_matches_fp32_reference_and_float64_on_the_same_bf16_values(core_cases, name):
    import gomatching_b200 as g
    c = core_cases.case(name)
    if c["loc"].shape[3] != 4 or c["loc"].shape[4] != 4:
        pytest.skip("paired kernel is instantiated for L = 4, P = 4")
    sh, ls = dev(c["shapes"]), dev(c["lsi"])
    paired = g.pair_value_bf16(dev(c["value"]), sh, ls)
    out = g.ms_deform_attn_forward_paired(paired, sh, ls, dev(c["loc"]), dev(c["attn"]))
    assert out.dtype == torch.bfloat16
    got = out.float().cpu().numpy()
    assert rel_err(got, c["out_f32"]) <= 2e-2                           # north_star bf16 bar vs the fp32 reference
    exact = _float64_on_bf16_values(c)
    # the only remaining differences: fp32 accumulation and ONE final bf16 rounding of the output (2^-9 relative)
    assert rel_err(got, exact) <= 2.0 ** -8
    out_from_bf16 = g.ms_deform_attn_forward_paired(g.pair_value_bf16(dev(c["value"]).bfloat16(), sh, ls), sh, ls,
                                                    dev(c["loc"]), dev(c["attn"]))
    assert torch.equal(out, out_from_bf16)                              # fp32 and bf16 inputs pair to the same bits


@pytest.mark.parametrize("kind,n", [("encoder", 1), ("decoder", 2)])
def test_paired_full_size_720p(kind, n):
    """BASELINE.json config 3 shape (decoder, 100 x 25 point queries) and the encoder shape, fused entry: against the
    fp32 fused kernel (2e-2) and against float64 on the bf16-rounded values with an fp32-rounded output check."""
    import gomatching_b200 as g
    from gomatching_b200 import synthetic as syn
    w = syn.make_workload(kind, 720, 1280, n=n, seed=13, dist="local")
    v, sh, ls = w.value.cuda(), w.shapes.cuda(), w.lsi.cuda()
    ref32 = g.ms_deform_attn_forward_fused(v, sh, ls, w.ref.cuda(), w.offsets.cuda(), w.logits.cuda())
    paired = g.pair_value_bf16(v, sh, ls)
    out = g.ms_deform_attn_forward_fused_paired(paired, sh, ls, w.ref.cuda(), w.offsets.cuda(), w.logits.cuda())
    core = g.ms_deform_attn_forward_paired(paired, sh, ls, w.loc.cuda(), w.attn.cuda())
    torch.cuda.synchronize()
    a, b = out.float().cpu().numpy(), ref32.cpu().numpy()
    assert rel_err(a, b) <= 2e-2
    assert rel_err(core.float().cpu().numpy(), b) <= 2e-2
    # fused and core entries of the paired kernel agree to the softmax's 1e-6
    assert rel_err(a, core.float().cpu().numpy()) <= 2.0 ** -7
    # zero padding / borders at full size: the uniform distribution puts ~30 % of the samples outside the maps
    wu = syn.make_workload(kind, 720, 1280, n=1, seed=14, dist="uniform")
    vu = wu.value.cuda()
    pu = g.pair_value_bf16(vu, sh, ls)
    got = g.ms_deform_attn_forward_paired(pu, sh, ls, wu.loc.cuda(), wu.attn.cuda()).float().cpu().numpy()
    want = g.ms_deform_attn_forward(vu.bfloat16().float(), sh, ls, wu.loc.cuda(), wu.attn.cuda(), 64).cpu().numpy()
    assert rel_err(got, want) <= 2.0 ** -8


@pytest.mark.parametrize("name", ["ref2_mask", "ref4_nomask"])
def test_module_in_paired_bf16_mode_stays_within_the_bf16_bar(module_cases, name):
    """MSDeformAttn.forward with ``paired_bf16_value`` against the REFERENCE module's fp32 output (golden fixture)."""
    import gomatching_b200 as g
    c = module_cases.case(name)
    d_model, levels, heads, points = (int(v) for v in c["cfg"])
    if (d_model // heads, levels, points) != (32, 4, 4):
        pytest.skip("fixture is not the D = 32, L = 4, P = 4 shape")
    mod = g.MSDeformAttn(d_model, levels, heads, points)
    mod.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in c.items() if k.startswith("sd/")}, strict=True)
    mod = mod.cuda().eval()
    mod.paired_bf16_value = True
    mask = dev(c["mask"]) if c["mask"].size else None
    with torch.no_grad():
        out = mod(dev(c["query"]), dev(c["ref"]), dev(c["src"]), dev(c["shapes"]), dev(c["lsi"]), mask)
    assert out.dtype == torch.float32
    assert rel_err(out.cpu().numpy(), c["out"]) <= 2e-2
