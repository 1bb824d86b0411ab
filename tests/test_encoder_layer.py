"""Encoder-layer drop-in: state-dict compatibility on the CPU, parity against outputs of the
reference's own DeformableTransformerEncoderLayer (tests/golden/make_golden_encoder_layer.py) on the GPU."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "encoder_layer_cases.npz")
CASES = ["mask_pos", "plain"]


def load(name):
    z = np.load(GOLD)
    c = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "/")}
    sd = {k[3:]: torch.from_numpy(v) for k, v in c.items() if k.startswith("sd/")}
    return c, sd


def build(c, sd):
    from gomatching_b200 import DeformableTransformerEncoderLayer
    d_model, d_ffn, heads, levels, points = [int(v) for v in c["cfg"]]
    layer = DeformableTransformerEncoderLayer(d_model, d_ffn, 0.1, "relu", levels, heads, points).eval()
    missing, unexpected = layer.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return layer


@pytest.mark.parametrize("name", CASES)
def test_state_dict_of_the_reference_layer_loads_unchanged(name):
    c, sd = load(name)
    layer = build(c, sd)
    assert sorted(layer.state_dict().keys()) == sorted(sd.keys())
    for k, v in layer.state_dict().items():
        assert v.shape == sd[k].shape, k


def test_cpu_tensors_raise_like_the_reference_operator():
    c, sd = load("plain")
    layer = build(c, sd)
    src = torch.from_numpy(c["src"])
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        layer(src, None, torch.from_numpy(c["ref"]), torch.from_numpy(c["shapes"]), torch.from_numpy(c["lsi"]), None)
    # the feed-forward block alone is plain torch on the CPU (eager reference sequence)
    out = layer.forward_ffn(src)
    assert float((out - torch.from_numpy(c["ffn_only"])).abs().max()) <= 1e-5


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("tc", [True, False])
def test_layer_matches_the_reference_layer_output(name, tc):
    c, sd = load(name)
    layer = build(c, sd).cuda()
    layer.tensor_core_ffn = tc
    layer.self_attn.tensor_core_projections = tc
    src = torch.from_numpy(c["src"]).cuda()
    pos = torch.from_numpy(c["pos"]).cuda() if c["pos"].size else None
    mask = torch.from_numpy(c["mask"]).cuda() if c["mask"].size else None
    ref, sh, lsi = torch.from_numpy(c["ref"]).cuda(), torch.from_numpy(c["shapes"]).cuda(), torch.from_numpy(c["lsi"]).cuda()
    with torch.no_grad():
        out = layer(src, pos, ref, sh, lsi, mask)
        ffn = layer.forward_ffn(src)
    assert rel(out.cpu(), torch.from_numpy(c["out"])) <= 1e-4           # the fp32 bar of the operator
    assert rel(ffn.cpu(), torch.from_numpy(c["ffn_only"])) <= 1e-5


@pytest.mark.gpu
def test_ffn_on_tensor_cores_is_fp32_grade_at_full_width():
    """d_ffn = 1024 (the DeepSolo setting), 3 000 tokens: the 3xTF32 feed-forward block against float64."""
    from gomatching_b200 import DeformableTransformerEncoderLayer
    torch.manual_seed(3)
    layer = DeformableTransformerEncoderLayer(256, 1024, 0.1, "relu", 4, 8, 4).cuda().eval()
    src = torch.randn(2, 1500, 256, device="cuda")
    with torch.no_grad():
        got = layer.forward_ffn(src)
        layer.tensor_core_ffn = False
        eager = layer.forward_ffn(src)
        l64 = DeformableTransformerEncoderLayer(256, 1024, 0.1, "relu", 4, 8, 4).double().cuda().eval()
        l64.load_state_dict({k: v.double() for k, v in layer.state_dict().items()})
        ref = l64.forward_ffn(src.double())
    e_tc, e_eager = rel(got.double(), ref), rel(eager.double(), ref)
    assert e_tc <= 2e-5, e_tc
    assert e_tc <= 20 * max(e_eager, 1e-7)
    # training mode with dropout keeps the eager path (and its randomness)
    layer.tensor_core_ffn = True
    layer.train()
    assert not layer._ffn_on_tensor_cores(src)


@pytest.mark.gpu
@pytest.mark.parametrize("rows,c,with_y,affine", [(19160, 256, True, True), (777, 128, True, True), (3, 1024, False, True),
                                                 (4097, 384, True, False), (64, 256, True, True)])
def test_add_layernorm_matches_torch(rows, c, with_y, affine):
    """out = LayerNorm(x + y): against torch in float64 and against the eager fp32 ops (deformable_transformer.py:251-252)."""
    from gomatching_b200.norm import add_layernorm
    torch.manual_seed(rows + c)
    norm = torch.nn.LayerNorm(c, elementwise_affine=affine).cuda()
    if affine:
        with torch.no_grad():
            norm.weight.normal_(1.0, 0.2)
            norm.bias.normal_(0.0, 0.2)
    x = torch.randn(2, rows, c, device="cuda") * 3 + 0.5
    y = torch.randn(2, rows, c, device="cuda") if with_y else None
    with torch.no_grad():
        got = add_layernorm(x, y, norm)
        eager = norm(x + y if with_y else x)
        n64 = torch.nn.LayerNorm(c, elementwise_affine=affine).double().cuda()
        if affine:
            n64.load_state_dict({k: v.double() for k, v in norm.state_dict().items()})
        ref = n64((x.double() + y.double()) if with_y else x.double())
    assert got.shape == eager.shape and got.dtype == torch.float32
    e_got, e_eager = rel(got.double(), ref), rel(eager.double(), ref)
    assert e_got <= 2e-6, e_got
    assert e_got <= 4 * max(e_eager, 1e-7)


@pytest.mark.gpu
def test_add_layernorm_fallbacks_and_cpu():
    from gomatching_b200.norm import add_layernorm
    norm = torch.nn.LayerNorm(100).cuda()                      # C not a multiple of 128: eager path, same result
    x, y = torch.randn(5, 100, device="cuda"), torch.randn(5, 100, device="cuda")
    assert torch.equal(add_layernorm(x, y, norm), norm(x + y))
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        add_layernorm(torch.randn(2, 256), None, torch.nn.LayerNorm(256))


# ---------------------------------------------------------------------------------------------------
# Decoder layer drop-in (deformable_transformer.py:326-427)
# ---------------------------------------------------------------------------------------------------
DEC_GOLD = os.path.join(HERE, "golden", "decoder_layer_cases.npz")
DEC_CASES = ["dec_mask", "dec_shared_ref"]


def load_dec(name):
    z = np.load(DEC_GOLD)
    c = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "/")}
    sd = {k[3:]: torch.from_numpy(v) for k, v in c.items() if k.startswith("sd/")}
    return c, sd


def build_dec(c, sd):
    from gomatching_b200 import DeformableCompositeTransformerDecoderLayer
    d_model, d_ffn, heads, levels, points = [int(v) for v in c["cfg"]]
    layer = DeformableCompositeTransformerDecoderLayer(d_model, d_ffn, 0.1, "relu", levels, heads, points).eval()
    missing, unexpected = layer.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return layer


@pytest.mark.parametrize("name", DEC_CASES)
def test_decoder_layer_state_dict_loads_unchanged(name):
    c, sd = load_dec(name)
    layer = build_dec(c, sd)
    assert sorted(layer.state_dict().keys()) == sorted(sd.keys())


@pytest.mark.gpu
@pytest.mark.parametrize("name", DEC_CASES)
@pytest.mark.parametrize("fast", [True, False])
def test_decoder_layer_matches_the_reference_layer_output(name, fast):
    c, sd = load_dec(name)
    layer = build_dec(c, sd).cuda()
    layer.tensor_core_ffn = layer.fused_add_norm = layer.fused_self_attention = fast
    layer.attn_cross.tensor_core_projections = fast
    t = lambda k: torch.from_numpy(c[k]).cuda()
    mask = t("mask") if c["mask"].size else None
    with torch.no_grad():
        out = layer(t("tgt"), t("qpos"), t("ref"), t("src"), t("shapes"), t("lsi"), mask)
    assert out.shape == tuple(c["out"].shape)
    assert rel(out.cpu(), torch.from_numpy(c["out"])) <= 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("B,L,bstride,sstride", [(100, 25, 25, 1), (25, 100, 1, 25), (3, 128, 128, 1), (7, 1, 1, 7)])
def test_small_mha_matches_nn_multihead_attention(B, L, bstride, sstride):
    """The decoder's intra / inter self-attention on this library's kernels vs nn.MultiheadAttention in float64, with the
    strided sequence addressing both call sites use (deformable_transformer.py:386-404)."""
    from gomatching_b200 import small_mha
    torch.manual_seed(B * 1000 + L)
    mha = torch.nn.MultiheadAttention(256, 8).cuda().eval()
    with torch.no_grad():
        mha.in_proj_bias.normal_(0, 0.1)
        mha.out_proj.bias.normal_(0, 0.1)
    T = B * L
    tok = torch.randn(T, 256, device="cuda")
    pos = torch.randn(T, 256, device="cuda")
    rows = (torch.arange(B)[:, None] * bstride + torch.arange(L)[None, :] * sstride).cuda()       # (B, L) row of each token
    assert sorted(rows.flatten().tolist()) == list(range(T))
    ref64 = torch.nn.MultiheadAttention(256, 8).cuda().double().eval()
    ref64.load_state_dict({k: v.double() for k, v in mha.state_dict().items()})
    packed = small_mha.PackedProjection()
    with torch.no_grad():
        for v_tokens in (None, tok):                                  # packed q=k=v, and q=k=tok+pos with v=tok
            qk = tok if v_tokens is None else tok + pos
            got = small_mha.self_attention(mha, qk, v_tokens, B, L, bstride, sstride, packed)
            q64 = qk.double()[rows].transpose(0, 1)                   # (L, B, E)
            v64 = tok.double()[rows].transpose(0, 1)
            want = ref64(q64, q64, v64)[0].transpose(0, 1)            # (B, L, E)
            err = float((got.double()[rows] - want).abs().max() / want.abs().max())
            assert err <= 2e-5, (err, v_tokens is None)


@pytest.mark.gpu
@pytest.mark.parametrize("bs,nq,npts,levels", [(1, 100, 25, 4), (2, 300, 25, 4), (3, 7, 5, 1)])
def test_decoder_glue_is_bit_identical_to_the_eager_expressions(bs, nq, npts, levels):
    """csrc/decoder_glue.cu against the eager tensor expressions it replaces -- the positional embedding of the current
    reference points (adet/modeling/model/utils.py:24-37 on reference_points * valid_ratios, level 0) and the
    reference-point refinement (deformable_transformer.py:483-486 with adet/utils/misc.py:115-119): every bit equal,
    including points outside [0, 1], exact 0 / 1, and NaN."""
    import math
    from gomatching_b200 import decoder_glue
    g = torch.Generator(device="cuda").manual_seed(bs * 100 + nq)
    ref = torch.rand(bs, nq, npts, 2, generator=g, device="cuda") * 1.4 - 0.2
    ref.view(-1)[:6] = torch.tensor([0.0, 1.0, -0.0, 1e-6, 1.0 - 1e-6, 0.5], device="cuda")
    vr = torch.rand(bs, levels, 2, generator=g, device="cuda") * 0.5 + 0.5
    d_model, temp = 256, 10000
    got = decoder_glue.point_pos_embed(ref, vr, d_model, temp)
    pts = (ref[:, :, :, None] * vr[:, None, None])[:, :, :, 0, :]
    dim = d_model // 2
    dim_t = torch.arange(dim, dtype=torch.float32, device="cuda")
    dim_t = temp ** (2 * torch.div(dim_t, 2, rounding_mode='trunc') / dim)
    px = (pts[:, :, :, 0] * (2 * math.pi))[:, :, :, None] / dim_t
    py = (pts[:, :, :, 1] * (2 * math.pi))[:, :, :, None] / dim_t
    px = torch.stack((px[:, :, :, 0::2].sin(), px[:, :, :, 1::2].cos()), dim=4).flatten(3)
    py = torch.stack((py[:, :, :, 0::2].sin(), py[:, :, :, 1::2].cos()), dim=4).flatten(3)
    want = torch.cat((px, py), dim=-1)
    assert got.shape == want.shape and torch.equal(got, want)
    assert torch.equal(decoder_glue.point_pos_embed(ref, None, d_model, temp),
                       decoder_glue.point_pos_embed(ref, torch.ones_like(vr), d_model, temp))

    tmp = torch.randn(bs, nq, npts, 2, generator=g, device="cuda") * 3
    refn = ref.clone()
    refn.view(-1)[7] = float("nan")
    x = refn.clamp(min=0, max=1)
    want = (tmp + torch.log(x.clamp(min=1e-5) / (1 - x).clamp(min=1e-5))).sigmoid()
    got = decoder_glue.refine_points(tmp, refn)
    assert torch.equal(torch.isnan(got), torch.isnan(want)) and bool(torch.isnan(got).view(-1)[7])
    assert torch.equal(torch.nan_to_num(got), torch.nan_to_num(want))
    with pytest.raises(RuntimeError):
        decoder_glue.refine_points(tmp.cpu(), refn.cpu())                        # no CPU path
