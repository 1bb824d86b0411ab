"""CPU tests: pin the oracle against outputs of the reference itself (tests/golden/*.npz, produced by
tests/golden/make_golden.py from the reference's ms_deform_attn_core_pytorch / MSDeformAttn /
DeformableTransformer).  The reference ships no golden vectors for this path."""
import numpy as np
import pytest
import torch

from conftest import Golden, rel_err
from oracle import msda_oracle as O

CORE = Golden("core_cases.npz").names()
MODULE = Golden("module_cases.npz").names()


def test_oracle_builds_and_loads():
    assert O.lib().msda_oracle_version() >= 1
    assert O.max_threads() >= 1


@pytest.mark.parametrize("name", CORE)
def test_kernel_restatement_matches_reference_fp32(core_cases, name):
    c = core_cases.case(name)
    out = O.forward_f32(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"])
    # fp32 tolerance of north_star: 1e-4 relative (max|a-b|/max|b|); measured ~1e-6
    assert rel_err(out, c["out_f32"]) <= 1e-4
    assert rel_err(out, c["out_f64"]) <= 1e-4


@pytest.mark.parametrize("name", CORE)
def test_kernel_restatement_matches_reference_fp64(core_cases, name):
    c = core_cases.case(name)
    out = O.forward_f64(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"])
    assert rel_err(out, c["out_f64"]) <= 1e-12


@pytest.mark.parametrize("name", CORE)
def test_exact_evaluation_matches_reference_fp64(core_cases, name):
    c = core_cases.case(name)
    out = O.forward_exact(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"])
    assert rel_err(out, c["out_f64"]) <= 1e-12


@pytest.mark.parametrize("name", CORE)
def test_gridsample_port_equals_reference_cpu_path(core_cases, name):
    c = core_cases.case(name)
    out = O.core_gridsample(torch.from_numpy(c["value"]), c["shapes"].tolist(), torch.from_numpy(c["loc"]),
                            torch.from_numpy(c["attn"])).numpy()
    # same torch ops in the same order as ms_deform_attn.py:40-60 -> equal up to thread-count effects
    assert rel_err(out, c["out_f32"]) <= 1e-6


@pytest.mark.parametrize("name", CORE)
def test_bf16_oracle_within_bf16_tolerance(core_cases, name):
    c = core_cases.case(name)
    vb = O.f32_to_bf16_bits(c["value"])
    out = O.bf16_bits_to_f32(O.forward_bf16(vb, c["shapes"], c["lsi"], c["loc"], c["attn"]))
    # bf16 tolerance of north_star: 2e-2 relative against the fp32 reference on the same inputs
    assert rel_err(out, c["out_f32"]) <= 2e-2
    # and it is exactly "fp32 kernel on up-cast bf16 value, rounded once"
    ref = O.forward_f32(O.bf16_bits_to_f32(vb), c["shapes"], c["lsi"], c["loc"], c["attn"])
    assert np.array_equal(O.f32_to_bf16_bits(ref), O.f32_to_bf16_bits(out))


def test_bf16_rounding_helpers_match_torch():
    x = torch.randn(4096, generator=torch.Generator().manual_seed(3)) * 37.0
    bits = O.f32_to_bf16_bits(x.numpy())
    assert np.array_equal(bits, x.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16))
    assert np.array_equal(O.bf16_bits_to_f32(bits), x.to(torch.bfloat16).float().numpy())


def test_index_oracle_on_edge_locations(core_cases):
    c = core_cases.case("edges_d32")
    idx = O.sample_index(c["loc"], c["shapes"], c["lsi"], M=8, D=32)
    L = c["shapes"].shape[0]
    for l in range(L):
        H, W = (int(v) for v in c["shapes"][l])
        # fused multiply-add == exact fp64 product and subtraction, rounded ONCE to fp32
        x = (c["loc"][:, :, :, l, :, 0].astype(np.float64) * W - 0.5).astype(np.float32)
        y = (c["loc"][:, :, :, l, :, 1].astype(np.float64) * H - 0.5).astype(np.float32)
        inr = (x > -1) & (y > -1) & (x < W) & (y < H)
        rec = idx[:, :, :, l, :]
        assert np.array_equal(rec["in_range"].astype(bool), inr)
        assert np.array_equal(rec["w_low"][inr], np.floor(x[inr]).astype(np.int32))
        assert np.array_equal(rec["h_low"][inr], np.floor(y[inr]).astype(np.int32))
        assert np.all(rec["level_offset"] == int(c["lsi"][l]) * 8 * 32)
        m = rec["corner_mask"][inr]
        hl, wl = rec["h_low"][inr], rec["w_low"][inr]
        assert np.array_equal((m & 1) != 0, (hl >= 0) & (wl >= 0))
        assert np.array_equal((m & 8) != 0, (hl + 1 <= H - 1) & (wl + 1 <= W - 1))


def test_index_oracle_is_the_fused_rounding():
    """h_low = floor(fmaf(loc, size, -0.5)) -- one rounding (reference SASS: FFMA R, size, loc, -0.5)."""
    rng = np.random.default_rng(0)
    W = 160
    loc = np.zeros((1, 200000, 1, 1, 1, 2), dtype=np.float32)
    loc[..., 0] = rng.random((1, 200000, 1, 1, 1), dtype=np.float32)
    loc[..., 1] = 0.5
    idx = O.sample_index(loc, [[8, W]], [0], M=1, D=1)
    exact = np.floor((loc[..., 0].astype(np.float64) * W - 0.5).astype(np.float32)).astype(np.int32)
    assert np.array_equal(idx["w_low"], exact)          # fp64 product + one fp32 rounding == fmaf
    # a location engineered so that the two-step fp32 form flips the floor: (k+0.5)/W just below
    k = 37
    v = np.float32((k + 0.5) / W)
    v = np.nextafter(v, np.float32(0), dtype=np.float32)
    loc2 = np.zeros((1, 1, 1, 1, 1, 2), dtype=np.float32)
    loc2[..., 0], loc2[..., 1] = v, 0.5
    got = int(O.sample_index(loc2, [[8, W]], [0], M=1, D=1)["w_low"].ravel()[0])
    assert got == int(np.floor(np.float32(np.float64(v) * W - 0.5)))


@pytest.mark.parametrize("name", MODULE)
def test_module_restatement_matches_reference_module(module_cases, name):
    c = module_cases.case(name)
    d_model, levels, heads, points = (int(v) for v in c["cfg"])
    sd = {k[3:]: torch.from_numpy(v) for k, v in c.items() if k.startswith("sd/")}
    mask = torch.from_numpy(c["mask"]) if c["mask"].size else None
    for core in ("kernel", "gridsample"):
        out, loc, aw = O.module_forward(sd, torch.from_numpy(c["query"]), torch.from_numpy(c["ref"]),
                                        torch.from_numpy(c["src"]), c["shapes"], c["lsi"], mask,
                                        n_heads=heads, n_levels=levels, n_points=points, core=core)
        assert rel_err(out.numpy(), c["out"]) <= 1e-4
        assert rel_err(loc.numpy(), c["loc"]) <= 1e-6
        assert rel_err(aw.numpy(), c["attn"]) <= 1e-6


@pytest.mark.parametrize("name", MODULE)
def test_location_and_softmax_glue(module_cases, name):
    c = module_cases.case(name)
    d_model, levels, heads, points = (int(v) for v in c["cfg"])
    sd = {k[3:]: torch.from_numpy(v) for k, v in c.items() if k.startswith("sd/")}
    q = torch.from_numpy(c["query"])
    N, Lq, _ = q.shape
    off = torch.nn.functional.linear(q, sd["sampling_offsets.weight"], sd["sampling_offsets.bias"]).view(
        N, Lq, heads, levels, points, 2).numpy()
    logits = torch.nn.functional.linear(q, sd["attention_weights.weight"], sd["attention_weights.bias"]).view(
        N, Lq, heads, levels * points).numpy()
    loc = O.locations(c["ref"], off, c["shapes"])
    # the same IEEE operations in the same order as the eager reference: bit-exact
    assert np.array_equal(loc, c["loc"])
    aw = O.softmax(logits).reshape(c["attn"].shape)
    assert rel_err(aw, c["attn"]) <= 1e-6
    with pytest.raises(ValueError):
        O.locations(np.zeros((N, Lq, levels, 3), np.float32), off, c["shapes"])


@pytest.mark.parametrize("tag", ["enc0", "dec0"])
def test_setC_default_init_network_capture(setc_cases, tag):
    c = setc_cases.case(tag)
    out = O.forward_f32(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"])
    assert rel_err(out, c["out"]) <= 1e-4
    if tag == "enc0":
        # default init: uniform attention 1/16 and offsets of exactly k pixels (ms_deform_attn.py:101-109)
        assert np.allclose(c["attn"], 1.0 / 16.0)
        idx = O.sample_index(c["loc"], c["shapes"], c["lsi"], M=8, D=32)
        assert 0.02 < 1.0 - idx["in_range"].mean() < 0.6


BACKWARD = Golden("backward_cases.npz").names()


@pytest.mark.parametrize("name", BACKWARD)
def test_gridsample_port_gradients_match_reference_autograd(name):
    """The backward checker used on the GPU box (autograd through oracle.core_gridsample) reproduces the
    gradients the reference's own CPU path gives (tests/golden/make_golden_backward.py)."""
    c = Golden("backward_cases.npz").case(name)
    value = torch.from_numpy(c["value"]).double().requires_grad_(True)
    loc = torch.from_numpy(c["loc"]).double().requires_grad_(True)
    attn = torch.from_numpy(c["attn"]).double().requires_grad_(True)
    out = O.core_gridsample(value, c["shapes"].tolist(), loc, attn)
    out.backward(torch.from_numpy(c["grad_out"]).double())
    # fixtures were produced from float64 inputs later stored as float32 -> compare at float32 input precision
    assert rel_err(out.detach().numpy(), c["out"]) <= 1e-5
    assert rel_err(value.grad.numpy(), c["grad_value"]) <= 1e-5
    assert rel_err(loc.grad.numpy(), c["grad_loc"]) <= 1e-4
    assert rel_err(attn.grad.numpy(), c["grad_attn"]) <= 1e-5
