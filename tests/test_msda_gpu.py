"""GPU parity tests (run on the B200 box with -m gpu).  Everything goes through the C-ABI library
(gomatching_b200/libmsda_b200.so) via the Python operator layer; the checker is the CPU oracle
(oracle/msda_oracle.c), the committed golden fixtures generated from the reference (tests/golden/), and --
when oracle/_ref/libmsda_refcuda.so was built -- the unmodified reference CUDA kernel itself.

Bars (BASELINE.json north_star): sampling indices and level offsets bit-exact; outputs within 1e-4
(fp32) / 2e-2 (bf16) of the reference as max|a-b|/max|b|.  The fp32 and bf16 kernels reproduce the
reference's FMA chain, so most output checks below are in fact bit-exact (np.array_equal).
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import Golden, rel_err
from oracle import msda_oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CORE = Golden("core_cases.npz").names()
MODULE = Golden("module_cases.npz").names()
REFCUDA = os.path.join(ROOT, "oracle", "_ref", "libmsda_refcuda.so")


def dev(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def run_core(c, tuning=None, dtype=torch.float32):
    import gomatching_b200 as g
    out = g.ms_deform_attn_forward(dev(c["value"], dtype), dev(c["shapes"]), dev(c["lsi"]), dev(c["loc"]),
                                   dev(c["attn"]), 64, tuning=tuning)
    torch.cuda.synchronize()
    return out


def bits(t):
    return t.view(torch.int16).cpu().numpy().view(np.uint16)


def test_native_library_is_what_runs():
    from gomatching_b200 import _native
    L = _native.lib()
    assert L.msda_b200_sm_count() >= 100          # B200: 148
    assert os.path.basename(_native.LIB_PATH) == "libmsda_b200.so"


@pytest.mark.parametrize("name", CORE)
def test_core_fp32_bit_exact_vs_oracle_and_in_tolerance_vs_reference(core_cases, name):
    c = core_cases.case(name)
    out = run_core(c).cpu().numpy()
    ora = O.forward_f32(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"])
    assert np.array_equal(out, ora), "fp32 kernel differs from the kernel-arithmetic oracle (max rel %g)" % rel_err(out, ora)
    assert rel_err(out, c["out_f32"]) <= 1e-4      # vs the reference's own ms_deform_attn_core_pytorch output
    assert rel_err(out, c["out_f64"]) <= 1e-4


@pytest.mark.parametrize("name", CORE)
def test_core_fp64_bit_exact_vs_oracle(core_cases, name):
    """The reference also instantiates double (ms_deform_attn_cuda.cu:64)."""
    import gomatching_b200 as g
    c = core_cases.case(name)
    out = g.ms_deform_attn_forward(dev(c["value"]).double(), dev(c["shapes"]), dev(c["lsi"]), dev(c["loc"]).double(),
                                   dev(c["attn"]).double(), 64).cpu().numpy()
    assert np.array_equal(out, O.forward_f64(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"]))
    assert rel_err(out, c["out_f64"]) <= 1e-12


@pytest.mark.parametrize("name", ["uniform_d32", "edges_d32", "d64_p8"])
def test_every_variant_and_mode_gives_identical_bits(core_cases, name):
    from gomatching_b200 import _native
    c = core_cases.case(name)
    base = run_core(c).cpu().numpy()
    for variant in range(_native.lib().msda_b200_variant_count()):
        for mode in (1, 3):
            for tq in (4, 32, 64):
                for v1 in (0, 1):
                    out = run_core(c, dict(mode=mode, variant=variant, tile_q=tq, ctas_per_sm=2, force_v1=v1)).cpu().numpy()
                    assert np.array_equal(out, base), (variant, mode, tq, v1)


@pytest.mark.parametrize("name", ["uniform_d32", "edges_d32", "d64_p8", "nonpow2_d12"])
def test_core_bf16_bit_exact_vs_oracle(core_cases, name):
    c = core_cases.case(name)
    vb = O.f32_to_bf16_bits(c["value"])
    out = run_core(c, dtype=torch.bfloat16)
    ora = O.forward_bf16(vb, c["shapes"], c["lsi"], c["loc"], c["attn"])
    assert np.array_equal(bits(out), ora)
    # north_star bf16 bar: 2e-2 relative vs the fp32 reference on the same inputs
    assert rel_err(out.float().cpu().numpy(), c["out_f32"]) <= 2e-2


@pytest.mark.parametrize("name", ["uniform_d32", "edges_d32", "wide_d32", "nonpow2_d12"])
def test_sampling_indices_and_level_offsets_bit_exact(core_cases, name):
    import gomatching_b200 as g
    c = core_cases.case(name)
    N, S, M, D = c["value"].shape
    got = g.sample_index(dev(c["loc"]), dev(c["shapes"]), dev(c["lsi"]), M, D)
    ora = O.sample_index(c["loc"], c["shapes"], c["lsi"], M=M, D=D)
    for k in ("h_low", "w_low", "in_range", "corner_mask", "level_offset"):
        assert np.array_equal(got[k].cpu().numpy(), ora[k]), k


@pytest.mark.parametrize("tag", ["enc0", "dec0"])
def test_setC_default_init_network_capture(setc_cases, tag):
    """Samples sit exactly on pixel centres: the worst case for the fused-rounding index contract."""
    import gomatching_b200 as g
    c = setc_cases.case(tag)
    out = run_core(c).cpu().numpy()
    assert np.array_equal(out, O.forward_f32(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"]))
    assert rel_err(out, c["out"]) <= 1e-4
    got = g.sample_index(dev(c["loc"]), dev(c["shapes"]), dev(c["lsi"]), 8, 32)
    ora = O.sample_index(c["loc"], c["shapes"], c["lsi"], M=8, D=32)
    for k in ("h_low", "w_low", "in_range", "corner_mask", "level_offset"):
        assert np.array_equal(got[k].cpu().numpy(), ora[k]), k
    if tag == "enc0":      # Lq == S: the pyramid tiling is the default; force the others too
        for tuning in (dict(mode=1), dict(mode=2, tile_h=4, tile_w=4), dict(mode=2, tile_h=16, tile_w=16), dict(mode=3),
                       dict(mode=2, force_v1=1), dict(mode=2, variant=3), dict(mode=2, variant=5, tile_h=2, tile_w=32),
                       dict(mode=4), dict(mode=4, tile_h=2), dict(mode=4, tile_h=3), dict(mode=5)):
            assert np.array_equal(run_core(c, tuning).cpu().numpy(), out), tuning


# ---------------------------------------------------------------------------------------------------
# Full-size workloads (BASELINE.json configs 1-3): oracle on the host, reference CUDA kernel when built
# ---------------------------------------------------------------------------------------------------
def _refcuda():
    if not os.path.exists(REFCUDA):
        return None
    lib = ctypes.CDLL(REFCUDA)
    lib.refcuda_msda_forward_f32.restype = ctypes.c_int
    lib.refcuda_msda_forward_f32.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 7 + [ctypes.c_void_p] * 2
    return lib


def _run_refcuda(lib, w):
    N, S, M, D, L, Lq, P = w.dims
    v, sh, ls, loc, at = dev(w.value), dev(w.shapes), dev(w.lsi), dev(w.loc), dev(w.attn)
    out = torch.empty(N, Lq, M * D, device="cuda")
    rc = lib.refcuda_msda_forward_f32(v.data_ptr(), sh.data_ptr(), ls.data_ptr(), loc.data_ptr(), at.data_ptr(), N, S, M,
                                      D, L, Lq, P, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("kind,dist,n", [("encoder", "local", 1), ("encoder", "uniform", 1), ("decoder", "local", 1),
                                         ("decoder", "uniform", 2), ("encoder", "local", 2)])
def test_full_size_720p_bit_exact(kind, dist, n):
    import gomatching_b200 as g
    from gomatching_b200 import synthetic as syn
    w = syn.make_workload(kind, 720, 1280, n=n, seed=11, dist=dist)
    c = dict(value=w.value.numpy(), shapes=w.shapes.numpy(), lsi=w.lsi.numpy(), loc=w.loc.numpy(), attn=w.attn.numpy())
    out = run_core(c)
    ora = O.forward_f32(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"])
    assert np.array_equal(out.cpu().numpy(), ora)
    if kind == "encoder":   # staged shared-memory-window kernel (windows + global fallback): same bits
        for tuning in (dict(mode=4), dict(mode=4, tile_h=2), dict(mode=5)):
            assert np.array_equal(run_core(c, tuning).cpu().numpy(), ora), tuning
            fz = g.ms_deform_attn_forward_fused(dev(w.value), dev(w.shapes), dev(w.lsi), dev(w.ref), dev(w.offsets),
                                                dev(w.logits), tuning=tuning)
            fz0 = g.ms_deform_attn_forward_fused(dev(w.value), dev(w.shapes), dev(w.lsi), dev(w.ref), dev(w.offsets),
                                                 dev(w.logits), tuning=dict(mode=1))
            assert torch.equal(fz, fz0), tuning
    # index parity at full size
    got = g.sample_index(dev(w.loc), dev(w.shapes), dev(w.lsi), 8, 32)
    oi = O.sample_index(c["loc"], c["shapes"], c["lsi"], M=8, D=32)
    for k in ("h_low", "w_low", "in_range", "corner_mask", "level_offset"):
        assert np.array_equal(got[k].cpu().numpy(), oi[k]), k
    # the reference's CPU path (grid_sample port) agrees within the fp32 bar
    gs = O.core_gridsample(w.value, w.shapes.tolist(), w.loc, w.attn).numpy()
    assert rel_err(out.cpu().numpy(), gs) <= 1e-4
    ref = _refcuda()
    if ref is not None:    # the unmodified reference kernel, compiled for sm_100a from /root/reference
        assert torch.equal(out, _run_refcuda(ref, w)), "differs from the reference CUDA kernel"
    # bf16 storage: bit-exact vs the oracle, inside 2e-2 of fp32
    ob = run_core(c, dtype=torch.bfloat16)
    assert np.array_equal(bits(ob), O.forward_bf16(O.f32_to_bf16_bits(c["value"]), c["shapes"], c["lsi"], c["loc"], c["attn"]))
    assert rel_err(ob.float().cpu().numpy(), ora) <= 2e-2


def test_reference_cuda_kernel_available_for_parity():
    """Not a failure if absent (it is built only where /root/reference exists) -- but say so loudly."""
    if _refcuda() is None:
        pytest.skip("oracle/_ref/libmsda_refcuda.so not built: parity vs the reference CUDA binary not exercised")


def test_1080p_encoder_properties():
    """Config 5 size: batch-split invariance, query-permutation equivariance, linearity in value."""
    import gomatching_b200 as g
    from gomatching_b200 import synthetic as syn
    w = syn.make_workload("encoder", 1080, 1920, n=2, seed=5, dist="local")
    v, sh, ls, loc, at = dev(w.value), dev(w.shapes), dev(w.lsi), dev(w.loc), dev(w.attn)
    out = g.ms_deform_attn_forward(v, sh, ls, loc, at, 64)
    # (1) a batch of 2 equals two batches of 1, bit for bit
    for b in range(2):
        ob = g.ms_deform_attn_forward(v[b:b + 1].contiguous(), sh, ls, loc[b:b + 1].contiguous(), at[b:b + 1].contiguous(), 64)
        assert torch.equal(ob[0], out[b])
    # (2) permuting the queries permutes the output rows (tiling never mixes queries)
    perm = torch.randperm(loc.shape[1], generator=torch.Generator().manual_seed(3)).cuda()
    op = g.ms_deform_attn_forward(v, sh, ls, loc[:, perm].contiguous(), at[:, perm].contiguous(), 64)
    assert torch.equal(op, out[:, perm])
    # (3) linear in value: f(2v) == 2 f(v) exactly (power-of-two scaling commutes with every rounding)
    assert torch.equal(g.ms_deform_attn_forward(v * 2, sh, ls, loc, at, 64), out * 2)
    # (4) value == 1 everywhere: out = sum over in-range samples of attn * (sum of valid corner weights) <= 1
    ones = torch.ones_like(v)
    o1 = g.ms_deform_attn_forward(ones, sh, ls, loc, at, 64)
    assert float(o1.max()) <= 1.0 + 1e-5 and float(o1.min()) >= 0.0
    assert torch.equal(o1.view(2, -1, 8, 32)[..., :1].expand(-1, -1, -1, 32), o1.view(2, -1, 8, 32))


@pytest.mark.parametrize("hw,n,dist", [((720, 1280), 1, "uniform"), ((1080, 1920), 1, "local"), ((192, 320), 3, "local"),
                                       ((96, 96), 2, "uniform")])
def test_staged_window_kernel_is_bit_identical(hw, n, dist):
    """tuning.mode = 4 (value windows in shared memory, msda_forward_staged.cu): window placement, zero-filled borders
    and the global fallback for samples outside a window never change a bit -- uniform locations put most samples
    outside the windows, tiny maps make windows larger than the map, box reference points use the other offset rule."""
    import gomatching_b200 as g
    from gomatching_b200 import synthetic as syn
    w = syn.make_workload("encoder", hw[0], hw[1], n=n, seed=21, dist=dist)
    v, sh, ls, loc, at = dev(w.value), dev(w.shapes), dev(w.lsi), dev(w.loc), dev(w.attn)
    base = g.ms_deform_attn_forward(v, sh, ls, loc, at, 64, tuning=dict(mode=1))
    rf, off, lg = dev(w.ref), dev(w.offsets), dev(w.logits)
    fbase = g.ms_deform_attn_forward_fused(v, sh, ls, rf, off, lg, tuning=dict(mode=1))
    box = torch.cat([rf, torch.full_like(rf, 0.05)], -1).contiguous()             # (N, Lq, L, 4): centre + (w, h)
    bbase = g.ms_deform_attn_forward_fused(v, sh, ls, box, off, lg, tuning=dict(mode=1))
    for tn in (dict(mode=4, tile_h=1), dict(mode=4, tile_h=2), dict(mode=4, tile_h=3), dict(mode=5)):
        assert torch.equal(g.ms_deform_attn_forward(v, sh, ls, loc, at, 64, tuning=tn), base), tn
        assert torch.equal(g.ms_deform_attn_forward_fused(v, sh, ls, rf, off, lg, tuning=tn), fbase), tn
        assert torch.equal(g.ms_deform_attn_forward_fused(v, sh, ls, box, off, lg, tuning=tn), bbase), tn
    # decoder shapes (Lq != S) and bf16 storage fall back to the register-gather kernel under the same tuning
    d = syn.make_workload("decoder", 192, 320, n=1, seed=2, dist="local")
    a = g.ms_deform_attn_forward(dev(d.value), dev(d.shapes), dev(d.lsi), dev(d.loc), dev(d.attn), 64, tuning=dict(mode=4))
    b = g.ms_deform_attn_forward(dev(d.value), dev(d.shapes), dev(d.lsi), dev(d.loc), dev(d.attn), 64)
    assert torch.equal(a, b)
    vb = v.to(torch.bfloat16)
    assert torch.equal(g.ms_deform_attn_forward(vb, sh, ls, loc, at, 64, tuning=dict(mode=4)),
                       g.ms_deform_attn_forward(vb, sh, ls, loc, at, 64))


@pytest.mark.parametrize("M,shapes", [(4, [(40, 56), (20, 28), (10, 14), (5, 7)]), (16, [(24, 33), (12, 17), (6, 9), (3, 5)]),
                                      (8, [(7, 9), (4, 5), (2, 3), (1, 2)])])
def test_window_kernels_with_other_head_counts_and_odd_maps(M, shapes):
    """The window kernels take any number of heads (tensor-map dimension M) and odd / tiny pyramids."""
    import gomatching_b200 as g
    gen = torch.Generator().manual_seed(M)
    S = sum(h * w for h, w in shapes)
    N, L, P, D = 2, 4, 4, 32
    value = torch.randn(N, S, M, D, generator=gen).cuda()
    sh = torch.as_tensor(shapes, dtype=torch.long).cuda()
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    ref = torch.rand(N, S, L, 2, generator=gen).cuda()
    off = (torch.randn(N, S, M, L, P, 2, generator=gen) * 3).cuda()
    lg = torch.randn(N, S, M, L * P, generator=gen).cuda()
    loc, attn = g.locations_softmax(sh, ref, off, lg, lanes_per_unit=8)
    base = g.ms_deform_attn_forward(value, sh, lsi, loc, attn, 64, tuning=dict(mode=1))
    fbase = g.ms_deform_attn_forward_fused(value, sh, lsi, ref, off, lg, tuning=dict(mode=1))
    for tn in (dict(mode=4), dict(mode=5)):
        assert torch.equal(g.ms_deform_attn_forward(value, sh, lsi, loc, attn, 64, tuning=tn), base), tn
        assert torch.equal(g.ms_deform_attn_forward_fused(value, sh, lsi, ref, off, lg, tuning=tn), fbase), tn


# ---------------------------------------------------------------------------------------------------
# Fused glue: softmax + offsets->locations inside the sampler
# ---------------------------------------------------------------------------------------------------
def _eager_glue(w_ref, off, logits, shapes, P):
    """ms_deform_attn.py:138-147 with torch on the GPU (what the reference executes)."""
    N, Lq, M, L, _, _ = off.shape
    attn = torch.softmax(logits, -1).view(N, Lq, M, L, P)
    if w_ref.shape[-1] == 2:
        norm = torch.stack([shapes[..., 1], shapes[..., 0]], -1)
        loc = w_ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = w_ref[:, :, None, :, None, :2] + off / P * w_ref[:, :, None, :, None, 2:] * 0.5
    return loc.contiguous(), attn.contiguous()


@pytest.mark.parametrize("ref_dim", [2, 4])
@pytest.mark.parametrize("lanes", [8, 4])
def test_glue_kernel_locations_bit_exact_softmax_matches_torch(ref_dim, lanes):
    import gomatching_b200 as g
    from gomatching_b200 import synthetic as syn
    w = syn.make_workload("decoder", 720, 1280, n=2, seed=21, dist="local")
    ref = w.ref
    if ref_dim == 4:
        gen = torch.Generator().manual_seed(2)
        ref = torch.cat([w.ref, torch.rand(w.ref.shape, generator=gen) * 0.3 + 0.05], -1).contiguous()
    sh, off, lg, rf = dev(w.shapes), dev(w.offsets), dev(w.logits), dev(ref)
    loc, attn = g.locations_softmax(sh, rf, off, lg, lanes_per_unit=lanes)
    eloc, eattn = _eager_glue(rf, off, lg, sh, 4)
    assert torch.equal(loc, eloc), "locations differ from the eager reference arithmetic"
    # and therefore the CPU oracle's restatement too
    assert np.array_equal(loc.cpu().numpy(), O.locations(ref.numpy(), w.offsets.numpy(), w.shapes.numpy()))
    # softmax: same operation order as torch's warp softmax -> expect identical bits; bar is 1e-6 relative
    assert rel_err(attn.cpu().numpy(), eattn.cpu().numpy()) <= 1e-6
    frac_equal = float((attn == eattn).float().mean())
    print("softmax bit-identical fraction vs torch: %.6f" % frac_equal)
    assert frac_equal > 0.99


@pytest.mark.parametrize("kind,dtype", [("encoder", torch.float32), ("decoder", torch.float32),
                                        ("encoder", torch.bfloat16), ("decoder", torch.bfloat16)])
def test_fused_equals_unfused(kind, dtype):
    import gomatching_b200 as g
    from gomatching_b200 import synthetic as syn
    w = syn.make_workload(kind, 360, 640, n=2, seed=4, dist="local")
    v, sh, ls = dev(w.value, dtype), dev(w.shapes), dev(w.lsi)
    rf, off, lg = dev(w.ref), dev(w.offsets), dev(w.logits)
    fused = g.ms_deform_attn_forward_fused(v, sh, ls, rf, off, lg)
    lanes = 8 if dtype == torch.float32 else 4
    loc, attn = g.locations_softmax(sh, rf, off, lg, lanes_per_unit=lanes)
    unfused = g.ms_deform_attn_forward(v, sh, ls, loc, attn, 64)
    assert torch.equal(fused, unfused)          # same arithmetic, same order -> same bits
    eloc, eattn = _eager_glue(rf, off, lg, sh, 4)
    eager = g.ms_deform_attn_forward(v, sh, ls, eloc, eattn, 64)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert rel_err(fused.float().cpu().numpy(), eager.float().cpu().numpy()) <= tol
    # against the CPU oracle end to end (fp32 only, it is the reference arithmetic)
    if dtype == torch.float32:
        ora = O.forward_f32(w.value.numpy(), w.shapes.numpy(), w.lsi.numpy(),
                            O.locations(w.ref.numpy(), w.offsets.numpy(), w.shapes.numpy()),
                            O.softmax(w.logits.numpy()).reshape(w.attn.shape))
        assert rel_err(fused.cpu().numpy(), ora) <= 1e-5


# ---------------------------------------------------------------------------------------------------
# Module drop-in
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", MODULE)
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("tc", [True, False])
def test_module_matches_reference_module_output(module_cases, name, fused, tc):
    import gomatching_b200 as g
    c = module_cases.case(name)
    d_model, levels, heads, points = (int(v) for v in c["cfg"])
    mod = g.MSDeformAttn(d_model, levels, heads, points)
    mod.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in c.items() if k.startswith("sd/")}, strict=True)
    mod = mod.cuda().eval()
    mod.use_fused = fused
    mod.tensor_core_projections = tc      # 3xTF32 tcgen05 projections vs F.linear (cuBLAS fp32)
    mask = dev(c["mask"]) if c["mask"].size else None
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        out = mod(dev(c["query"]), dev(c["ref"]), dev(c["src"]), dev(c["shapes"]), dev(c["lsi"]), mask)
    assert rel_err(out.cpu().numpy(), c["out"]) <= 1e-4      # reference MSDeformAttn.forward output (fp32 bar)


def test_module_error_behaviour_matches_reference():
    import gomatching_b200 as g
    m = g.MSDeformAttn(256, 4, 8, 4).cuda()
    sh = torch.tensor([[4, 4], [2, 2], [1, 1], [1, 1]]).cuda()
    ls = torch.tensor([0, 16, 20, 21]).cuda()
    with pytest.raises(ValueError, match="must be 2 or 4"):
        m(torch.zeros(1, 3, 256).cuda(), torch.zeros(1, 3, 4, 3).cuda(), torch.zeros(1, 22, 256).cuda(), sh, ls)
    # non-contiguous operator input (ms_deform_attn_cuda.cu:28)
    v = torch.zeros(1, 22, 32, 8).cuda().transpose(2, 3)
    with pytest.raises(RuntimeError, match="value tensor has to be contiguous"):
        g.ms_deform_attn_forward(v, sh, ls, torch.zeros(1, 3, 8, 4, 4, 2).cuda(), torch.zeros(1, 3, 8, 4, 4).cuda(), 64)
    # batch rule (ms_deform_attn_cuda.cu:50-52): N=65 is rejected exactly like the reference
    with pytest.raises(RuntimeError, match="must divide im2col_step"):
        g.ms_deform_attn_forward(torch.zeros(65, 22, 8, 32).cuda(), sh, ls, torch.zeros(65, 3, 8, 4, 4, 2).cuda(),
                                 torch.zeros(65, 3, 8, 4, 4).cuda(), 64)
    m.strict_shape_check = True
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 3, 256).cuda(), torch.zeros(1, 3, 4, 2).cuda(), torch.zeros(1, 23, 256).cuda(), sh, ls)


def test_batch_64_and_empty_like_edges(core_cases):
    """N=64 (= im2col_step) in one launch; a level of 1x1; queries whose samples are all out of range."""
    import gomatching_b200 as g
    gen = torch.Generator().manual_seed(9)
    shapes = [(3, 5), (1, 1)]
    S = 16
    v = torch.randn(64, S, 2, 32, generator=gen)
    loc = torch.rand(64, 7, 2, 2, 4, 2, generator=gen) * 1.4 - 0.2
    loc[:, 0] = 5.0                       # every sample of query 0 is out of range -> exact zeros
    loc[:, 1] = float("nan")              # NaN locations fail every comparison -> skipped (cuh:288)
    attn = torch.softmax(torch.randn(64, 7, 2, 8, generator=gen), -1).view(64, 7, 2, 2, 4)
    sh = torch.tensor(shapes)
    ls = torch.tensor([0, 15])
    out = g.ms_deform_attn_forward(v.cuda(), sh.cuda(), ls.cuda(), loc.cuda(), attn.cuda(), 64).cpu().numpy()
    ora = O.forward_f32(v.numpy(), sh.numpy(), ls.numpy(), loc.numpy(), attn.numpy())
    assert np.array_equal(out, ora)
    assert not out[:, :2].any()


def test_cuda_graph_capture_and_side_stream(core_cases):
    """The call only enqueues work on the current stream: capturable, no hidden sync or allocation."""
    import gomatching_b200 as g
    c = core_cases.case("uniform_d32")
    v, sh, ls, loc, at = dev(c["value"]), dev(c["shapes"]), dev(c["lsi"]), dev(c["loc"]), dev(c["attn"])
    eager = g.ms_deform_attn_forward(v, sh, ls, loc, at, 64)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        warm = g.ms_deform_attn_forward(v, sh, ls, loc, at, 64)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        captured = g.ms_deform_attn_forward(v, sh, ls, loc, at, 64)
    captured.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(captured, eager) and torch.equal(warm, eager)


def test_host_buffer_entry_point(core_cases):
    """msda_b200_forward_f32_host: the call a non-PyTorch host makes."""
    from gomatching_b200 import _native
    L = _native.lib()
    c = core_cases.case("uniform_d32")
    N, S, M, D = c["value"].shape
    _, Lq, _, Lv, P, _ = c["loc"].shape
    ctx = ctypes.c_void_p()
    assert L.msda_b200_host_ctx_create(ctypes.byref(ctx), 0) == 0
    out = np.empty((N, Lq, M * D), np.float32)
    arrs = [np.ascontiguousarray(c[k]) for k in ("value", "shapes", "lsi", "loc", "attn")]
    rc = L.msda_b200_forward_f32_host(ctx, *[a.ctypes.data_as(ctypes.c_void_p) for a in arrs], N, S, M, D, Lv, Lq, P,
                                      out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    L.msda_b200_host_ctx_destroy(ctx)
    assert np.array_equal(out, O.forward_f32(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"]))


# ---------------------------------------------------------------------------------------------------
# Backward (the "next" row a13): vs the reference CPU path under autograd
# ---------------------------------------------------------------------------------------------------
BACKWARD = Golden("backward_cases.npz").names()


@pytest.mark.parametrize("name", BACKWARD)
def test_backward_matches_reference_autograd(name):
    import gomatching_b200 as g
    c = Golden("backward_cases.npz").case(name)
    value = dev(c["value"]).requires_grad_(True)
    loc = dev(c["loc"]).requires_grad_(True)
    attn = dev(c["attn"]).requires_grad_(True)
    out = g.MSDeformAttnFunction.apply(value, dev(c["shapes"]), dev(c["lsi"]), loc, attn, 64)
    out.backward(dev(c["grad_out"]))
    torch.cuda.synchronize()
    # fp32 kernels vs float64 reference gradients: the north_star fp32 bar (1e-4, max|a-b|/max|b|)
    assert rel_err(out.detach().cpu().numpy(), c["out"]) <= 1e-4
    assert rel_err(value.grad.cpu().numpy(), c["grad_value"]) <= 1e-4
    assert rel_err(attn.grad.cpu().numpy(), c["grad_attn"]) <= 1e-4
    # grad_loc multiplies by the level size (up to 160 px): same relative bar against its own scale
    assert rel_err(loc.grad.cpu().numpy(), c["grad_loc"]) <= 1e-4


def test_backward_full_size_decoder_vs_autograd_of_cpu_path():
    import gomatching_b200 as g
    from gomatching_b200 import synthetic as syn
    w = syn.make_workload("decoder", 360, 640, n=1, seed=8, dist="local")
    gen = torch.Generator().manual_seed(1)
    go = torch.randn(1, w.loc.shape[1], 256, generator=gen)
    v = w.value.clone().requires_grad_(True)
    lo = w.loc.clone().requires_grad_(True)
    at = w.attn.clone().requires_grad_(True)
    O.core_gridsample(v, w.shapes.tolist(), lo, at).backward(go)
    vg, lg, ag = dev(w.value).requires_grad_(True), dev(w.loc).requires_grad_(True), dev(w.attn).requires_grad_(True)
    g.MSDeformAttnFunction.apply(vg, dev(w.shapes), dev(w.lsi), lg, ag, 64).backward(go.cuda())
    assert rel_err(vg.grad.cpu().numpy(), v.grad.numpy()) <= 1e-4
    assert rel_err(ag.grad.cpu().numpy(), at.grad.numpy()) <= 1e-4
    assert rel_err(lg.grad.cpu().numpy(), lo.grad.numpy()) <= 1e-3      # fp32 vs fp32, location gradients x160


def test_module_trains_through_the_autograd_function():
    """With grad enabled the module takes the reference's eager glue + MSDeformAttnFunction path."""
    import gomatching_b200 as g
    m = g.MSDeformAttn(256, 4, 8, 4).cuda()
    sh = torch.tensor([[8, 12], [4, 6], [2, 3], [1, 2]]).cuda()
    ls = torch.tensor([0, 96, 120, 126]).cuda()
    q = torch.randn(2, 5, 256, device="cuda", requires_grad=True)
    src = torch.randn(2, 128, 256, device="cuda", requires_grad=True)
    ref = torch.rand(2, 5, 4, 2, device="cuda")
    out = m(q, ref, src, sh, ls)
    out.square().mean().backward()
    assert q.grad is not None and src.grad is not None and m.value_proj.weight.grad is not None
    assert torch.isfinite(q.grad).all() and float(src.grad.abs().sum()) > 0
    with torch.no_grad():
        fused = m(q, ref, src, sh, ls)
    assert rel_err(fused.cpu().numpy(), out.detach().cpu().numpy()) <= 1e-5


def test_nccl_gather_of_frame_records_two_gpus(tmp_path):
    """The exchange step on real NVLink (needs >= 2 GPUs; the driver's scaling run has them)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("single-GPU box")
    import subprocess
    import sys
    code = (
        "import os,torch,torch.distributed as dist\n"
        "from gomatching_b200 import video as V\n"
        "r=int(os.environ['RANK']);torch.cuda.set_device(r);dist.init_process_group('nccl')\n"
        "s=V.RecordSchema(max_instances=8)\n"
        "def spot(f,t):\n"
        "    g=torch.Generator().manual_seed(t);n=t%5\n"
        "    return ({'reid_features':torch.randn(n,1024,generator=g).cuda(),'pred_boxes':torch.rand(n,4,generator=g).cuda(),"
        "'scores':torch.rand(n,generator=g).cuda(),'pred_classes':torch.zeros(n,dtype=torch.int64).cuda(),"
        "'ctrl_points':torch.rand(n,50,generator=g).cuda(),'recs':torch.randint(0,9,(n,25),generator=g).cuda(),"
        "'bd':torch.rand(n,25,4,generator=g).cuda()},(720,1280))\n"
        "def asso(d,st,state):\n"
        "    state=state or []\n"
        "    state+= [(x['frame_index'],float(x['fields']['reid_features'].sum())) for x in d];return state\n"
        "st=V.run_clip(list(range(11)),spot,asso,s,chunk=4,device='cuda')\n"
        "if r==0:\n"
        "    assert [a for a,_ in st]==list(range(11))\n"
        "    import torch as T\n"
        "    for t,v in st: assert abs(v-float(spot(None,t)[0]['reid_features'].sum()))<1e-3\n"
        "    print('ok')\n"
        "dist.destroy_process_group()\n")
    script = tmp_path / "nccl_gather.py"
    script.write_text("import sys\nsys.path.insert(0, %r)\n" % ROOT + code)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True,
                       text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_pitched_fused_call_equals_dense(core_cases):
    """offsets / logits as column slices of one merged 256->384 projection output (no copy)."""
    import gomatching_b200 as g
    from gomatching_b200 import synthetic as syn
    w = syn.make_workload("decoder", 360, 640, n=2, seed=12, dist="local")
    N, S, M, D, L, Lq, P = w.dims
    v, sh, ls, rf = dev(w.value), dev(w.shapes), dev(w.lsi), dev(w.ref)
    dense = g.ms_deform_attn_forward_fused(v, sh, ls, rf, dev(w.offsets), dev(w.logits))
    merged = torch.cat([w.offsets.view(N, Lq, -1), w.logits.view(N, Lq, -1)], -1).cuda()       # (N, Lq, 384)
    off_view = merged[..., :M * L * P * 2].view(N, Lq, M, L, P, 2)
    lg_view = merged[..., M * L * P * 2:].view(N, Lq, M, L * P)
    assert not off_view.is_contiguous() and not lg_view.is_contiguous()
    assert torch.equal(g.ms_deform_attn_forward_fused(v, sh, ls, rf, off_view, lg_view), dense)
    # bf16 storage too
    vb = v.to(torch.bfloat16)
    assert torch.equal(g.ms_deform_attn_forward_fused(vb, sh, ls, rf, off_view, lg_view),
                       g.ms_deform_attn_forward_fused(vb, sh, ls, rf, dev(w.offsets), dev(w.logits)))


def test_module_merged_projection_matches_separate_projections(module_cases):
    import gomatching_b200 as g
    c = module_cases.case("ref2_mask")
    mod = g.MSDeformAttn(256, 4, 8, 4)
    mod.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in c.items() if k.startswith("sd/")}, strict=True)
    mod = mod.cuda().eval()
    args = (dev(c["query"]), dev(c["ref"]), dev(c["src"]), dev(c["shapes"]), dev(c["lsi"]), dev(c["mask"]))
    with torch.no_grad():
        a = mod(*args)
        mod.merge_query_projections = False
        b = mod(*args)
        # the cache follows in-place parameter updates
        mod.merge_query_projections = True
        mod.sampling_offsets.bias.add_(0.25)
        c2 = mod(*args)
        mod.merge_query_projections = False
        d2 = mod(*args)
    assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-5
    assert rel_err(c2.cpu().numpy(), d2.cpu().numpy()) <= 1e-5
    assert rel_err(a.cpu().numpy(), c2.cpu().numpy()) > 1e-4


def test_module_tensor_core_projections_match_cublas_fp32(module_cases):
    """Same module, projections on the tcgen05 tensor cores (3xTF32) vs cuBLAS fp32: outputs agree to fp32 rounding,
    far inside the 1e-4 bar -- so the projections cannot move a sampling index by more than a cuBLAS re-ordering would."""
    import gomatching_b200 as g
    torch.backends.cuda.matmul.allow_tf32 = False
    for name in ("ref2_mask", MODULE[0]):
        c = module_cases.case(name)
        d_model, levels, heads, points = (int(v) for v in c["cfg"])
        mod = g.MSDeformAttn(d_model, levels, heads, points)
        mod.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in c.items() if k.startswith("sd/")}, strict=True)
        mod = mod.cuda().eval()
        mask = dev(c["mask"]) if c["mask"].size else None
        args = (dev(c["query"]), dev(c["ref"]), dev(c["src"]), dev(c["shapes"]), dev(c["lsi"]), mask)
        with torch.no_grad():
            mod.tensor_core_projections = True
            a = mod(*args)
            mod.tensor_core_projections = False
            b = mod(*args)
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 2e-5


# ---------------------------------------------------------------------------------------------------
# The TMA window kernel as the DEFAULT encoder path: host geometry keyed by value, validated on the device
# ---------------------------------------------------------------------------------------------------
def _encoder_workload(h, w, seed, transpose=False):
    """fp32 encoder self-attention workload on an (h, w) frame; ``transpose`` swaps H and W of every level, which keeps
    S (and every other host-visible size) but is a different pyramid."""
    from gomatching_b200 import synthetic as syn
    wk = syn.make_workload("encoder", h, w, n=2, seed=seed, dist="uniform")
    shapes = wk.shapes.clone()
    if transpose:
        shapes = shapes.flip(1).contiguous()
    return dict(value=wk.value.numpy(), shapes=shapes.numpy(), lsi=wk.lsi.numpy(), loc=wk.loc.numpy(), attn=wk.attn.numpy())


def test_window_kernel_is_the_default_and_its_shape_guard_holds(monkeypatch):
    """(1) Without any tuning the fp32 encoder call learns the pyramid once and runs mode 5; bits == mode 1 == oracle.
    (2) A second pyramid with the SAME S (H and W swapped) hits the value-keyed cache with the wrong geometry: the
    device-side guard must make the result that of the device tensors anyway, report the mismatch, and the cache must
    relearn.  (3) A wrong ``spatial_shapes_list`` is equally harmless."""
    from gomatching_b200 import _native
    import gomatching_b200 as g
    _native._window_geometry.clear()
    monkeypatch.setattr(_native, "WINDOW_MIN_ITEMS_PER_SM", 0)        # small test maps: take the window kernel anyway
    a = _encoder_workload(96, 160, 21)
    b = _encoder_workload(96, 160, 22, transpose=True)
    learned0 = _native.window_stats["learned"]
    out_a = run_core(a).cpu().numpy()                                  # auto -> mode 5
    assert _native.window_stats["learned"] == learned0 + 1
    assert np.array_equal(out_a, run_core(a, dict(mode=1)).cpu().numpy())
    assert np.array_equal(out_a, O.forward_f32(a["value"], a["shapes"], a["lsi"], a["loc"], a["attn"]))
    run_core(a)
    assert _native.window_stats["learned"] == learned0 + 1             # no second device->host read
    ep0 = _native.lib().msda_b200_shape_mismatch_epoch()
    out_b = run_core(b).cpu().numpy()                                  # cache says pyramid A: guard must catch it
    want_b = O.forward_f32(b["value"], b["shapes"], b["lsi"], b["loc"], b["attn"])
    assert np.array_equal(out_b, want_b), "stale host geometry changed the result"
    torch.cuda.synchronize()
    assert _native.lib().msda_b200_shape_mismatch_epoch() != ep0       # reported through pinned host memory
    out_b2 = run_core(b).cpu().numpy()                                 # cache dropped and relearned for B
    assert np.array_equal(out_b2, want_b) and _native.window_stats["learned"] == learned0 + 2
    ep1 = _native.lib().msda_b200_shape_mismatch_epoch()
    run_core(b)
    torch.cuda.synchronize()
    assert _native.lib().msda_b200_shape_mismatch_epoch() == ep1       # and no further mismatch
    # (3) lying list: same S, wrong geometry
    lie = [tuple(int(v) for v in hw) for hw in a["shapes"]]
    out = g.ms_deform_attn_forward(dev(b["value"]), dev(b["shapes"]), dev(b["lsi"]), dev(b["loc"]), dev(b["attn"]), 64,
                                   spatial_shapes_list=lie).cpu().numpy()
    assert np.array_equal(out, want_b)
    with pytest.raises(ValueError):
        g.ms_deform_attn_forward(dev(b["value"]), dev(b["shapes"]), dev(b["lsi"]), dev(b["loc"]), dev(b["attn"]), 64,
                                 spatial_shapes_list=[(1, 1)] * 4)
    _native._window_geometry.clear()


def test_window_kernel_default_under_cuda_graph_capture(monkeypatch):
    """Inside a capture an unknown pyramid cannot be learned (no device->host read): the call must fall back to the
    register-gather kernel and still be capturable; a known pyramid captures the window kernel."""
    from gomatching_b200 import _native
    import gomatching_b200 as g
    monkeypatch.setattr(_native, "WINDOW_MIN_ITEMS_PER_SM", 0)
    a = _encoder_workload(64, 96, 31)
    t = {k: dev(v) for k, v in a.items()}
    want = O.forward_f32(a["value"], a["shapes"], a["lsi"], a["loc"], a["attn"])
    for known in (False, True):
        _native._window_geometry.clear()
        if known:
            run_core(a)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            g.ms_deform_attn_forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["attn"], 64)    # warm-up outside capture
            torch.cuda.synchronize()
            if not known:
                _native._window_geometry.clear()
            with torch.cuda.graph(graph, stream=s):
                out = g.ms_deform_attn_forward(t["value"], t["shapes"], t["lsi"], t["loc"], t["attn"], 64)
        graph.replay()
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), want), known


def test_1080p_encoder_bit_exact_vs_oracle_and_reference_kernel():
    """BASELINE.json config 5 size (1080x1920, S = 43110): the default path (window kernel) against the CPU oracle and
    the unmodified reference CUDA kernel -- not only against itself."""
    from gomatching_b200 import synthetic as syn
    w = syn.make_workload("encoder", 1080, 1920, n=1, seed=5, dist="local")
    c = dict(value=w.value.numpy(), shapes=w.shapes.numpy(), lsi=w.lsi.numpy(), loc=w.loc.numpy(), attn=w.attn.numpy())
    out = run_core(c)
    ora = O.forward_f32(c["value"], c["shapes"], c["lsi"], c["loc"], c["attn"])
    assert np.array_equal(out.cpu().numpy(), ora)
    ref = _refcuda()
    if ref is not None:
        assert torch.equal(out, _run_refcuda(ref, w)), "differs from the reference CUDA kernel"
    wd = syn.make_workload("decoder", 1080, 1920, n=1, seed=6, dist="local")
    cd = dict(value=wd.value.numpy(), shapes=wd.shapes.numpy(), lsi=wd.lsi.numpy(), loc=wd.loc.numpy(), attn=wd.attn.numpy())
    assert np.array_equal(run_core(cd).cpu().numpy(), O.forward_f32(cd["value"], cd["shapes"], cd["lsi"], cd["loc"], cd["attn"]))


def test_tuning_variant_never_changes_results_in_the_product_build(core_cases):
    """the staged kernel's time-attribution builds (wrong results) are compiled only with -DMSDA_DIAG."""
    a = _encoder_workload(64, 96, 41)
    want = O.forward_f32(a["value"], a["shapes"], a["lsi"], a["loc"], a["attn"])
    for mode in (4, 5):
        for variant in range(0, 6):
            assert np.array_equal(run_core(a, dict(mode=mode, variant=variant)).cpu().numpy(), want), (mode, variant)


def test_backward_accepts_other_dtypes_and_returns_them():
    import gomatching_b200 as g
    from gomatching_b200 import synthetic as syn
    w = syn.make_workload("decoder", 96, 160, n=1, seed=2, dist="uniform")
    v, sh, ls = w.value.cuda(), w.shapes.cuda(), w.lsi.cuda()
    go = torch.randn(1, w.loc.shape[1], 256, device="cuda")
    gv, gl, ga = g.ms_deform_attn_backward(v, sh, ls, w.loc.cuda(), w.attn.cuda(), go, 64)
    gv2, gl2, ga2 = g.ms_deform_attn_backward(v, sh, ls, w.loc.cuda().double(), w.attn.cuda().double(), go.double(), 64)
    assert gl2.dtype == torch.float64 and ga2.dtype == torch.float64
    assert torch.allclose(gl2.float(), gl, rtol=1e-4, atol=1e-5) and torch.allclose(ga2.float(), ga, rtol=1e-4, atol=1e-5)
    assert torch.allclose(gv2, gv, rtol=1e-3, atol=1e-4)           # atomics: summation order differs run to run


def test_second_gpu_in_the_same_process():
    """kernel attributes are per device, and the projection wrappers must follow the tensor's device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import gomatching_b200 as g
    from gomatching_b200 import synthetic as syn
    w = syn.make_workload("encoder", 96, 160, n=1, seed=3, dist="local")
    outs = []
    for d in (0, 1):
        dv = torch.device("cuda", d)
        m = g.MSDeformAttn(256, 4, 8, 4).to(dv).eval()
        torch.manual_seed(0)
        for p_ in m.parameters():
            p_.data.normal_(0, 0.05)
        m.invalidate_caches()
        q = torch.randn(1, w.ref.shape[1], 256, generator=torch.Generator().manual_seed(1)).to(dv)
        src = torch.randn(1, w.value.shape[1], 256, generator=torch.Generator().manual_seed(2)).to(dv)
        with torch.no_grad():                                     # current device stays 0 for both
            outs.append(m(q, w.ref.to(dv), src, w.shapes.to(dv), w.lsi.to(dv)).cpu())
    assert torch.equal(outs[0], outs[1])
