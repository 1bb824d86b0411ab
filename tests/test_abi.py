"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol
include/msda_b200.h declares (no compute calls: there is no GPU here and the library has no CPU path)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "msda_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msda_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from gomatching_b200 import _native, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = header_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), "header declares %s but the .so does not export it" % name
    assert sorted(_native.SYMBOLS) == declared, "gomatching_b200/_native.py SYMBOLS out of sync with the header"


def test_abi_version_and_error_strings():
    from gomatching_b200 import _native
    L = _native.lib()
    assert L.msda_b200_abi_version() == _native.ABI_VERSION == 4
    assert L.msda_b200_variant_count() >= 1
    assert b"NULL" in L.msda_b200_error_string(-1)
    assert b"2 or 4" in L.msda_b200_error_string(-4)
    assert b"no CPU path" in L.msda_b200_error_string(-6)
    assert L.msda_b200_error_string(0) == b"success"


def test_argument_errors_without_a_gpu():
    """Validation happens before any CUDA call, so it can be exercised on the CPU box."""
    from gomatching_b200 import _native
    L = _native.lib()
    z = ctypes.c_void_p(0)
    assert L.msda_b200_forward_f32(z, z, z, z, z, 1, 1, 1, 32, 1, 1, 1, z, z) == -1          # NULL pointers
    buf = (ctypes.c_char * 256)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert L.msda_b200_forward_f32(p, p, p, p, p, 0, 1, 1, 32, 1, 1, 1, p, z) == -2          # N = 0
    assert L.msda_b200_forward_f32(p, p, p, p, p, 1, 1 << 24, 8, 32, 1, 1, 1, p, z) == -2    # value map >= 2 GiB
    assert L.msda_b200_forward_fused_f32(p, p, p, p, 3, p, p, 1, 1, 1, 32, 1, 1, 1, p, z, None) == -4   # ref_dim 3


def test_operator_raises_on_cpu_tensors_like_the_reference():
    """ms_deform_attn.h:38 AT_ERROR("Not implemented on the CPU") -> RuntimeError; no CPU fallback."""
    import gomatching_b200 as g
    v = torch.zeros(1, 4, 1, 32)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        g.ms_deform_attn_forward(v, torch.tensor([[2, 2]]), torch.tensor([0]), torch.zeros(1, 1, 1, 1, 4, 2),
                                 torch.zeros(1, 1, 1, 1, 4), 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        g.MSDeformAttnFunction.apply(v, torch.tensor([[2, 2]]), torch.tensor([0]), torch.zeros(1, 1, 1, 1, 4, 2),
                                     torch.zeros(1, 1, 1, 1, 4), 64)


def test_module_is_state_dict_compatible_with_the_reference(module_cases):
    """Parameter names/shapes of ms_deform_attn.py:94-97 -- a reference state dict loads strictly."""
    import gomatching_b200 as g
    for name in module_cases.names():
        c = module_cases.case(name)
        d_model, levels, heads, points = (int(v) for v in c["cfg"])
        mod = g.MSDeformAttn(d_model, levels, heads, points)
        sd = {k[3:]: torch.from_numpy(v) for k, v in c.items() if k.startswith("sd/")}
        assert set(sd) == set(mod.state_dict())
        mod.load_state_dict(sd, strict=True)
        assert mod.im2col_step == 64


def test_module_constructor_errors_and_default_init():
    import gomatching_b200 as g
    with pytest.raises(ValueError, match="d_model must be divisible by n_heads"):
        g.MSDeformAttn(250, 4, 8, 4)
    m = g.MSDeformAttn(256, 4, 8, 4)
    # ms_deform_attn.py:101-115: zero offset weights, compass-rose bias, uniform attention
    assert float(m.sampling_offsets.weight.abs().max()) == 0.0
    b = m.sampling_offsets.bias.view(8, 4, 4, 2)
    assert torch.allclose(b[0, :, :, 0], torch.tensor([1., 2., 3., 4.]).expand(4, 4))
    assert float(m.attention_weights.bias.abs().max()) == 0.0
    with pytest.raises(ValueError, match="must be 2 or 4"):
        m(torch.zeros(1, 3, 256), torch.zeros(1, 3, 4, 3), torch.zeros(1, 5, 256), torch.tensor([[1, 5]]),
          torch.tensor([0]))


def test_synthetic_shapes_match_the_survey_table():
    from gomatching_b200 import synthetic as syn
    assert syn.level_shapes(720, 1280) == [(90, 160), (45, 80), (23, 40), (12, 20)]
    assert syn.level_shapes(1080, 1920) == [(135, 240), (68, 120), (34, 60), (17, 30)]
    assert syn.level_start_index(syn.level_shapes(720, 1280)).tolist() == [0, 14400, 18000, 18920]
    w = syn.make_workload("decoder", 96, 160, seed=1)
    assert w.dims == (1, 12 * 20 + 6 * 10 + 3 * 5 + 2 * 3, 8, 32, 4, 2500, 4)
    enc = syn.make_workload("encoder", 96, 160, seed=1)
    N, S, M, D, L, Lq, P = enc.dims
    assert Lq == S
    assert enc.algorithmic_bytes() == 4 * S * 256 + 12 * S * 8 * 16 + 4 * S * 256


def test_bench_figures_match_the_survey():
    """bench.py's algorithmic bytes per launch are SURVEY s8d's B_alg: 68.67 MB per 720p encoder call, 26.02 MB per
    fp32 decoder call (value counted once, capped by the gather volume)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.algorithmic_bytes(1, 19160, 19160) == 68669440
    assert bench.algorithmic_bytes(8, 19160, 19160) == 549355520
    assert bench.algorithmic_bytes(1, 19160, 2500) == 19619840 + 12 * 2500 * 8 * 16 + 4 * 2500 * 256
    assert abs(bench.algorithmic_bytes(1, 19160, 2500) / 1e6 - 26.02) < 0.01
    assert (bench.ENC_LAYERS, bench.DEC_LAYERS, bench.HEIGHT, bench.WIDTH) == (6, 6, 720, 1280)
