"""The host-free ``DeformableTransformer`` drop-in (gomatching_b200/transformer_dropin.py) against the reference's own
forward: same tensor operations with Python-int loop bounds, so every output must be BIT-identical.

CPU: only the transformer / encoder containers are swapped (the reference's layers and its CPU operator path stay), on
the real GoMatching model.  GPU (tests/test_clip_gpu.py) exercises install level "transformer" end to end."""
import sys

import numpy as np
import pytest
import torch

import clip_common as C

pytestmark = pytest.mark.skipif(not C.have_reference(), reason="reference tree not available")


def _swap_transformer_only():
    from gomatching_b200.transformer_dropin import make_dropin_classes
    ns = C.L.load_reference()
    C.L.restore_reference_classes()
    new_t, new_e = make_dropin_classes(ns.dt)
    ns.dt.DeformableTransformer, ns.dt.DeformableTransformerEncoder = new_t, new_e
    sys.modules["adet.modeling.model.detection_transformer_wobackbone"].DeformableTransformer = new_t
    return new_t, new_e


def test_host_free_transformer_is_bit_identical_on_the_real_model():
    torch.set_num_threads(2)
    cfg = C.small_cfg(enc=2, dec=2)
    ref_model = C.L.build_gomatching(cfg, seed=0)
    sd = ref_model.state_dict()
    frames = C.L.synthetic_clip(5, 128, 192, seed=2)
    inputs = C.L.frames_to_inputs(frames)
    with torch.no_grad():
        images = ref_model.preprocess_image(inputs[:2])                      # N = 2: ragged valid ratios are exercised below
        feats, pos = ref_model.backbone(images)
        want = ref_model.detection_transformer(feats, pos, ref_model.backbone)
    ref, id_count = C.reference_loop(ref_model, frames)
    try:
        new_t, new_e = _swap_transformer_only()
        from tools.refhost import d2_standins as D2
        torch.manual_seed(0)
        model = D2.build_model(cfg)
        model.load_state_dict(sd, strict=True)                               # inherited constructor: same state-dict keys
        model.eval()
        tr = model.detection_transformer.transformer
        assert isinstance(tr, new_t) and isinstance(tr.encoder, new_e)
        with torch.no_grad():
            images = model.preprocess_image(inputs[:2])
            feats, pos = model.backbone(images)
            got = model.detection_transformer(feats, pos, model.backbone)
        for k in want:
            if want[k] is not None:
                assert torch.equal(want[k], got[k]), k
        out, id_count2 = C.reference_loop(model, frames)
        C.assert_identical(C.summarize(ref), C.summarize(out), "host-free transformer")
        assert id_count2 == id_count
    finally:
        C.L.restore_reference_classes()


def test_host_free_transformer_with_padding_masks():
    """Two frames of different sizes in one forward: padding masks and valid ratios differ per image."""
    torch.set_num_threads(2)
    cfg = C.small_cfg(enc=1, dec=1)
    ref_model = C.L.build_gomatching(cfg, seed=0)
    sd = ref_model.state_dict()
    a = C.L.frames_to_inputs(C.L.synthetic_clip(1, 128, 192, seed=3))[0]
    b = C.L.frames_to_inputs(C.L.synthetic_clip(1, 96, 160, seed=4))[0]

    def run(model):
        with torch.no_grad():
            images = model.preprocess_image([a, b])
            feats, pos = model.backbone(images)
            return model.detection_transformer(feats, pos, model.backbone)
    want = run(ref_model)
    try:
        _swap_transformer_only()
        from tools.refhost import d2_standins as D2
        torch.manual_seed(0)
        model = D2.build_model(cfg)
        model.load_state_dict(sd, strict=True)
        model.eval()
        got = run(model)
        for k in want:
            if want[k] is not None:
                assert torch.equal(want[k], got[k]), k
    finally:
        C.L.restore_reference_classes()


def test_tensor_core_linears_keep_parameters_and_cpu_behaviour():
    """Level "heads" re-classes the spotter's nn.Linear layers in place: same parameters and state-dict keys, the
    tracker's association head untouched, and anything the kernel does not take (here: CPU tensors) is nn.Linear itself,
    so the whole clip still reproduces the reference bit for bit on the host."""
    import gomatching_b200
    from gomatching_b200 import TensorCoreLinear
    torch.set_num_threads(2)
    cfg = C.small_cfg(enc=1, dec=1)
    model = C.L.build_gomatching(cfg, seed=0)
    keys = list(model.state_dict().keys())
    frames = C.L.synthetic_clip(4, 128, 192, seed=3)
    ref, id_count = C.reference_loop(model, frames)
    n = gomatching_b200.accelerate_spotter(model)
    assert n >= 10 and n == sum(isinstance(m, TensorCoreLinear) for m in model.modules())
    assert list(model.state_dict().keys()) == keys
    assert not any(isinstance(m, TensorCoreLinear) for m in model.roi_heads.asso_head.modules())
    assert gomatching_b200.accelerate_spotter(model) == 0                       # idempotent
    out, id_count2 = C.reference_loop(model, frames)
    C.assert_identical(C.summarize(ref), C.summarize(out), "TensorCoreLinear on the host")
    assert id_count2 == id_count
    lin = TensorCoreLinear(48, 64)
    x = torch.randn(5, 48, requires_grad=True)
    lin(x).sum().backward()                                                     # gradients: the nn.Linear path
    assert x.grad is not None and lin.weight.grad is not None
    with pytest.raises(ValueError):
        gomatching_b200.install_into_adet(level="everything")
