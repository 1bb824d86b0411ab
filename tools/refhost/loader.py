"""Run the reference's OWN video model (GoMatching = frozen DeepSolo spotter + LST-Matcher) without detectron2.

What this is: host scaffolding that lets the unmodified reference Python -- gomatching/modeling/meta_arch/
gom_lstmatcher.py (GoMatching), gomatching/modeling/roi_heads/lstmatcher.py (LSTMatcher), third_party/adet/... --
be imported and driven from tests and from ``bench.py``'s clip workload, in the role the reference's ``eval.py``
plays.  The B200 kernels enter through ``gomatching_b200.install_into_adet()`` exactly as they would in a real
deployment; the tracker is the reference's code, untouched (north_star: "left unchanged").

Where the reference comes from (first hit wins):
  1. ``$GOM_REFERENCE``
  2. ``/root/reference``            (the build container)
  3. ``<repo>/baseline/_ref``       (a file-for-file copy of the reference's Python made by ``stage()`` during
                                    ``__graft_entry__.build()``; git-ignored, travels to the GPU box like the
                                    built ``.so`` files.  Nothing in it is edited.)

Import recipe: the packages' ``__init__`` files pull in detectron2 and the dataset registry, so
``adet``, ``adet.layers``, ..., ``gomatching``, ``gomatching.modeling``, ... are registered as bare namespace modules
pointing at the reference directories, ``adet._C`` is a stub, and the detectron2 / fvcore names resolve to
``d2_standins``.  ``adet._C.ms_deform_attn_forward`` defaults to the reference's own
``ms_deform_attn_core_pytorch`` (BASELINE.json config 1: "MSDeformAttn via ms_deform_attn_core_pytorch"); the
reference CUDA op itself raises on CPU tensors (csrc/DeformAttn/ms_deform_attn.h:38).
"""
from __future__ import annotations

import importlib
import os
import shutil
import sys
import types
from typing import Optional

import torch

from . import d2_standins as D2

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
STAGED = os.path.join(REPO, "baseline", "_ref")

# what stage() copies: the reference's Python on the video path + its configs (no csrc, no weights, no data)
_STAGE_DIRS = ("gomatching", os.path.join("third_party", "adet"), "configs")


def reference_root() -> Optional[str]:
    for cand in (os.environ.get("GOM_REFERENCE"), "/root/reference", STAGED):
        if cand and os.path.isfile(os.path.join(cand, "gomatching", "modeling", "meta_arch", "gom_lstmatcher.py")):
            return cand
    return None


def stage(src: str = "/root/reference", dst: str = STAGED) -> Optional[str]:
    """Copy the reference's Python (unmodified) to ``baseline/_ref`` so it can travel to the GPU box.  No-op when the
    source tree is absent (the GPU box) -- the previously staged copy is then used as is."""
    if not os.path.isdir(os.path.join(src, "gomatching")):
        return dst if os.path.isdir(os.path.join(dst, "gomatching")) else None
    for sub in _STAGE_DIRS:
        for root, dirs, files in os.walk(os.path.join(src, sub)):
            dirs[:] = [d for d in dirs if d not in ("csrc", "__pycache__")]
            rel = os.path.relpath(root, src)
            for f in files:
                if f.endswith((".py", ".yaml")):
                    os.makedirs(os.path.join(dst, rel), exist_ok=True)
                    shutil.copyfile(os.path.join(root, f), os.path.join(dst, rel, f))
    return dst


_loaded = {}


def _ns(name: str, path: str):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


def load_reference(root: Optional[str] = None):
    """Import the reference's model modules.  Returns a namespace with ``msda``, ``dt`` (deformable_transformer),
    ``meta`` (gom_lstmatcher), ``roi`` (lstmatcher) and ``root``.  Idempotent."""
    root = root or reference_root()
    if root is None:
        raise FileNotFoundError("reference tree not found (looked at $GOM_REFERENCE, /root/reference, %s)" % STAGED)
    if _loaded.get("root") == root:
        return _loaded["ns"]
    D2.install()
    adet = os.path.join(root, "third_party", "adet")
    gom = os.path.join(root, "gomatching")
    for name, path in (("adet", adet), ("adet.layers", os.path.join(adet, "layers")),
                       ("adet.utils", os.path.join(adet, "utils")), ("adet.modeling", os.path.join(adet, "modeling")),
                       ("adet.modeling.model", os.path.join(adet, "modeling", "model")),
                       ("adet.config", os.path.join(adet, "config")),
                       ("gomatching", gom), ("gomatching.modeling", os.path.join(gom, "modeling")),
                       ("gomatching.modeling.meta_arch", os.path.join(gom, "modeling", "meta_arch")),
                       ("gomatching.modeling.roi_heads", os.path.join(gom, "modeling", "roi_heads"))):
        _ns(name, path)
    stub = types.ModuleType("adet._C")
    sys.modules["adet._C"] = stub
    sys.modules["adet"]._C = stub
    msda = importlib.import_module("adet.layers.ms_deform_attn")
    use_reference_cpu_path(stub, msda)
    dt = importlib.import_module("adet.layers.deformable_transformer")
    roi = importlib.import_module("gomatching.modeling.roi_heads.lstmatcher")
    importlib.import_module("gomatching.modeling.roi_heads.shared_ffn_crsattn")     # registers SHA_FFN_CRSATTN (GoMatching++ configs)
    meta = importlib.import_module("gomatching.modeling.meta_arch.gom_lstmatcher")
    ns = types.SimpleNamespace(root=root, msda=msda, dt=dt, roi=roi, meta=meta, C=stub,
                               original=dict(MSDeformAttn=msda.MSDeformAttn,
                                             EncoderLayer=dt.DeformableTransformerEncoderLayer,
                                             DecoderLayer=dt.DeformableCompositeTransformerDecoderLayer,
                                             Transformer=dt.DeformableTransformer,
                                             Encoder=dt.DeformableTransformerEncoder))
    _loaded.update(root=root, ns=ns)
    return ns


def use_reference_cpu_path(stub=None, msda=None):
    """Route ``adet._C.ms_deform_attn_forward`` to the reference's ms_deform_attn_core_pytorch
    (third_party/adet/layers/ms_deform_attn.py:40-60) -- works on any device, pure torch."""
    stub = stub or sys.modules["adet._C"]
    msda = msda or sys.modules["adet.layers.ms_deform_attn"]

    def fwd(value, shapes, lsi, loc, attn, im2col_step):
        return msda.ms_deform_attn_core_pytorch(value, shapes.tolist(), loc, attn)

    def bwd(*a, **k):
        raise NotImplementedError("reference CPU route is forward-only")

    stub.ms_deform_attn_forward = fwd
    stub.ms_deform_attn_backward = bwd


def use_reference_cuda_kernel(lib_path: Optional[str] = None) -> bool:
    """Route ``adet._C.ms_deform_attn_forward`` to the UNMODIFIED reference CUDA kernel (ms_deform_im2col_cuda.cuh compiled
    where it lies by oracle/Makefile -> oracle/_ref/libmsda_refcuda.so), with the reference host wrapper's behaviour
    (ms_deform_attn_cuda.cu:20-80: contiguous CUDA fp32 tensors, zero-initialised output).  Checker / baseline use only.
    False if the library was not built."""
    import ctypes

    lib_path = lib_path or os.path.join(REPO, "oracle", "_ref", "libmsda_refcuda.so")
    if not os.path.exists(lib_path):
        return False
    lib = ctypes.CDLL(lib_path)
    lib.refcuda_msda_forward_f32.restype = ctypes.c_int
    lib.refcuda_msda_forward_f32.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 7 + [ctypes.c_void_p] * 2

    def fwd(value, shapes, lsi, loc, attn, im2col_step):
        assert value.is_cuda and value.dtype == torch.float32
        value, loc, attn = value.contiguous(), loc.contiguous(), attn.contiguous()
        N, S, M, D = value.shape
        Lq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
        out = torch.zeros(N, Lq, M * D, device=value.device)
        with torch.cuda.device(value.device):
            rc = lib.refcuda_msda_forward_f32(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(),
                                              attn.data_ptr(), N, S, M, D, L, Lq, P, out.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        return out

    sys.modules["adet._C"].ms_deform_attn_forward = fwd
    return True


def restore_reference_classes():
    """Undo ``gomatching_b200.install_into_adet()``'s class swaps (the reference's own modules again)."""
    ns = load_reference()
    ns.msda.MSDeformAttn = ns.original["MSDeformAttn"]
    ns.dt.MSDeformAttn = ns.original["MSDeformAttn"]
    ns.dt.DeformableTransformerEncoderLayer = ns.original["EncoderLayer"]
    ns.dt.DeformableCompositeTransformerDecoderLayer = ns.original["DecoderLayer"]
    ns.dt.DeformableTransformer = ns.original["Transformer"]
    ns.dt.DeformableTransformerEncoder = ns.original["Encoder"]
    wob = sys.modules.get("adet.modeling.model.detection_transformer_wobackbone")
    if wob is not None:
        wob.DeformableTransformer = ns.original["Transformer"]
    use_reference_cpu_path()


# ----------------------------------------------------------------------------------------------- config
def _d2_base_cfg():
    """The detectron2 v0.6 default keys the reference's config files extend or its model code reads."""
    CN = D2.CfgNode
    c = CN()
    c.MODEL = CN(DEVICE="cuda", META_ARCHITECTURE="GeneralizedRCNN", WEIGHTS="", PIXEL_MEAN=[103.53, 116.28, 123.675],
                 PIXEL_STD=[1.0, 1.0, 1.0], MASK_ON=False, KEYPOINT_ON=False)
    c.MODEL.BACKBONE = CN(NAME="build_resnet_backbone", FREEZE_AT=2)
    c.MODEL.RESNETS = CN(DEPTH=50, OUT_FEATURES=["res4"], STRIDE_IN_1X1=True, NORM="FrozenBN")
    c.MODEL.ROI_HEADS = CN(NAME="Res5ROIHeads", NUM_CLASSES=80, IN_FEATURES=["res4"], IOU_THRESHOLDS=[0.5],
                           IOU_LABELS=[0, 1], BATCH_SIZE_PER_IMAGE=512, POSITIVE_FRACTION=0.25,
                           PROPOSAL_APPEND_GT=True)
    c.INPUT = CN(MIN_SIZE_TEST=800, MAX_SIZE_TEST=1333, FORMAT="BGR")
    c.INPUT.CROP = CN(ENABLED=False)
    c.DATASETS = CN(TRAIN=(), TEST=())
    c.DATALOADER = CN(SAMPLER_TRAIN="TrainingSampler")
    c.SOLVER = CN()
    c.SOLVER.CLIP_GRADIENTS = CN(ENABLED=False)
    c.TEST = CN()
    c.OUTPUT_DIR = "./output"
    return c


def build_cfg(config: str = "GoMatching_ICDAR15", device: str = "cpu", **overrides):
    """cfg = detectron2 base keys + the reference's adet/config/defaults.py + gomatching/config.py add_gom_config +
    configs/<config>.yaml + eval.py:220's ASSO_THRESH_TEST rule.  ``overrides`` use dotted keys with ``__`` for dots,
    e.g. ``MODEL__TRANSFORMER__ENC_LAYERS=2``."""
    import yaml

    ns = load_reference()
    base = _d2_base_cfg()
    defaults_mod = types.ModuleType("detectron2.config.defaults")
    defaults_mod._C = base
    sys.modules["detectron2.config.defaults"] = defaults_mod
    sys.modules["detectron2.config"].defaults = defaults_mod
    sys.modules.pop("adet.config.defaults", None)
    importlib.import_module("adet.config.defaults")                # extends ``base`` in place (adet/config/defaults.py)
    gom_cfg = importlib.import_module("gomatching.config")
    gom_cfg.add_gom_config(base)                                   # gomatching/config.py:3-80
    with open(os.path.join(ns.root, "configs", config + ".yaml")) as f:
        base.merge(yaml.safe_load(f))
    base.MODEL.DEVICE = device
    for k, v in overrides.items():
        node = base
        parts = k.split("__")
        for p in parts[:-1]:
            node = node[p]
        node[parts[-1]] = v
    base.MODEL.ASSO_HEAD.ASSO_THRESH_TEST = base.MODEL.TRANSFORMER.INFERENCE_TH_TEST      # eval.py:220
    return base


# ----------------------------------------------------------------------------------------------- model
def build_gomatching(cfg, seed: int = 0, b200=False, state_dict=None):
    """Instantiate the reference's GoMatching with its own initialisers (seeded).  ``b200`` (True = "layers", or
    "op" / "module" / "transformer" / "heads") installs the B200 operator / module / layers first (gomatching_b200.install_into_adet) so
    DeepSolo's encoder and decoder are built from them; pass ``state_dict`` of a reference-built model to get identical weights."""
    load_reference()
    restore_reference_classes()
    if b200:
        import gomatching_b200
        gomatching_b200.install_into_adet(level="layers" if b200 is True else b200)
    torch.manual_seed(seed)
    model = D2.build_model(cfg)
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    model.eval()
    for p in model.parameters():
        p.requires_grad_(False)
    if b200 == "heads":
        gomatching_b200.accelerate_spotter(model)
    return model


@torch.no_grad()
def calibrate_detections(model, frame_input, keep: int = 40) -> float:
    """Pick the score threshold that lets about ``keep`` of the queries through on ``frame_input`` and install it the
    way eval.py:219-220 does (INFERENCE_TH_TEST, and ASSO_THRESH_TEST tied to it).  A default-initialised model has
    no sharp scores: at the config's 0.3 all 100 queries pass and the tracker sees 100 detections per frame, 5-10x a
    real text video; SURVEY.md s8d asks for 20-60.  Returns the threshold."""
    images = model.preprocess_image([frame_input])
    features, pos = model.backbone(images)
    out = model.detection_transformer(features, pos, model.backbone)
    prob = out["pred_logits"].mean(-2).sigmoid()                       # gom_lstmatcher.py:592-601
    scores = prob.max(-1)[0]
    if model.with_rescore:
        re = model.roi_heads.rescoring_head(out["query_features"]).mean(-2).sigmoid().max(-1)[0]
        scores = torch.where(scores > re, scores, re)
    s = scores.flatten().sort(descending=True)[0]
    keep = max(1, min(int(keep), s.numel() - 1))
    thr = float((s[keep - 1] + s[keep]) / 2)
    model.test_score_threshold = thr
    model.roi_heads.asso_thresh_test = thr
    return thr


def synthetic_clip(n_frames: int, height: int, width: int, seed: int = 0):
    """Seeded uint8 BGR frames (H, W, 3) with smooth structure that drifts from frame to frame, as numpy arrays --
    the format ``read_image(path, format="BGR")`` hands to the reference's predictor (eval.py:326)."""
    import numpy as np

    g = torch.Generator().manual_seed(seed)
    base = torch.rand(1, 3, height // 8 + 2, width // 8 + 8, generator=g)
    big = torch.nn.functional.interpolate(base, scale_factor=8, mode="bilinear", align_corners=False)[0]
    noise = torch.rand(n_frames, 3, height, width, generator=g) * 0.1
    frames = []
    for t in range(n_frames):
        shift = (2 * t) % 56
        img = big[:, :height, shift:shift + width] * 0.9 + noise[t]
        frames.append((img.clamp(0, 1) * 255).to(torch.uint8).permute(1, 2, 0).contiguous().numpy())
    return frames


def frames_to_inputs(frames_bgr, input_format: str = "RGB"):
    """GoMBatchPredictor.__call__'s host preprocessing (text_track_visualizer.py:313-324) without the resize
    (frames are produced at the test size): BGR->RGB flip, HWC uint8 -> CHW float32, dict per frame."""
    out = []
    h, w = frames_bgr[0].shape[:2]
    for x in frames_bgr:
        if input_format == "RGB":
            x = x[:, :, ::-1]
        out.append({"image": torch.as_tensor(x.astype("float32").transpose(2, 0, 1)), "height": h, "width": w,
                    "video_id": 0})
    return out


def new_time_cost():
    return {'total_time': 0, 'pre_process': 0, 'backbone': 0, 'detector': 0, 'rescore': 0, 'tracker': 0,
            'long_match': 0, 'short_match': 0, 'post_process': 0}      # eval.py:299-300
