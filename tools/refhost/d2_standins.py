"""Stand-ins for the detectron2 / fvcore symbols the reference's video path imports.

The reference's GoMatching meta-architecture and LST-Matcher head (gomatching/modeling/meta_arch/gom_lstmatcher.py:7-24,
gomatching/modeling/roi_heads/lstmatcher.py:6-23, association_head.py:6-7) are written against detectron2 v0.6 and
fvcore, neither of which is installed (and neither is vendored by the reference).  This module provides exactly the
symbols those files import, with detectron2's documented behaviour, so that the reference's own model code runs
unmodified.  It is HOST scaffolding for tests and the clip benchmark -- nothing here is on the accelerated path,
and nothing here comes from the reference tree.

Behaviour that matters for identical results:
  * ``pairwise_iou`` is 0 (not NaN) where the intersection is empty -- degenerate boxes are common at default init.
  * ``Instances`` indexing applies the index to every field; ``set`` checks lengths.
  * ``ImageList.from_tensors`` zero-pads to the batch maximum and keeps the true sizes.
  * the backbone is torchvision's ResNet-50 with frozen batch norm, returning res3 / res4 / res5 at strides
    8 / 16 / 32 (STRIDE_IN_1X1: False == torchvision's v1.5 bottleneck).
"""
from __future__ import annotations

import functools
import inspect
import sys
import types
from typing import Any, Dict, List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------- config
class CfgNode(dict):
    """Attribute-style nested dict (the subset of yacs/detectron2 CfgNode the model code reads)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    @classmethod
    def from_dict(cls, d):
        out = cls()
        for k, v in d.items():
            out[k] = cls.from_dict(v) if isinstance(v, dict) else v
        return out

    def merge(self, d):
        for k, v in d.items():
            if isinstance(v, dict) and isinstance(self.get(k), CfgNode):
                self[k].merge(v)
            else:
                self[k] = CfgNode.from_dict(v) if isinstance(v, dict) else v
        return self


def _called_with_cfg(*args, **kwargs):
    if len(args) and isinstance(args[0], CfgNode):
        return True
    return isinstance(kwargs.get("cfg", None), CfgNode)


def configurable(init_func=None, *, from_config=None):
    """detectron2.config.configurable for ``__init__``: ``Cls(cfg, ...)`` becomes ``Cls(**Cls.from_config(cfg, ...))``."""
    assert init_func is not None and inspect.isfunction(init_func) and from_config is None

    @functools.wraps(init_func)
    def wrapped(self, *args, **kwargs):
        if _called_with_cfg(*args, **kwargs):
            explicit = type(self).from_config(*args, **kwargs)
            init_func(self, **explicit)
        else:
            init_func(self, *args, **kwargs)

    return wrapped


class Registry:
    def __init__(self, name):
        self._name = name
        self._map: Dict[str, Any] = {}

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self._map[o.__name__] = o
                return o
            return deco
        self._map[obj.__name__] = obj
        return obj

    def get(self, name):
        if name not in self._map:
            raise KeyError("No object named '%s' found in '%s' registry!" % (name, self._name))
        return self._map[name]


META_ARCH_REGISTRY = Registry("META_ARCH")
ROI_HEADS_REGISTRY = Registry("ROI_HEADS")
BACKBONE_REGISTRY = Registry("BACKBONE")


# ----------------------------------------------------------------------------------------------- structures
class Boxes:
    def __init__(self, tensor):
        if not isinstance(tensor, torch.Tensor):
            tensor = torch.as_tensor(tensor, dtype=torch.float32, device=torch.device("cpu"))
        else:
            tensor = tensor.to(torch.float32)
        if tensor.numel() == 0:
            tensor = tensor.reshape((-1, 4)).to(dtype=torch.float32)
        assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
        self.tensor = tensor

    def clone(self):
        return Boxes(self.tensor.clone())

    def to(self, device):
        return Boxes(self.tensor.to(device=device))

    def area(self):
        b = self.tensor
        return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])

    def __getitem__(self, item):
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        b = self.tensor[item]
        assert b.dim() == 2, "Indexing on Boxes with {} failed to return a matrix!".format(item)
        return Boxes(b)

    def __len__(self):
        return self.tensor.shape[0]

    def __repr__(self):
        return "Boxes(" + str(self.tensor) + ")"

    @property
    def device(self):
        return self.tensor.device

    def __iter__(self):
        yield from self.tensor


def pairwise_intersection(boxes1: Boxes, boxes2: Boxes) -> torch.Tensor:
    a, b = boxes1.tensor, boxes2.tensor
    wh = torch.min(a[:, None, 2:], b[:, 2:]) - torch.max(a[:, None, :2], b[:, :2])
    wh.clamp_(min=0)
    return wh.prod(dim=2)


def pairwise_iou(boxes1: Boxes, boxes2: Boxes) -> torch.Tensor:
    area1, area2 = boxes1.area(), boxes2.area()
    inter = pairwise_intersection(boxes1, boxes2)
    return torch.where(inter > 0, inter / (area1[:, None] + area2 - inter),
                       torch.zeros(1, dtype=inter.dtype, device=inter.device))


class Instances:
    def __init__(self, image_size: Tuple[int, int], **kwargs: Any):
        self._image_size = image_size
        self._fields: Dict[str, Any] = {}
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, name, val):
        if name.startswith("_"):
            super().__setattr__(name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name):
        if name == "_fields" or name not in self._fields:
            raise AttributeError("Cannot find field '{}' in the given Instances!".format(name))
        return self._fields[name]

    def set(self, name, value):
        data_len = len(value)
        if len(self._fields):
            assert len(self) == data_len, "Adding a field of length {} to a Instances of length {}".format(
                data_len, len(self))
        self._fields[name] = value

    def has(self, name):
        return name in self._fields

    def remove(self, name):
        del self._fields[name]

    def get(self, name):
        return self._fields[name]

    def get_fields(self):
        return self._fields

    def to(self, *args, **kwargs):
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            if hasattr(v, "to"):
                v = v.to(*args, **kwargs)
            ret.set(k, v)
        return ret

    def __getitem__(self, item):
        if type(item) == int:
            if item >= len(self) or item < -len(self):
                raise IndexError("Instances index out of range!")
            item = slice(item, None, len(self))
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v[item])
        return ret

    def __len__(self):
        for v in self._fields.values():
            return v.__len__()
        raise NotImplementedError("Empty Instances does not support __len__!")

    def __iter__(self):
        raise NotImplementedError("`Instances` object is not iterable!")

    def __repr__(self):
        return "Instances(num_instances=%d, image_size=%s, fields=[%s])" % (
            len(self) if self._fields else 0, self._image_size, ", ".join(self._fields))


class ImageList:
    def __init__(self, tensor, image_sizes):
        self.tensor = tensor
        self.image_sizes = image_sizes

    def __len__(self):
        return len(self.image_sizes)

    @property
    def device(self):
        return self.tensor.device

    def to(self, *a, **k):
        return ImageList(self.tensor.to(*a, **k), self.image_sizes)

    @staticmethod
    def from_tensors(tensors: List[torch.Tensor], size_divisibility: int = 0, pad_value: float = 0.0):
        assert len(tensors) > 0
        image_sizes = [(im.shape[-2], im.shape[-1]) for im in tensors]
        max_h = max(s[0] for s in image_sizes)
        max_w = max(s[1] for s in image_sizes)
        if size_divisibility > 1:
            st = size_divisibility
            max_h, max_w = (max_h + st - 1) // st * st, (max_w + st - 1) // st * st
        if len(tensors) == 1:
            h, w = image_sizes[0]
            batched = F.pad(tensors[0], [0, max_w - w, 0, max_h - h], value=pad_value).unsqueeze_(0)
        else:
            batched = tensors[0].new_full((len(tensors),) + tuple(tensors[0].shape[:-2]) + (max_h, max_w), pad_value)
            for img, pad_img in zip(tensors, batched):
                pad_img[..., : img.shape[-2], : img.shape[-1]].copy_(img)
        return ImageList(batched.contiguous(), image_sizes)


# ----------------------------------------------------------------------------------------------- layers
class ShapeSpec:
    def __init__(self, channels=None, height=None, width=None, stride=None):
        self.channels, self.height, self.width, self.stride = channels, height, width, stride


Linear = nn.Linear


def nms(boxes, scores, iou_threshold):
    import torchvision
    return torchvision.ops.nms(boxes.float(), scores.float(), iou_threshold)


class FrozenBatchNorm2d(nn.Module):
    """Batch norm with fixed statistics and affine parameters (detectron2.layers.FrozenBatchNorm2d)."""

    def __init__(self, num_features, eps=1e-5):
        super().__init__()
        self.num_features, self.eps = num_features, eps
        self.register_buffer("weight", torch.ones(num_features))
        self.register_buffer("bias", torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features) - eps)

    def forward(self, x):
        if x.requires_grad:
            scale = self.weight * (self.running_var + self.eps).rsqrt()
            bias = self.bias - self.running_mean * scale
            return x * scale.reshape(1, -1, 1, 1).to(x.dtype) + bias.reshape(1, -1, 1, 1).to(x.dtype)
        # detectron2's inference branch: one fused kernel, no intermediate tensors
        return F.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias, training=False, eps=self.eps)


class ResNet50Features(nn.Module):
    """build_resnet_backbone stand-in: DEPTH 50, OUT_FEATURES res3/res4/res5 (configs/GoMatching_ICDAR15.yaml:7-12)."""

    def __init__(self, cfg=None):
        super().__init__()
        import torchvision
        net = torchvision.models.resnet50(weights=None, norm_layer=FrozenBatchNorm2d)
        self.stem = nn.Sequential(net.conv1, net.bn1, net.relu, net.maxpool)
        self.res2, self.res3, self.res4, self.res5 = net.layer1, net.layer2, net.layer3, net.layer4

    def forward(self, x):
        x = self.res2(self.stem(x))
        out = {}
        x = self.res3(x); out["res3"] = x
        x = self.res4(x); out["res4"] = x
        x = self.res5(x); out["res5"] = x
        return out

    def output_shape(self):
        return {"res3": ShapeSpec(channels=512, stride=8), "res4": ShapeSpec(channels=1024, stride=16),
                "res5": ShapeSpec(channels=2048, stride=32)}


def build_backbone(cfg, input_shape=None):
    name = cfg.MODEL.BACKBONE.NAME
    if name != "build_resnet_backbone":
        raise NotImplementedError("stand-in backbone only for build_resnet_backbone, got %s" % name)
    return ResNet50Features(cfg)


def build_roi_heads(cfg, input_shape):
    return ROI_HEADS_REGISTRY.get(cfg.MODEL.ROI_HEADS.NAME)(cfg, input_shape)


def build_model(cfg):
    model = META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)(cfg)
    model.to(torch.device(cfg.MODEL.DEVICE))
    return model


class ROIHeads(nn.Module):
    pass


class Matcher:
    def __init__(self, thresholds, labels, allow_low_quality_matches=False):
        self.thresholds, self.labels, self.allow_low_quality_matches = thresholds, labels, allow_low_quality_matches

    def __call__(self, match_quality_matrix):   # training only
        raise NotImplementedError("Matcher is a training-time component; the video path never calls it")


def _training_only(*a, **k):
    raise NotImplementedError("training-time detectron2 utility; not on the video inference path")


class _Metadata(types.SimpleNamespace):
    def get(self, key, default=None):
        return getattr(self, key, default)


class _MetadataCatalog:
    def __init__(self):
        self._m: Dict[str, _Metadata] = {}

    def get(self, name):
        if name not in self._m:
            self._m[name] = _Metadata(name=name)
        return self._m[name]


MetadataCatalog = _MetadataCatalog()


# ----------------------------------------------------------------------------------------------- fvcore
def c2_xavier_fill(module: nn.Module) -> None:
    nn.init.kaiming_uniform_(module.weight, a=1)
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


def c2_msra_fill(module: nn.Module) -> None:
    nn.init.kaiming_normal_(module.weight, mode="fan_out", nonlinearity="relu")
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


# ----------------------------------------------------------------------------------------------- registration
def _mod(name, **symbols):
    m = types.ModuleType(name)
    m.__dict__.update(symbols)
    m.__standin__ = True
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


def install() -> None:
    """Register the stand-in modules under the names the reference imports.  A real detectron2, if importable,
    is left alone."""
    if "detectron2" in sys.modules and not getattr(sys.modules["detectron2"], "__standin__", False):
        return
    _mod("detectron2")
    _mod("detectron2.config", configurable=configurable, CfgNode=CfgNode)
    _mod("detectron2.structures", Boxes=Boxes, pairwise_iou=pairwise_iou, ImageList=ImageList, Instances=Instances)
    _mod("detectron2.layers", nms=nms, Linear=Linear, ShapeSpec=ShapeSpec, FrozenBatchNorm2d=FrozenBatchNorm2d)
    _mod("detectron2.data", MetadataCatalog=MetadataCatalog)
    _mod("detectron2.modeling", build_backbone=build_backbone, build_roi_heads=build_roi_heads,
         build_model=build_model)
    _mod("detectron2.modeling.meta_arch")
    _mod("detectron2.modeling.meta_arch.build", META_ARCH_REGISTRY=META_ARCH_REGISTRY, build_model=build_model)
    _mod("detectron2.modeling.proposal_generator")
    _mod("detectron2.modeling.proposal_generator.proposal_utils", add_ground_truth_to_proposals=_training_only)
    _mod("detectron2.modeling.matcher", Matcher=Matcher)
    _mod("detectron2.modeling.sampling", subsample_labels=_training_only)
    _mod("detectron2.modeling.roi_heads")
    _mod("detectron2.modeling.roi_heads.roi_heads", ROI_HEADS_REGISTRY=ROI_HEADS_REGISTRY, ROIHeads=ROIHeads)
    _mod("detectron2.utils")
    _mod("detectron2.utils.events", get_event_storage=_training_only)
    _mod("detectron2.utils.comm", get_world_size=lambda: 1)
    _mod("fvcore")
    _mod("fvcore.nn")
    _mod("fvcore.nn.weight_init", c2_xavier_fill=c2_xavier_fill, c2_msra_fill=c2_msra_fill)
