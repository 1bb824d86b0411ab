"""GPU diagnostic: how far do the spotter's outputs move when the B200 module / layers replace the reference's?
    python tools/clip_diff.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import clip_common as C

def run(model, frame):
    with torch.no_grad():
        images = model.preprocess_image([frame])
        features, pos = model.backbone(images)
        out = model.detection_transformer(features, pos, model.backbone)
        out["re"] = model.roi_heads.rescoring_head(out["query_features"])
    return {k: v.double() for k, v in out.items() if v is not None}

cfg = C.L.build_cfg(device="cuda")
frames = C.L.frames_to_inputs(C.L.synthetic_clip(2, 720, 1280, seed=1))
ref_model = C.L.build_gomatching(cfg, seed=0)
C.use_reference_cuda_kernel()
sd = {k: v.clone() for k, v in ref_model.state_dict().items()}
ref = run(ref_model, frames[0])
ref2 = run(ref_model, frames[0])
print("reference run-to-run:", {k: float((ref[k] - ref2[k]).abs().max()) for k in ref})
for level in ("op", "module", "layers"):
    m = C.L.build_gomatching(cfg, seed=0, b200=level, state_dict=sd)
    got = run(m, frames[0])
    print(level, {k: "%.3g (rel %.3g)" % (float((ref[k] - got[k]).abs().max()), float((ref[k] - got[k]).abs().max() / ref[k].abs().max())) for k in ref})
    sc = got["pred_logits"].mean(-2).sigmoid().flatten(); rs = got["re"].mean(-2).sigmoid().flatten()
    fin = torch.maximum(sc, rs)
    print("   scores: min %.4f max %.4f; within 1e-4 of 0.3: %d; passing: %d" % (float(fin.min()), float(fin.max()), int(((fin - 0.3).abs() < 1e-4).sum()), int((fin > 0.3).sum())))
