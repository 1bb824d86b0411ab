"""Small launches of every shared-memory / mbarrier / TMA kernel, for compute-sanitizer:
    compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_run.py
    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool synccheck python tools/sanitize_run.py
Every result is also compared with the register-gather kernel (mode 1), so a sanitizer run is a parity run too."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gomatching_b200 as g  # noqa: E402
from gomatching_b200 import synthetic as syn  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
w = syn.make_workload("encoder", 128, 192, n=2, seed=1, dist="uniform")
t = dict(v=w.value.cuda(), sh=w.shapes.cuda(), ls=w.lsi.cuda(), loc=w.loc.cuda(), attn=w.attn.cuda(), ref=w.ref.cuda(),
         off=w.offsets.cuda(), lg=w.logits.cuda())
base = g.ms_deform_attn_forward(t["v"], t["sh"], t["ls"], t["loc"], t["attn"], 64, tuning=dict(mode=1))
base_f = g.ms_deform_attn_forward_fused(t["v"], t["sh"], t["ls"], t["ref"], t["off"], t["lg"], tuning=dict(mode=1))
ran = []
for mode in (2, 4, 5):
    if which not in ("all", "mode%d" % mode):
        continue
    o = g.ms_deform_attn_forward(t["v"], t["sh"], t["ls"], t["loc"], t["attn"], 64, tuning=dict(mode=mode))
    of = g.ms_deform_attn_forward_fused(t["v"], t["sh"], t["ls"], t["ref"], t["off"], t["lg"], tuning=dict(mode=mode))
    torch.cuda.synchronize()
    assert torch.equal(o, base) and torch.equal(of, base_f), mode
    ran.append("mode %d core+fused" % mode)
if which in ("all", "paired"):
    p = g.pair_value_bf16(t["v"], t["sh"], t["ls"])
    o = g.ms_deform_attn_forward_paired(p, t["sh"], t["ls"], t["loc"], t["attn"])
    of = g.ms_deform_attn_forward_fused_paired(p, t["sh"], t["ls"], t["ref"], t["off"], t["lg"])
    torch.cuda.synchronize()
    assert float((o.float() - base).abs().max() / base.abs().max()) < 2e-2
    ran.append("paired bf16")
if which in ("all", "gemm"):
    from gomatching_b200.projections import linear_3xtf32
    x = torch.randn(1000, 256, device="cuda")
    wt = torch.randn(384, 256, device="cuda") * 0.05
    y = linear_3xtf32(x, wt, torch.zeros(384, device="cuda"))
    torch.cuda.synchronize()
    assert float((y - x.double().matmul(wt.double().t()).float()).abs().max()) < 1e-3
    ran.append("tcgen05 3xTF32 GEMM")
if which in ("all", "backward"):
    go = torch.randn_like(base)
    g.ms_deform_attn_backward(t["v"], t["sh"], t["ls"], t["loc"], t["attn"], go, 64)
    torch.cuda.synchronize()
    ran.append("backward")
print("sanitize_run ok:", "; ".join(ran))
