"""Summarise an ncu launch list (gpu__time_duration.sum csv) by kernel family.
    python tools/launch_summary.py gpurun_out/launches.csv > profiles/<name>.txt"""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except Exception:
        continue
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    rows.append((r["Kernel Name"], us))
def family(n):
    n0 = n
    n = re.sub(r"<.*", "", n)
    n = n.replace("void ", "")
    for key, fam in (("msda_fwd", "msda sampler (this library)"), ("proj_gemm", "tcgen05 3xTF32 GEMM (this library)"),
                     ("add_layernorm", "add + LayerNorm (this library)"), ("frames_u8", "frame batcher (this library)"),
                     ("frame_batch", "frame batcher (this library)"), ("split_tf32", "TF32 weight split (this library)"),
                     ("cudnn", "cuDNN convolution"), ("implicit_convolve", "cuDNN convolution"), ("conv", "cuDNN convolution"),
                     ("sm100_xmma", "cuDNN / cuBLAS xmma"), ("sm90_xmma", "cuDNN / cuBLAS xmma"), ("cutlass", "cuBLAS / CUTLASS GEMM"),
                     ("gemm", "cuBLAS GEMM"), ("gemv", "cuBLAS GEMM"), ("fmha", "attention (torch SDPA)"), ("flash", "attention (torch SDPA)"),
                     ("softmax", "torch softmax"), ("layer_norm", "torch LayerNorm / GroupNorm"), ("group_norm", "torch LayerNorm / GroupNorm"),
                     ("RowwiseMoments", "torch LayerNorm / GroupNorm"), ("elementwise", "torch elementwise"), ("reduce", "torch reduce"),
                     ("index", "torch indexing"), ("sort", "torch sort / topk"), ("topk", "torch sort / topk"), ("nchw", "cuDNN layout"), ("nhwc", "cuDNN layout")):
        if key.lower() in n0.lower():
            return fam
    return "other: " + n[:60]
tot = sum(u for _, u in rows)
fam = defaultdict(lambda: [0.0, 0])
for n, u in rows:
    f = fam[family(n)]; f[0] += u; f[1] += 1
print("launches: %d   total device time: %.1f us (ncu: serialised, cold cache -- compare SHARES)" % (len(rows), tot))
for k, (u, c) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
    print("%8.1f us  %5.1f %%  %5d launches  %s" % (u, 100 * u / tot, c, k))
print("\ntop individual kernels:")
byname = defaultdict(lambda: [0.0, 0])
for n, u in rows:
    b = byname[re.sub(r"\(.*", "", n)[:110]]; b[0] += u; b[1] += 1
for k, (u, c) in sorted(byname.items(), key=lambda kv: -kv[1][0])[:25]:
    print("%8.1f us  %5.1f %%  %5d x  %s" % (u, 100 * u / tot, c, k))
