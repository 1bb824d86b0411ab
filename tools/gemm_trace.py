"""Per-CTA clock64 accounting of the persistent 3xTF32 GEMM (MSDA trace buffer): where the MMA issuer and the epilogue
wait.  python tools/gemm_trace.py [M N K] ; MSDA_GEMM_BN picks the tile width; MSDA_GEMM_DIAG needs a build.py --diag library."""
import sys, os, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gomatching_b200 import _native
from gomatching_b200.projections import linear_3xtf32
dev = torch.device("cuda", 0)
M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (19160 * 8, 256, 256)
w = torch.randn(N, K, device=dev) / 16; b = torch.randn(N, device=dev)
x = torch.randn(M, K, device=dev); out = torch.empty(M, N, device=dev)
for _ in range(3):
    linear_3xtf32(x, w, b, out=out)
buf = torch.zeros(4096, 8, dtype=torch.int64, device=dev)
L = _native.lib()
L.msda_b200_linear_set_trace(buf.data_ptr())
linear_3xtf32(x, w, b, out=out)
torch.cuda.synchronize()
L.msda_b200_linear_set_trace(None)
t = buf.cpu()
n = int((t[:, 0] != 0).sum())
t = t[:n].double()
print("M %d N %d K %d cfg %s: CTAs traced %d" % (M, N, K, os.environ.get("MSDA_GEMM_BN", "auto"), n))
rows = [("MMA warp lifetime (start -> last issue)", t[:, 4] - t[:, 0]),
        ("  of which waiting for TMA (full)", t[:, 1]),
        ("  of which waiting for the splitter", t[:, 2]),
        ("  of which waiting for a free accumulator", t[:, 3]),
        ("epilogue warp lifetime (start -> done)", t[:, 5] - t[:, 0]),
        ("  of which waiting for an accumulator", t[:, 6]),
        ("  of which draining", t[:, 7])]
for nm, v in rows:
    print("%-44s mean %9.0f  min %9.0f  max %9.0f clk" % (nm, float(v.mean()), float(v.min()), float(v.max())))
