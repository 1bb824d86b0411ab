import sys, os, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gomatching_b200 import _native
from gomatching_b200.projections import linear_3xtf32
dev = torch.device("cuda", 0)
M, N, K = int(sys.argv[1]) if len(sys.argv) > 1 else 19160 * 8, 256, 256
w = torch.randn(N, K, device=dev) / 16; b = torch.randn(N, device=dev)
x = torch.randn(M, K, device=dev); out = torch.empty(M, N, device=dev)
for _ in range(3):
    linear_3xtf32(x, w, b, out=out)
buf = torch.zeros(4096, 8, dtype=torch.int64, device=dev)
L = _native.lib()
L.msda_b200_linear_set_trace.argtypes = [ctypes.c_void_p]
L.msda_b200_linear_set_trace.restype = None
L.msda_b200_linear_set_trace(buf.data_ptr())
linear_3xtf32(x, w, b, out=out)
torch.cuda.synchronize()
L.msda_b200_linear_set_trace(None)
t = buf.cpu()
n = int((t[:, 0] != 0).sum())
t = t[:n].double()
names = ["start->first full", "full->split", "split->last mma issued", "last issue->accum ready", "accum ready->epilogue done"]
d = [t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3], t[:, 5] - t[:, 4]]
print("CTAs traced:", n, " total cycles per CTA: mean %.0f" % float((t[:, 5] - t[:, 0]).mean()))
for nm, v in zip(names, d):
    print("%-30s mean %8.0f  min %8.0f  max %8.0f cycles" % (nm, float(v.mean()), float(v.min()), float(v.max())))
first = t[:148]
print("first wave only: total mean %.0f ; " % float((first[:, 5] - first[:, 0]).mean()), [round(float((a[:148]).mean())) for a in d])
