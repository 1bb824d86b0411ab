import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gomatching_b200.projections import linear_3xtf32
dev = torch.device("cuda", 0)
M, N, K = 19160 * 8, int(sys.argv[1]) if len(sys.argv) > 1 else 256, 256
w = torch.randn(N, K, device=dev) / 16; b = torch.randn(N, device=dev)
xs = [torch.randn(M, K, device=dev) for _ in range(2)]
outs = [torch.empty(M, N, device=dev) for _ in range(2)]
for i in range(4):
    linear_3xtf32(xs[i % 2], w, b, out=outs[i % 2])
torch.cuda.synchronize()
