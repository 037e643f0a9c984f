import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asyrp_official_b200 import ops
dev = torch.device("cuda:0")
N, H, W, Cin, Cout = 16, 256, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 64, 128
x = torch.randn(N, H, W, Cin, device=dev).half()
w = (torch.randn(Cout, 9 * Cin, device=dev) / (3 * Cin ** 0.5)).half()
out = torch.empty(N, H, W, Cout, device=dev, dtype=torch.float16)
op = ops.ConvOp([(x, ops.MODE_3x3)], w, out=out, stats=ops.new_stats(N, H, W, Cout, dev, True))
for _ in range(3): op.launch()
torch.cuda.synchronize()
