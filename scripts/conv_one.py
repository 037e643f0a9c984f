"""Launch one or more synthetic convs three times each (for ncu captures / quick CUDA-event timing).
usage: conv_one.py SPEC [SPEC ...]   SPEC = H,Cin,Cout[,nseg[,fused]]   (3x3 segments, batch 16)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asyrp_official_b200 import ops
dev = torch.device("cuda:0")
N = 16
keep = []
for spec in sys.argv[1:] or ["256,64,128"]:
    f = [int(v) for v in spec.split(",")]
    H, Cin, Cout = f[:3]
    nseg = f[3] if len(f) > 3 else 1
    fused = f[4] if len(f) > 4 else 0
    W = H
    segs = []
    for _ in range(nseg):
        x = torch.randn(N, H, W, Cin, device=dev).half()
        aff = torch.stack([torch.ones(N, Cin, device=dev), torch.zeros(N, Cin, device=dev)], -1).contiguous()
        segs.append((x, ops.MODE_3x3, aff, 0, 1) if fused else (x, ops.MODE_3x3))
    w = (torch.randn(Cout, 9 * Cin * nseg, device=dev) / (3 * (Cin * nseg) ** 0.5)).half()
    out = torch.empty(N, H, W, Cout, device=dev, dtype=torch.float16)
    op = ops.ConvOp(segs, w, out=out, stats=ops.new_stats(N, H, W, Cout, dev, True))
    keep.append((op, segs, w, out))
    for _ in range(3): op.launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): op.launch()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * N * H * W * Cout * 9 * Cin * nseg
    print(f"{spec}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TF/s")
