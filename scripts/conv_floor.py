"""Micro-benchmark: per-tile floor of short-K convs (Cin=64: one K stage) for the tile variants, with/without stats."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asyrp_official_b200 import ops
dev = torch.device("cuda:0")
N, H, W = 16, 256, 256
for Cout in (128, 256, 64):
    for Cin in (64, 128):
        for stats in (1, 0):
            x = torch.randn(N, H, W, Cin, device=dev).half()
            w = (torch.randn(Cout, 9 * Cin, device=dev) / (3 * Cin ** 0.5)).half()
            out = torch.empty(N, H, W, Cout, device=dev, dtype=torch.float16)
            st = ops.new_stats(N, H, W, Cout, dev, True) if stats else None
            op = ops.ConvOp([(x, ops.MODE_3x3)], w, out=out, stats=st)
            for _ in range(3): op.launch()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): op.launch()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            fl = 2.0 * N * H * W * Cout * 9 * Cin
            outputs = N * H * W * Cout
            print(f"Cout={Cout:3d} Cin={Cin:3d} stats={stats}  {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TF/s  out-write {outputs*2/ms/1e9:6.2f} TB/s")
