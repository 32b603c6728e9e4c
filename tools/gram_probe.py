"""Tensor-path Gram timing probe: python tools/gram_probe.py   (VS_GRAM_WARPS=15 restores the 15-warp 2x2 form)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import varsens_b200 as vb
ctx = vb.Context.get(0); ctx.set_timing(True); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for k, n, l in ((50, 1 << 22, 1), (30, 1 << 22, 1), (24, 1 << 22, 1), (100, 1 << 20, 1), (20, 1 << 20, 3), (20, 1 << 22, 1)):
    vals = torch.rand(((2 + 2 * k) * n, l), dtype=torch.float64, device="cuda") + 1.0
    ts = []
    for _ in range(4):
        ctx.partials_from_values(k, l, n, vals, shift=[1.5] * l); torch.cuda.synchronize(); ts.append(ctx.last_kernel_ms())
    nbytes = vals.numel() * 8
    print("warps=%s k=%d n=2^%d l=%d kernel_ms %.3f  (%.0f GB/s)" % (os.environ.get("VS_GRAM_WARPS", "31"), k, n.bit_length() - 1, l, min(ts[1:]), nbytes / (min(ts[1:]) * 1e-3) / 1e9))
    del vals
