"""Export-mode timing probe: python tools/export_trace.py [gen|w16|w64|shard ...]   (VS_TRACE=file adds CTA 0's phase stamps,
VS_EXPORT_COPIES=1 / VS_EXPORT_SLOW_GEN=1 select the comparison forms).  C4 geometry: k=50, n=2^22."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import varsens_b200 as vb
ctx = vb.Context.get(0); ctx.set_timing(True); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
k, n = 50, 1 << 22
p = torch.from_numpy(numpy.random.RandomState(1).permutation(n).astype(numpy.int32)).cuda()
tag = "copies=%s slowgen=%s" % (os.environ.get("VS_EXPORT_COPIES", "2"), os.environ.get("VS_EXPORT_SLOW_GEN", "0"))
for what in sys.argv[1:] or ["gen", "w16"]:
    if what == "gen":
        out = torch.empty((2, k), dtype=torch.float64, device="cuda"); r0, r1 = n - 1, n + 1
    elif what == "full":
        rows = 2 * n * (1 + k); out = torch.empty((rows, k), dtype=torch.float64, device="cuda"); r0, r1 = 0, rows
    elif what in ("w16", "w64", "w1"):
        rows = int({"w16": 16e9, "w64": 64e9, "w1": 1e9}[what] / (k * 8)); out = torch.empty((rows, k), dtype=torch.float64, device="cuda"); r0, r1 = n + 12345, n + 12345 + rows
    ts = []
    for _ in range(3):
        if what == "shard":
            rows = n // 8
            out = torch.empty((2 + 2 * k, rows, k), dtype=torch.float64, device="cuda") if _ == 0 else out
            ctx.sample_flat_shard(k, n, p, 0, rows, out=out)
        else:
            ctx.sample_flat(k, n, p, row_begin=r0, row_end=r1, out=out)
        torch.cuda.synchronize(); ts.append(ctx.last_kernel_ms())
    print("%s %-6s kernel_ms %s  (%.0f GB/s)" % (tag, what, [round(t, 3) for t in ts], out.numel() * 8 / (min(ts) * 1e-3) / 1e9))
    del out
