import json, os, sys
sys.path.insert(0, os.getcwd())
import numpy, torch
import varsens_b200 as vb
ctx = vb.Context.get(0); ctx.set_timing(True); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
k, n = 50, 1 << 22
p = torch.from_numpy(numpy.random.RandomState(1).permutation(n).astype(numpy.int32)).cuda()
rows = int(16e9 / (k * 8)); buf = torch.empty((rows, k), dtype=torch.float64, device="cuda")
for r0 in (n + 12345, 0):
    ts = []
    for _ in range(4):
        ctx.sample_flat(k, n, p, row_begin=r0, row_end=r0 + rows, out=buf); torch.cuda.synchronize(); ts.append(ctx.last_kernel_ms())
    print("window16GB r0=%d kernel_ms %s frac %.3f" % (r0, [round(t, 3) for t in ts], 16e9 / (min(ts[1:]) * 1e-3) / 1e9 / 6388))
# generation only: a window of ONE flat row per block would need all rows... use a tiny window spanning 2 blocks: rows [n-1, n+1)
small = torch.empty((2, k), dtype=torch.float64, device="cuda")
ts = []
for _ in range(4):
    ctx.sample_flat(k, n, p, row_begin=n - 1, row_end=n + 1, out=small); torch.cuda.synchronize(); ts.append(ctx.last_kernel_ms())
print("generation of all n base rows (2-row window across a block boundary): kernel_ms", [round(t, 3) for t in ts])
# small windows (the torch-objective route materialises 256 MB windows: Objective._evaluate_vectorized)
for nbytes in (1 << 28, 1 << 30, 4 << 30):
    rows = nbytes // (k * 8)
    w = torch.empty((rows, k), dtype=torch.float64, device="cuda")
    for r0 in (0, 3 * n + 1001):
        ts = []
        for _ in range(4):
            ctx.sample_flat(k, n, p, row_begin=r0, row_end=r0 + rows, out=w); torch.cuda.synchronize(); ts.append(ctx.last_kernel_ms())
        print("window %5d MB r0=%d kernel_ms %.3f -> %.0f GB/s" % (nbytes >> 20, r0, min(ts[1:]), nbytes / (min(ts[1:]) * 1e-3) / 1e9))
    del w
