"""One vs_partials_from_values launch on resident random values (for ncu --set full; not a benchmark).
usage: python tools/gram_one.py <k> <log2 rows>"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from varsens_b200 import Context
ctx = Context(0)
ctx.set_timing(True)
k, rows = int(sys.argv[1]), 1 << int(sys.argv[2])
m = 2 + 2 * k
vals = torch.rand(m * rows, dtype=torch.float64, device="cuda") + 1.0
out = torch.empty(4 + m * (m + 1) // 2, dtype=torch.float64, device="cuda")
for _ in range(2):
    ctx.partials_from_values(k, 1, rows, vals, shift=[1.5], out=out)
ctx.synchronize()
print("ms", ctx.last_kernel_ms())
