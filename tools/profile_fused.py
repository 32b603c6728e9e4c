"""One fused launch of the C3 kernel at reduced n (for ncu --set full; not a benchmark)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import varsens_b200 as vb
from varsens_b200 import _cabi
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 1
A = [0, 0.5, 3, 9, 99, 99] + [99.0] * 14
ctx = vb.Context.get(0)
ctx.set_timing(True)
perm = torch.from_numpy(numpy.random.RandomState(1).permutation(n).astype(numpy.int32)).cuda()
for _ in range(3):
    r = ctx.run_fused(20, n, perm, _cabi.OBJ_GFUNCTION, A, flags=flags)
print("n=%d flags=%d kernel_ms=%.4f var_y=%.12f" % (n, flags, ctx.last_kernel_ms(), r.var_y[0]))
