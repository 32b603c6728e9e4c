"""Fused-kernel variant 5 (8 E + 4 S warps) against variant 6 (12 E + 4 S) for several k at n = 2^23 (picks the default)."""
import os, sys, numpy, torch
sys.path.insert(0, os.getcwd())
import varsens_b200 as vb
from varsens_b200 import _cabi
ctx = vb.Context.get(0)
ctx.set_timing(True)
n = 1 << 23
perm = torch.from_numpy(numpy.random.RandomState(1).permutation(n).astype(numpy.int32)).cuda()
for k in (4, 6, 8, 10, 11, 12, 14, 15, 16, 18, 20):
    a = ([0, .5, 3, 9, 99, 99] + [99.0] * 14)[:k]
    out = []
    for v in ("5", "6"):
        os.environ["VS_FUSED_VARIANT"] = v
        ctx.reload_env()
        ts = []
        for _ in range(4):
            r = ctx.run_fused(k, n, perm, _cabi.OBJ_GFUNCTION, a)
            ts.append(ctx.last_kernel_ms())
        out.append(min(ts[1:]))
    print(k, "v5 %.3f ms  v6 %.3f ms" % tuple(out), "->", "5" if out[0] < out[1] else "6")
