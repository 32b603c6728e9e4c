import os, sys, numpy, torch
sys.path.insert(0, os.getcwd())
import varsens_b200 as vb
from varsens_b200 import _cabi
ctx = vb.Context.get(0)
for k, n in ((50, 1 << 22), (24, 1 << 22)):
    perm = torch.from_numpy(numpy.random.RandomState(1).permutation(n).astype(numpy.int32)).cuda()
    a = ([0, .5, 3, 9, 99, 99] + [99.0] * 44)[:k]
    for _ in range(2):
        r = ctx.run_fused(k, n, perm, _cabi.OBJ_GFUNCTION, a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = ctx.run_fused(k, n, perm, _cabi.OBJ_GFUNCTION, a); e1.record(); torch.cuda.synchronize()
    print("two-phase k=%d n=2^22: %.2f ms" % (k, e0.elapsed_time(e1)), "var_y", float(r.var_y[0]), "sens_2[0,1]", float(r.sens_2[0, 0, 1, 0]))
