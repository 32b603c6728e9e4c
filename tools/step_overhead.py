"""Host-side cost of one fused step: wall time of plan.run() on a tiny design (k=20, n=4096: ~20 us of GPU work), i.e. the foreign
call, the kernel launch, the stream synchronisation and the result unpacking; plus the same through Context.run_fused."""
import json, os, sys, time
import numpy, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import varsens_b200 as vb
from varsens_b200 import _cabi, saltelli

k, n = 20, 4096
a = [0, .5, 3, 9, 99, 99] + [99.0] * 14
ctx = vb.Context.get(0)
perm = torch.from_numpy(saltelli._reference_permutation(n).astype(numpy.int32)).cuda()
plan = ctx.fused_plan(k, n, perm, _cabi.OBJ_GFUNCTION, a)
out = {}
for name, fn in (("plan.run", plan.run), ("ctx.run_fused", lambda: ctx.run_fused(k, n, perm, _cabi.OBJ_GFUNCTION, a))):
    for _ in range(50):
        fn()
    ts = []
    for _ in range(400):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    out[name + "_us_median"] = round(1e6 * float(numpy.median(ts)), 2)
    out[name + "_us_p10"] = round(1e6 * float(numpy.percentile(ts, 10)), 2)
ctx.set_timing(True)
plan.run()
out["kernel_us_n4096"] = round(1e3 * ctx.last_kernel_ms(), 2)
out["tail_us"] = [round(t / 1e3, 2) for t in ctx.last_tail_ns(k).tolist()]
print(json.dumps(out))
