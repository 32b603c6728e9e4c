"""One launch (after a warm-up launch) of every kernel that carries a claim, for `ncu`:
    python tools/profile_kernels.py [fused|ishigami|export|evalpf|rk4|gram20|gram50|gramreg|all]
Prints one JSON line per kernel with the library's own event time.  Run it plain first, then under
    ncu --set full --clock-control none -k regex:<kernel> -s 1 -c 1 ...   (the warm-up launch is skipped with -s)."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import varsens_b200 as vb
from varsens_b200 import _cabi, saltelli

which = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = vb.Context.get(0)
ctx.set_timing(True)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
dev = torch.device("cuda", 0)


def perm_dev(n):
    return torch.from_numpy(saltelli._reference_permutation(n).astype(numpy.int32)).to(dev)


def report(name, **kw):
    torch.cuda.synchronize()
    kw.update(kernel=name, kernel_ms=ctx.last_kernel_ms())
    print(json.dumps(kw))


A20 = [0, .5, 3, 9, 99, 99] + [99.0] * 14
if which in ("fused", "all"):
    n = 1 << 24
    p = perm_dev(n)
    for _ in range(2):
        ctx.run_fused(20, n, p, _cabi.OBJ_GFUNCTION, A20)
    report("fused_wsd_kernel<20,GFunctionReg,...,2,1,26>", n=n, k=20)
if which in ("ishigami", "all"):
    n = 1 << 22
    p = perm_dev(n)
    sc = _cabi.Scale(_cabi.SCALE_LINEAR, numpy.full(3, -math.pi), numpy.full(3, math.pi))
    for _ in range(2):
        ctx.run_fused(3, n, p, _cabi.OBJ_ISHIGAMI, [7.0, 0.1], scale=sc)
    report("fused_wsd_kernel<3,IshigamiReg,...>", n=n, k=3)
if which in ("export", "all"):
    k, n = 50, 1 << 22
    p = perm_dev(n)
    rows = n // 8
    out = torch.empty((2 + 2 * k, rows, k), dtype=torch.float64, device=dev)
    for _ in range(2):
        ctx.sample_flat_shard(k, n, p, 0, rows, out=out)
    report("sample_flat_bulk_kernel (C4, base rows [0, n/8): 21.39 GB)", bytes=out.numel() * 8)
    del out
if which in ("evalpf", "all"):
    k, n = 50, 1 << 22
    p = perm_dev(n)
    vals = torch.empty((2 + 2 * k, n), dtype=torch.float64, device=dev)
    a50 = [0, .5, 3, 9, 99, 99] + [99.0] * 44
    for _ in range(2):
        ctx.eval_values(k, n, p, _cabi.OBJ_GFUNCTION, a50, out=vals)
    report("eval_values_pf_kernel<GFunction> (C4 second-order block, values to HBM)", bytes_written=vals.numel() * 8)
    for _ in range(2):
        part = ctx.partials_from_values(k, 1, n, vals, shift=[0.0])
    report("gram_mma_kernel k=50 (C4 second-order block, Gram of the values)", bytes_read=vals.numel() * 8)
    del vals
if which in ("rk4", "all"):
    k, n = 20, 1 << 18
    p = perm_dev(n)
    ref = numpy.array([1.0] * 10 + [0.5] * 10)
    sc = _cabi.Scale(_cabi.SCALE_POWER, ref / 10.0, ref * 10.0)
    vals = torch.empty((2 + 2 * k, n), dtype=torch.float64, device=dev)
    for _ in range(2):
        ctx.eval_values(k, n, p, _cabi.OBJ_RK4_CHAIN, [0.01, 1000], scale=sc, out=vals)
    report("eval_values_kernel<RK4Chain<10>> (C5)", trajectories=vals.numel())
    del vals
if which in ("gram20", "all"):
    k, n = 20, 1 << 22
    vals = torch.rand((2 + 2 * k, n), dtype=torch.float64, device=dev) + 1.0
    for _ in range(2):
        ctx.partials_from_values(k, 1, n, vals, shift=[1.5])
    report("gram_mma_kernel k=20 n=2^22", bytes_read=vals.numel() * 8)
    del vals
if which in ("gramreg", "all"):
    k, n, l = 6, 1 << 20, 3
    vals = torch.rand(((2 + 2 * k) * n, l), dtype=torch.float64, device=dev) + 1.0
    for _ in range(2):
        ctx.partials_from_values(k, l, n, vals, shift=[1.5] * l)
    report("gram kernel for l=3 outputs, k=6 n=2^20", bytes_read=vals.numel() * 8)
    del vals
