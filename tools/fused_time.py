"""Kernel time of the fused C3 step (k=20, n=2^24, resident permutation) for the library selected with VS_LIB, plus a
digest of the result bits (variants must agree bit for bit).  Used by tools/sweep_libs.sh for kernel experiments."""
import hashlib, json, os, sys
import numpy, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import varsens_b200 as vb
from varsens_b200 import _cabi, saltelli

k = int(os.environ.get("VS_K", "20"))
n = 1 << int(os.environ.get("VS_LOGN", "24"))
a = ([0, .5, 3, 9, 99, 99] + [99.0] * 14)[:k]
ctx = vb.Context.get(0)
ctx.set_timing(True)
perm = torch.from_numpy(saltelli._reference_permutation(n).astype(numpy.int32)).cuda()
flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
ts, res = [], None
for it in range(8):
    flush.zero_()
    torch.cuda.synchronize()
    res = ctx.run_fused(k, n, perm, _cabi.OBJ_GFUNCTION, a)
    ts.append(ctx.last_kernel_ms())
h = hashlib.sha256()
for name in ("E_2", "var_y", "U_j", "U_nj", "sens", "sens_t", "sens_2", "sens_2n"):
    h.update(numpy.ascontiguousarray(getattr(res, name)).tobytes())
tail = ctx.last_tail_ns(k).tolist()
print(json.dumps({"lib": os.path.basename(_cabi.LIB_PATH), "k": k, "n": n, "kernel_ms_min": min(ts[2:]), "kernel_ms_med": float(numpy.median(ts[2:])),
                  "tail_us": [round(t / 1e3, 1) for t in tail], "digest": h.hexdigest()[:16], "sens0": float(res.sens[0, 0])}))
