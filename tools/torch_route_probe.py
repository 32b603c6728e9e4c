"""The torch-objective route: Varsens(@vectorized torch function, ...) -- sample windows generated on the GPU by the
export kernel, the user's torch code evaluates them, vs_indices_from_values reduces the values.  Prints one JSON line:
    python tools/torch_route_probe.py [k] [log2 n]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import varsens_b200 as vb

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
a = torch.tensor(([0, .5, 3, 9, 99, 99] + [99.0] * 44)[:k], dtype=torch.float64, device="cuda")


@vb.vectorized
def g_torch(X):                                   # (rows, k) CUDA tensor -> (rows,)
    return torch.prod((torch.abs(4.0 * X - 2.0) + a) / (1.0 + a), dim=1)


ts = []
for _ in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v = vb.Varsens(g_torch, lambda x: x, k, n, verbose=False)
    s0 = float(v.sens[0, 0])
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
ref = vb.Varsens(vb.GFunction(a.cpu().numpy()), lambda x: x, k, n, verbose=False)
evals = 2 * n * (1 + k)
print(json.dumps({"route": "torch @vectorized objective", "k": k, "n": n, "evals": evals, "seconds": min(ts), "evals_per_s": evals / min(ts),
                  "sample_bytes_generated": evals * k * 8, "sens0": s0, "max_abs_diff_vs_fused_sens": float(numpy.abs(v.sens - ref.sens).max())}))
