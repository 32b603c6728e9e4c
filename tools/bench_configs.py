"""Timings of the BASELINE.json configurations other than the headline (C3 is bench.py): C1 parity config,
C2 Ishigami fused, C4 export mode (HBM-write roofline) + its k=50 Gram through the two-phase path, C5 RK4 chain.
Prints one JSON document; run on a B200:  python tools/bench_configs.py > gpurun_out/configs.json"""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import varsens_b200 as vb
from varsens_b200 import _cabi

ctx = vb.Context.get(0)
ctx.set_timing(True)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
dev = torch.device("cuda", 0)
out = {"gpu": torch.cuda.get_device_name(0)}
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
HBM = float(peaks.get("hbm_gbs", 6650.0))
out["hbm_peak_gbs"] = HBM
out["hbm_peak_source"] = "MEASURED_PEAKS.json (copy, read+write)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
out["fp64_peak_tflops"] = ctx.measure_fp64_peak()


def perm_dev(n):
    return torch.from_numpy(numpy.random.RandomState(1).permutation(n).astype(numpy.int32)).to(dev)


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(numpy.median(ts)), r


# ---- C1: README g-function k=6 n=1024 (parity config; latency-bound)
a6 = [0, .5, 3, 9, 99, 99]
p = perm_dev(1024)
ms, r = timeit(lambda: ctx.run_fused(6, 1024, p, _cabi.OBJ_GFUNCTION, a6))
out["C1_gfunction_k6_n1024"] = {"ms": ms, "evals": 14336, "var_y": float(r.var_y[0]), "sens": r.sens[:, 0].tolist()}

# ---- C2: Ishigami k=3 n=2^22, linear scaling to [-pi, pi], first/total/second order, fused
n = 1 << 22
p = perm_dev(n)
sc = _cabi.Scale(_cabi.SCALE_LINEAR, numpy.full(3, -math.pi), numpy.full(3, math.pi))
ms, r = timeit(lambda: ctx.run_fused(3, n, p, _cabi.OBJ_ISHIGAMI, [7.0, 0.1], scale=sc))
out["C2_ishigami_k3_n2p22"] = {"ms": ms, "kernel_ms": ctx.last_kernel_ms(), "evals": 2 * n * 4, "evals_per_s": 2 * n * 4 / (ms * 1e-3),
                               "var_y": float(r.var_y[0]), "sens": r.sens[:, 0].tolist(), "sens_t": r.sens_t[:, 0].tolist(),
                               "sens_2_02": float(r.sens_2[0, 0, 2, 0]),
                               "analytic": {"var_y": 13.8446, "sens": [0.3139, 0.4424, 0.0], "sens_t": [0.5576, 0.4424, 0.2437]}}

# ---- C4: export mode k=50, n=2^22: windows of the flat matrix at HBM write bandwidth
k, n = 50, 1 << 22
p = perm_dev(n)
total_rows = 2 * n * (1 + k)
c4 = {"total_bytes": total_rows * k * 8}
free, _tot = torch.cuda.mem_get_info()
for label, gb in (("window_16GB", 16), ("window_64GB", 64), ("full_171GB", total_rows * k * 8 / 1e9)):
    rows = min(total_rows, int(gb * 1e9 / (k * 8)))
    need = rows * k * 8
    free, _tot = torch.cuda.mem_get_info()
    if need > free - (3 << 30):
        c4[label] = {"skipped": "needs %.1f GB, %.1f GB free" % (need / 1e9, free / 1e9)}
        continue
    buf = torch.empty((rows, k), dtype=torch.float64, device=dev)
    r0 = 0 if rows == total_rows else n + 12345                      # a window that straddles block boundaries
    ms, _ = timeit(lambda: ctx.sample_flat(k, n, p, row_begin=r0, row_end=r0 + rows, out=buf) is None, reps=3, warm=1)
    kms = ctx.last_kernel_ms()
    c4[label] = {"rows": rows, "bytes": need, "ms": ms, "kernel_ms": kms, "write_gbs": need / (kms * 1e-3) / 1e9,
                 "frac_of_hbm_peak": need / (kms * 1e-3) / 1e9 / HBM, "checksum": float(buf[::max(1, rows // 4096)].sum())}
    del buf
    torch.cuda.empty_cache()
# linear-scaled window as well
rows = int(16e9 / (k * 8))
buf = torch.empty((rows, k), dtype=torch.float64, device=dev)
sc = _cabi.Scale(_cabi.SCALE_LINEAR, numpy.linspace(-1, 0, k), numpy.linspace(1, 5, k))
ms, _ = timeit(lambda: ctx.sample_flat(k, n, p, scale=sc, row_begin=0, row_end=rows, out=buf) is None, reps=3, warm=1)
c4["window_16GB_linear"] = {"bytes": rows * k * 8, "kernel_ms": ctx.last_kernel_ms(), "write_gbs": rows * k * 8 / (ctx.last_kernel_ms() * 1e-3) / 1e9}
del buf
torch.cuda.empty_cache()
# the k x k second-order block from g-function values (two-phase path: values to HBM, then the Gram)
a50 = [0, .5, 3, 9, 99, 99] + [99.0] * 44
ms, r = timeit(lambda: ctx.run_fused(k, n, p, _cabi.OBJ_GFUNCTION, a50), reps=3, warm=1)
c4["gram_k50_two_phase"] = {"ms": ms, "evals": total_rows, "evals_per_s": total_rows / (ms * 1e-3), "var_y": float(r.var_y[0]),
                            "sens_2_01": float(r.sens_2[0, 0, 1, 0])}
out["C4_export_k50_n2p22"] = c4

# ---- C5: RK4 mass-action chain, k=20, magnitude scaling (orders=1), n=2^18, dt=0.01, 1000 steps
k, n = 20, 1 << 18
p = perm_dev(n)
ref = numpy.array([1.0] * 10 + [0.5] * 10)
sc = _cabi.Scale(_cabi.SCALE_POWER, ref / 10.0, ref * 10.0)
ms, r = timeit(lambda: ctx.run_fused(k, n, p, _cabi.OBJ_RK4_CHAIN, [0.01, 1000], scale=sc), reps=3, warm=1)
traj = 2 * n * (1 + k)
flops = traj * 1000.0 * 4 * (10 * 4 + 11 * 4)        # 4 stages x (10 links x (mul, fma, sub = 4 flops) + 11 species x (axpy 2 + accumulate 2))
out["C5_rk4_chain_k20_n2p18"] = {"ms": ms, "trajectories": traj, "trajectories_per_s": traj / (ms * 1e-3), "flops_est": flops,
                                 "tflops_est": flops / (ms * 1e-3) / 1e12, "var_y": float(r.var_y[0]), "sens": r.sens[:, 0].tolist()}

# ---- estimators on given values (Objective(objective_vals=...) route), k=20 n=2^22
k, n = 20, 1 << 22
vals = torch.rand((2 * n * (1 + k), 1), dtype=torch.float64, device=dev) + 1.0
ms, r = timeit(lambda: ctx.indices_from_values(k, 1, n, n, vals), reps=5, warm=2)
out["indices_from_values_k20_n2p22"] = {"ms": ms, "kernel_ms": ctx.last_kernel_ms(), "bytes_read": vals.numel() * 8,
                                        "read_gbs": vals.numel() * 8 / (ctx.last_kernel_ms() * 1e-3) / 1e9,
                                        "frac_of_hbm_peak": vals.numel() * 8 / (ctx.last_kernel_ms() * 1e-3) / 1e9 / HBM}
print(json.dumps(out, indent=1))
